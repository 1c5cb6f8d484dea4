"""ShapeNet part-segmentation records (/root/reference/shapenet_seg/train_shapenet.py:155-180; writer
io/make_tfrecord_shapenet.py): bytes features xyz_raw (float32 x 3) and part_label (int32) per shape; batches are padded
with -1 like the S3DIS blocks.  -> rows [x y z part_label]."""
import numpy as np

from . import tfrecord
from .s3dis_input import padded_batch

INPUT_DIM = 3                 # train_shapenet.py:62


def parse_fn(item):
    f = tfrecord.parse_example(item)
    xyz = np.frombuffer(f["xyz_raw"][0], dtype="<f4").reshape(-1, 3)
    part = np.frombuffer(f["part_label"][0], dtype="<i4").reshape(-1, 1)
    if len(xyz) != len(part):
        raise tfrecord.RecordError("feature lengths disagree: %d %d" % (len(xyz), len(part)))
    return np.concatenate((xyz, part.astype(np.float32)), axis=-1)


def input_fn(filelist, batch_size=16, buffer_size=10000, rng=None, check_crc=True):
    records = tfrecord.shuffled_records(filelist, buffer_size, rng, check_crc)
    for batch in tfrecord.batched((parse_fn(rec) for rec in records), batch_size):
        yield padded_batch(batch)
