"""ModelNet40 classification records (/root/reference/modelnet40_cls/train_modelnet.py:118-138; writer
io/make_tfrecord_modelnet.py): a bytes feature xyz_raw (float32 x 3, the same point count in every record) and an int64
feature label.  Batches are plain stacks (dataset.batch, no padding)."""
import numpy as np

from . import tfrecord


def parse_fn(item):
    f = tfrecord.parse_example(item)
    xyz = np.frombuffer(f["xyz_raw"][0], dtype="<f4").reshape(-1, 3)
    return xyz, np.int32(f["label"][0])


def input_fn(filelist, batch_size=16, buffer_size=10000, rng=None, check_crc=True):
    """yields (batch_xyz (b, N, 3) float32, batch_label (b,) int32)"""
    records = tfrecord.shuffled_records(filelist, buffer_size, rng, check_crc)
    for batch in tfrecord.batched((parse_fn(rec) for rec in records), batch_size):
        yield np.stack([x for x, _ in batch]).astype(np.float32), np.asarray([l for _, l in batch], dtype=np.int32)
