"""S3DIS block pipeline of the reference's trainer (/root/reference/s3dis_seg/train_s3dis.py), without TensorFlow.

  parse_fn        :145-171   one Example -> (n, 8) float32 rows [x y z r g b seg_label inner_label]
  input_fn        :174-182   TFRecord files -> shuffled stream -> padded batches (pad value -1, no drop_remainder)
  select_points   :321-350   strip the padding, draw exactly num_point points per block (with replacement only when the
                             block holds fewer), split into input / label / inner mask
  augment_fn      :113-142   shuffle clouds and point order; first third of the batch rotated about z + perturbed,
                             second third jittered
ScanNet blocks carry the same features and the same pipeline (scannet_seg/train_scannet.py:130-160, INPUT_DIM = 6;
io/make_tfrecord_scannet.py:188-194), so this module reads them too.
The arrays `select_points` returns are what the model call graph consumes (models/SPH3D_s3dis.get_model / get_loss).
"""
import numpy as np

from . import tfrecord
from ..utils import data_util

INPUT_DIM = 6                 # xyz + rgb (train_s3dis.py:57)


def parse_fn(item):
    f = tfrecord.parse_example(item)
    col = lambda name, dt, w: np.frombuffer(f[name][0], dtype=dt).reshape(-1, w)
    xyz, rgb = col("xyz_raw", "<f4", 3), col("rgb_raw", "<f4", 3)
    seg, inner = col("seg_label", "<i4", 1), col("inner_label", "<i4", 1)
    if not (len(xyz) == len(rgb) == len(seg) == len(inner)):
        raise tfrecord.RecordError("feature lengths disagree: %d %d %d %d" % (len(xyz), len(rgb), len(seg), len(inner)))
    return np.concatenate((xyz, rgb, seg.astype(np.float32), inner.astype(np.float32)), axis=-1)


def padded_batch(items, pad_value=-1.0):
    """tf.data padded_batch(padded_shapes=(None, INPUT_DIM+2), padding_values=-1.0) of a list of (n_i, D) arrays"""
    longest = max(len(it) for it in items)
    out = np.full((len(items), longest, items[0].shape[1]), pad_value, dtype=np.float32)
    for b, it in enumerate(items):
        out[b, :len(it)] = it
    return out


def input_fn(filelist, batch_size=16, buffer_size=10000, rng=None, check_crc=True):
    """Generator of padded batches over all records of `filelist`, shuffled through a `buffer_size` reservoir like
    tf.data.Dataset.shuffle; the last batch may be smaller (drop_remainder=False)."""
    records = tfrecord.shuffled_records(filelist, buffer_size, rng, check_crc)
    for batch in tfrecord.batched((parse_fn(rec) for rec in records), batch_size):
        yield padded_batch(batch)


def select_points(padded_all, num_point, rng=None):
    """-> batch_input (b, num_point, INPUT_DIM) float32, batch_label (b, num_point) int32, batch_inner (b, num_point) int32"""
    rng = np.random.default_rng() if rng is None else rng
    b = padded_all.shape[0]
    batch_input = np.zeros((b, num_point, INPUT_DIM), dtype=np.float32)
    batch_label = np.zeros((b, num_point), dtype=np.int32)
    batch_inner = np.zeros((b, num_point), dtype=np.int32)
    for i in range(b):
        pad = np.nonzero(padded_all[i, :, -1] < 0)[0]                  # the inner mask is 0/1: -1 marks padding
        num = padded_all.shape[1] if len(pad) == 0 else int(pad[0])
        if num == 0:
            raise ValueError("empty block in batch")
        pick = rng.choice(num, num_point, replace=num < num_point)
        batch_input[i] = padded_all[i, pick, 0:-2]
        batch_label[i] = padded_all[i, pick, -2]
        batch_inner[i] = padded_all[i, pick, -1]
    return batch_input, batch_label, batch_inner


def augment_fn(batch_input, batch_label, batch_inner, rng=None):
    rng = np.random.default_rng() if rng is None else rng
    bsize, num_point, _ = batch_input.shape
    order = rng.permutation(bsize)
    batch_input, batch_label, batch_inner = batch_input[order], batch_label[order], batch_inner[order]
    order = rng.permutation(num_point)
    batch_input, batch_label, batch_inner = batch_input[:, order], batch_label[:, order], batch_inner[:, order]
    batch_input = np.array(batch_input, dtype=np.float32)
    third = int(bsize / 3.0)
    if third:
        xyz = data_util.rotate_point_cloud(batch_input[0:third, :, 0:3], rng=rng)
        batch_input[0:third, :, 0:3] = data_util.rotate_perturbation_point_cloud(xyz, rng=rng)
        batch_input[third:2 * third, :, 0:3] = data_util.jitter_point_cloud(batch_input[third:2 * third, :, 0:3], rng=rng)
    return batch_input, batch_label, batch_inner
