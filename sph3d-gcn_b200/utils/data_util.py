"""Point-cloud augmentations of the reference's trainers (/root/reference/utils/data_util.py:8-234), vectorised.

Same transformations and parameter defaults; two deliberate differences: every function draws from an explicit
numpy Generator (`rng`, default = a fresh default_rng()) instead of the global numpy state, and a whole batch is
transformed with one batched product instead of a Python loop per cloud.  All functions return new float32 arrays of
shape (B, N, 3) (the reference's in-place shift / scale also return their argument).  Row-vector convention of the
reference: points are multiplied from the left, p' = p R.
"""
import numpy as np


def _rng(rng):
    return np.random.default_rng() if rng is None else rng


def rot_x(angle):
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float32)


def rot_y(angle):
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float32)


def rot_z(angle):
    c, s = np.cos(angle), np.sin(angle)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)


def _apply(batch_data, matrices):
    return np.einsum("bnc,bcd->bnd", np.asarray(batch_data, np.float32), matrices.astype(np.float32)).astype(np.float32)


def shuffle_data(data, labels, rng=None):
    """permute the clouds of a batch; -> (data, labels, permutation)"""
    idx = _rng(rng).permutation(len(labels))
    return data[idx, ...], labels[idx], idx


def shuffle_points(batch_data, rng=None):
    """one permutation of the point order for the whole batch (changes what FPS and the first-K ball query pick)"""
    return batch_data[:, _rng(rng).permutation(batch_data.shape[1]), :]


def shuffle_points_and_label(batch_data, batch_label, rng=None):
    idx = _rng(rng).permutation(batch_data.shape[1])
    return batch_data[:, idx, :], batch_label[:, idx]


def rotate_point_cloud(batch_data, max_angle=2 * np.pi, rng=None):
    """one uniform rotation about the up (z) axis per cloud"""
    angles = _rng(rng).uniform(size=batch_data.shape[0]) * max_angle
    return _apply(batch_data, np.stack([rot_z(a) for a in angles]))


def rotate_point_cloud_by_angle(batch_data, rotation_angle):
    B = batch_data.shape[0]
    return _apply(batch_data, np.broadcast_to(rot_z(rotation_angle), (B, 3, 3)))


def _perturbation_matrices(B, angle_sigma, angle_clip, rng):
    ang = np.clip(angle_sigma * _rng(rng).standard_normal((B, 3)), -angle_clip, angle_clip)
    return np.stack([rot_z(a[2]) @ rot_y(a[1]) @ rot_x(a[0]) for a in ang])


def rotate_perturbation_point_cloud(batch_data, angle_sigma=0.06, angle_clip=0.18, rng=None):
    """small random rotation R = Rz Ry Rx per cloud, angles ~ N(0, sigma) clipped to +-clip"""
    return _apply(batch_data, _perturbation_matrices(batch_data.shape[0], angle_sigma, angle_clip, rng))


def rotate_point_cloud_with_normal(batch_xyz_normal, max_angle=2 * np.pi, rng=None):
    """(B, N, 6) = xyz + normal: both halves get the cloud's z rotation"""
    angles = _rng(rng).uniform(size=batch_xyz_normal.shape[0]) * max_angle
    R = np.stack([rot_z(a) for a in angles])
    return np.concatenate([_apply(batch_xyz_normal[:, :, 0:3], R), _apply(batch_xyz_normal[:, :, 3:6], R)], axis=2)


def rotate_perturbation_point_cloud_with_normal(batch_data, angle_sigma=0.06, angle_clip=0.18, rng=None):
    R = _perturbation_matrices(batch_data.shape[0], angle_sigma, angle_clip, rng)
    return np.concatenate([_apply(batch_data[:, :, 0:3], R), _apply(batch_data[:, :, 3:6], R)], axis=2)


def jitter_point_cloud(batch_data, sigma=0.01, clip=0.02, rng=None):
    """independent N(0, sigma) noise per coordinate, clipped to +-clip"""
    assert clip > 0
    noise = np.clip(sigma * _rng(rng).standard_normal(batch_data.shape), -clip, clip)
    return (np.asarray(batch_data, np.float32) + noise).astype(np.float32)


def shift_point_cloud(batch_data, shift_range=0.1, rng=None):
    shifts = _rng(rng).uniform(-shift_range, shift_range, (batch_data.shape[0], 1, 3))
    return (np.asarray(batch_data, np.float32) + shifts).astype(np.float32)


def random_scale_point_cloud(batch_data, scale_low=0.8, scale_high=1.25, rng=None):
    scales = _rng(rng).uniform(scale_low, scale_high, (batch_data.shape[0], 1, 1))
    return (np.asarray(batch_data, np.float32) * scales).astype(np.float32)
