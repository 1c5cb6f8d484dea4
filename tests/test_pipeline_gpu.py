"""File -> batch -> network: synthetic S3DIS blocks written as TFRecords (io/tfrecord.py), read back through the block
pipeline of s3dis_seg/train_s3dis.py (io/s3dis_input.py: shuffle, padded batch, resample to num_point, augmentation) and fed
to the SPH3D_s3dis call graph for one training step (SURVEY.md 8(f) N3 -> N1)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tfrecord_to_training_step(pkg, tmp_path):
    tfr, si, u, M = pkg.io.tfrecord, pkg.io.s3dis_input, pkg.sph3gcn_util, pkg.models
    rng = np.random.default_rng(31)
    recs = []
    for n in (1500, 900, 2000, 1100):                            # blocks of different sizes, like real rooms
        xyz = (rng.random((n, 3), dtype=np.float32) * np.float32(1.5)).astype(np.float32)
        recs.append(tfr.make_example({"xyz_raw": xyz.tobytes(), "rgb_raw": rng.random((n, 3), dtype=np.float32).tobytes(),
                                      "seg_label": rng.integers(0, 13, n).astype(np.int32).tobytes(),
                                      "inner_label": (rng.random(n) < 0.7).astype(np.int32).tobytes()}))
    path = str(tmp_path / "blocks.tfrecord")
    tfr.write_records(path, recs)
    num_point = 1024
    cfg = M.configs.s3dis(num_point)
    u.reset_variables()
    losses = []
    for padded in si.input_fn([path], batch_size=2, buffer_size=4, rng=np.random.default_rng(32)):
        inp, lab, inner = si.select_points(padded, num_point, rng=np.random.default_rng(33))
        inp, lab, inner = si.augment_fn(inp, lab, inner, rng=np.random.default_rng(34))
        assert inp.shape == (2, num_point, 6)
        pts = torch.from_numpy(inp).to("cuda:0")
        u.clear_collections()
        for p in u.trainable_variables():
            p.grad = None
        pred, end = M.SPH3D_s3dis.get_model(pts, True, cfg)
        loss = M.SPH3D_s3dis.get_loss(pred, torch.from_numpy(lab).to("cuda:0"), end, torch.from_numpy(inner).to("cuda:0"))
        loss.backward()
        assert pred.shape == (2, num_point, 13) and torch.isfinite(loss)
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in u.trainable_variables())
        losses.append(float(loss))
    assert len(losses) == 2
