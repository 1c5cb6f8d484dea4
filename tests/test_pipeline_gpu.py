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


def test_block_overlap_evaluation_and_scene_merge_with_the_network(pkg):
    """SURVEY.md 8(f) N4 on the GPU: blocks of one synthetic scene -> predict_blocks_with_overlap around
    SPH3D_s3dis.get_model (inference mode) -> per-block summed logits -> scene merge (io/s3dis_merge.py =
    post-merging/s3dis_merge.m) -> per-point labels of the scene and of its full-resolution cloud."""
    ev, mg, u, M = pkg.io.s3dis_eval, pkg.io.s3dis_merge, pkg.sph3gcn_util, pkg.models
    rng = np.random.default_rng(41)
    P, num_point, C = 3000, 512, 13
    scene_xyz = (rng.random((P, 3)) * np.array([3.0, 1.5, 1.0])).astype(np.float32)        # a 3 m x 1.5 m strip
    scene_rgb = rng.random((P, 3)).astype(np.float32)
    scene_gt = rng.integers(0, C, P).astype(np.int32)
    # two 1.5 m blocks with 0.5 m of context on either side (make_tfrecord_s3dis.py: inner = inside the block proper)
    blocks = []
    for x0 in (0.0, 1.5):
        member = np.nonzero((scene_xyz[:, 0] >= x0 - 0.5) & (scene_xyz[:, 0] < x0 + 2.0))[0]
        inner = ((scene_xyz[member, 0] >= x0) & (scene_xyz[member, 0] < x0 + 1.5)).astype(np.float32)
        rows = np.concatenate([scene_xyz[member] - np.array([x0, 0, 0], np.float32), scene_rgb[member],
                               scene_gt[member, None].astype(np.float32), inner[:, None]], axis=1)
        blocks.append((member, rows))
    maxn = max(len(r) for _, r in blocks)
    padded = np.full((len(blocks), maxn, 8), -1.0, dtype=np.float32)                      # padding rows carry -1 (padded_batch)
    for i, (_, rows) in enumerate(blocks):
        padded[i, :len(rows)] = rows
    cfg = M.configs.s3dis(num_point)
    u.reset_variables()
    calls = []

    def predict_fn(batch_input):
        with torch.no_grad():
            u.clear_collections()
            pred, _ = M.SPH3D_s3dis.get_model(torch.from_numpy(batch_input).to("cuda:0"), False, cfg)
        calls.append(batch_input.shape)
        return pred.cpu().numpy()

    summed, counts, rounds = ev.predict_blocks_with_overlap(padded, num_point, predict_fn, C, rng=np.random.default_rng(42))
    assert rounds == len(calls) >= 2 and all(s == (2, num_point, 6) for s in calls)
    metrics = ev.SegmentationMetrics(C)
    merge_in = []
    for i, (member, rows) in enumerate(blocks):
        inner = rows[:, -1] == 1
        assert (counts[i][inner] > 0).all() and np.isfinite(summed[i]).all()               # every inner point was drawn
        assert (np.abs(summed[i][counts[i] == 0]) == 0).all()
        metrics.update(summed[i], rows[:, -2], rows[:, -1])
        merge_in.append((summed[i], rows[:, -1].astype(np.int32), member))
    res = metrics.result()
    assert 0.0 <= res["accuracy"] <= 1.0 and metrics.seen == sum(int((r[:, -1] == 1).sum()) for _, r in blocks)
    pred, label = mg.merge_scene(P, merge_in, C)
    covered = np.zeros(P, bool)
    for member, rows in blocks:
        covered[member[rows[:, -1] == 1]] = True
    assert covered.all()                                                                   # the two inner strips tile the scene
    assert np.allclose(pred.sum(1), 1.0)              # each scene point is inner to exactly one block: one softmax row each
    assert label.shape == (P,) and label.min() >= 0 and label.max() < C
    # a scene point's merged label = the argmax of its own block's summed logits (softmax of a unit vector keeps the order)
    for (member, rows), s in zip(blocks, summed):
        inner = rows[:, -1] == 1
        assert (label[member[inner]] == s[inner].argmax(1)).all()
    full_xyz = scene_xyz[rng.integers(0, P, 5000)] + rng.normal(0, 1e-4, (5000, 3)).astype(np.float32)
    full_label = mg.propagate_to_full_cloud(scene_xyz, label, full_xyz)
    iou = mg.SceneIoU(C)
    iou.update(full_label, rng.integers(0, C, 5000))
    assert 0.0 <= iou.result()["mean_iou"] <= 1.0
