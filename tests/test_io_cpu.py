"""TFRecord container, tf.train.Example codec and the S3DIS block pipeline without TensorFlow
(sph3d-gcn_b200/io, utils/data_util.py).  Pins: CRC-32C test vectors of RFC 3720 B.4; the Example wire format against the
protobuf runtime (dynamic descriptors restating tensorflow/core/example/{example,feature}.proto)."""
import os
import struct

import numpy as np
import pytest


@pytest.fixture()
def tfr(pkg):
    return pkg.io.tfrecord


def test_crc32c_rfc3720_vectors(tfr):
    assert tfr.crc32c(b"123456789") == 0xE3069283
    assert tfr.crc32c(bytes(32)) == 0x8A9136AA
    assert tfr.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tfr.crc32c(bytes(range(32))) == 0x46DD794E
    assert tfr.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert tfr.crc32c(b"") == 0
    rng = np.random.default_rng(0)

    def bytewise(data):                                         # bit-at-a-time definition, reflected polynomial
        c = 0xFFFFFFFF
        for b in data:
            c ^= b
            for _ in range(8):
                c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
        return c ^ 0xFFFFFFFF
    for n in (1, 7, 8, 9, 15, 16, 17, 63, 64, 100, 1001):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert tfr.crc32c(data) == bytewise(data), n
    a, b = b"hello ", b"world"
    assert tfr.crc32c(b, tfr.crc32c(a)) == tfr.crc32c(a + b)    # incremental form
    c = tfr.crc32c(b"abc")
    assert tfr.masked_crc32c(b"abc") == ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _example_messages():
    """tensorflow/core/example/feature.proto + example.proto restated as dynamic protobuf descriptors"""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="sph3d_test_example.proto", package="sph3dtest", syntax="proto3")

    def msg(name, *fields):
        m = fd.message_type.add(name=name)
        for fname, num, typ, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                f.type_name = tname
        return m
    msg("BytesList", ("value", 1, F.TYPE_BYTES, F.LABEL_REPEATED, None))
    msg("FloatList", ("value", 1, F.TYPE_FLOAT, F.LABEL_REPEATED, None))
    msg("Int64List", ("value", 1, F.TYPE_INT64, F.LABEL_REPEATED, None))
    feat = msg("Feature", ("bytes_list", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".sph3dtest.BytesList"),
               ("float_list", 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".sph3dtest.FloatList"),
               ("int64_list", 3, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".sph3dtest.Int64List"))
    feat.oneof_decl.add(name="kind")
    for f in feat.field:
        f.oneof_index = 0
    feats = msg("Features", ("feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".sph3dtest.Features.FeatureEntry"))
    entry = feats.nested_type.add(name="FeatureEntry")
    entry.options.map_entry = True
    entry.field.add(name="key", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    entry.field.add(name="value", number=2, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL, type_name=".sph3dtest.Feature")
    msg("Example", ("features", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".sph3dtest.Features"))
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    if get is None:
        factory = message_factory.MessageFactory(pool)
        get = factory.GetPrototype
    return get(pool.FindMessageTypeByName("sph3dtest.Example"))


def test_example_codec_against_protobuf_runtime(tfr):
    Example = _example_messages()
    rng = np.random.default_rng(1)
    xyz = rng.random((50, 3), dtype=np.float32)
    seg = rng.integers(0, 13, 50).astype(np.int32)
    ex = Example()
    ex.features.feature["xyz_raw"].bytes_list.value.append(xyz.tobytes())
    ex.features.feature["seg_label"].bytes_list.value.append(seg.tobytes())
    ex.features.feature["scene_label"].int64_list.value.extend([7, -3, 1 << 40])
    ex.features.feature["weights"].float_list.value.extend([0.5, -1.25, 3.0])
    got = tfr.parse_example(ex.SerializeToString())            # written by protobuf, read by this codec
    assert set(got) == {"xyz_raw", "seg_label", "scene_label", "weights"}
    assert got["xyz_raw"] == [xyz.tobytes()] and got["seg_label"] == [seg.tobytes()]
    assert got["scene_label"].tolist() == [7, -3, 1 << 40]
    assert got["weights"].tolist() == [0.5, -1.25, 3.0]
    mine = tfr.make_example({"xyz_raw": xyz.tobytes(), "seg_label": seg.tobytes(),
                             "scene_label": np.array([7, -3, 1 << 40]), "weights": np.array([0.5, -1.25, 3.0], np.float32)})
    back = Example()
    back.ParseFromString(mine)                                  # written by this codec, read by protobuf
    assert back.features.feature["xyz_raw"].bytes_list.value[0] == xyz.tobytes()
    assert list(back.features.feature["scene_label"].int64_list.value) == [7, -3, 1 << 40]
    assert list(back.features.feature["weights"].float_list.value) == [0.5, -1.25, 3.0]
    assert back == ex


def _block(rng, n):
    xyz = rng.random((n, 3), dtype=np.float32) * 1.5
    rgb = rng.random((n, 3), dtype=np.float32)
    seg = rng.integers(0, 13, n).astype(np.int32)
    inner = (rng.random(n) < 0.7).astype(np.int32)
    return xyz, rgb, seg, inner


def _record(tfr, xyz, rgb, seg, inner):
    # the feature set of io/make_tfrecord_s3dis.py:233-241
    return tfr.make_example({"rgb_raw": rgb.tobytes(), "seg_label": seg.tobytes(), "inner_label": inner.tobytes(),
                             "index_label": np.arange(len(xyz), dtype=np.int32).tobytes(), "scene_label": np.array([3]),
                             "scene_idx": np.array([11]), "rel_xyz_raw": (xyz / 2).tobytes(), "xyz_raw": xyz.tobytes()})


def test_tfrecord_round_trip_and_corruption(tfr, tmp_path):
    rng = np.random.default_rng(2)
    blocks = [_block(rng, n) for n in (5, 40, 17)]
    path = str(tmp_path / "Area_1.tfrecord")
    tfr.write_records(path, [_record(tfr, *b) for b in blocks])
    recs = list(tfr.read_records(path))
    assert len(recs) == 3
    f = tfr.parse_example(recs[1])
    assert np.array_equal(np.frombuffer(f["xyz_raw"][0], "<f4").reshape(-1, 3), blocks[1][0])
    assert f["scene_idx"].tolist() == [11]
    raw = bytearray(open(path, "rb").read())
    first_len, = struct.unpack("<Q", raw[:8])
    assert struct.unpack("<I", raw[8:12])[0] == tfr.masked_crc32c(bytes(raw[:8]))
    raw[12 + first_len // 2] ^= 0x40                            # flip a payload bit of record 0
    bad = str(tmp_path / "bad.tfrecord")
    open(bad, "wb").write(raw)
    with pytest.raises(tfr.RecordError, match="payload"):
        list(tfr.read_records(bad))
    assert len(list(tfr.read_records(bad, check_crc=False))) == 3
    open(bad, "wb").write(bytes(raw[:-3]))
    with pytest.raises(tfr.RecordError, match="truncated"):
        list(tfr.read_records(bad, check_crc=False))


def test_s3dis_pipeline_semantics(pkg, tfr, tmp_path):
    si = pkg.io.s3dis_input
    rng = np.random.default_rng(3)
    sizes = (30, 100, 64, 7, 55)
    blocks = [_block(rng, n) for n in sizes]
    path = str(tmp_path / "a.tfrecord")
    tfr.write_records(path, [_record(tfr, *b) for b in blocks])
    one = si.parse_fn(next(tfr.read_records(path)))             # train_s3dis.py:145-171
    assert one.shape == (30, 8) and one.dtype == np.float32
    assert np.array_equal(one[:, 0:3], blocks[0][0]) and np.array_equal(one[:, 3:6], blocks[0][1])
    assert np.array_equal(one[:, 6], blocks[0][2].astype(np.float32)) and np.array_equal(one[:, 7], blocks[0][3].astype(np.float32))
    batches = list(si.input_fn([path], batch_size=2, buffer_size=3, rng=np.random.default_rng(4)))
    assert [b.shape[0] for b in batches] == [2, 2, 1]           # drop_remainder=False
    seen = sorted(int((b[i, :, -1] >= 0).sum()) for b in batches for i in range(b.shape[0]))
    assert seen == sorted(sizes)                                # every block once, padding = -1
    for b in batches:
        assert b.shape[2] == si.INPUT_DIM + 2 and (b[b[:, :, -1] < 0] == -1.0).all()
    padded = si.padded_batch([si.parse_fn(r) for r in tfr.read_records(path)])
    assert padded.shape == (5, 100, 8)
    inp, lab, inner = si.select_points(padded, 64, rng=np.random.default_rng(5))   # :321-350
    assert inp.shape == (5, 64, 6) and lab.shape == (5, 64) and inner.shape == (5, 64) and lab.dtype == np.int32
    for i, (xyz, rgb, seg, inn) in enumerate(blocks):
        rows = {tuple(r) for r in np.concatenate([xyz, rgb], 1).tolist()}
        assert all(tuple(r) in rows for r in inp[i].tolist())    # only real points, never padding
        n_unique = len({tuple(r) for r in inp[i].tolist()})
        if len(xyz) >= 64:
            assert n_unique == 64                               # without replacement when the block is large enough
        else:
            assert n_unique <= len(xyz)
        assert set(lab[i].tolist()) <= set(seg.tolist()) and set(inner[i].tolist()) <= {0, 1}
    a_in, a_lab, a_inner = si.augment_fn(inp.copy(), lab.copy(), inner.copy(), rng=np.random.default_rng(6))   # :113-142
    assert a_in.shape == inp.shape and sorted(a_lab.sum(1).tolist()) == sorted(lab.sum(1).tolist())
    assert np.allclose(np.sort(a_in[:, :, 3:6].sum((1, 2))), np.sort(inp[:, :, 3:6].sum((1, 2))), rtol=1e-5)   # rgb only permuted


def test_augmentations(pkg):
    du = pkg.utils.data_util
    rng = np.random.default_rng(7)
    pts = rng.standard_normal((6, 200, 3)).astype(np.float32)
    rot = du.rotate_point_cloud(pts, rng=np.random.default_rng(8))
    assert rot.dtype == np.float32 and np.allclose(rot[:, :, 2], pts[:, :, 2], atol=1e-6)          # about the up axis
    assert np.allclose(np.linalg.norm(rot, axis=2), np.linalg.norm(pts, axis=2), atol=1e-5)
    assert not np.allclose(rot, pts)
    assert np.array_equal(rot, du.rotate_point_cloud(pts, rng=np.random.default_rng(8)))           # reproducible
    quarter = du.rotate_point_cloud_by_angle(pts, np.pi / 2)                                        # p' = p Rz: (x, y) -> (y, -x)
    assert np.allclose(quarter[:, :, 0], pts[:, :, 1], atol=1e-6) and np.allclose(quarter[:, :, 1], -pts[:, :, 0], atol=1e-6)
    per = du.rotate_perturbation_point_cloud(pts, rng=np.random.default_rng(9))
    assert np.allclose(np.linalg.norm(per, axis=2), np.linalg.norm(pts, axis=2), atol=1e-5)
    cosang = (per * pts).sum(2) / np.maximum(np.linalg.norm(pts, axis=2) ** 2, 1e-12)
    assert cosang.min() > np.cos(3 * 0.18 + 1e-3)                                                    # three clipped angles
    jit = du.jitter_point_cloud(pts, rng=np.random.default_rng(10))
    assert np.abs(jit - pts).max() <= 0.02 + 1e-7 and np.abs(jit - pts).std() > 0.005
    sh = du.shift_point_cloud(pts, rng=np.random.default_rng(11))
    d = sh - pts
    assert np.allclose(d, d[:, :1, :], atol=1e-6) and np.abs(d).max() <= 0.1 + 1e-6
    sc = du.random_scale_point_cloud(pts, rng=np.random.default_rng(12))
    ratio = np.linalg.norm(sc, axis=2) / np.linalg.norm(pts, axis=2)
    assert np.allclose(ratio, ratio[:, :1], rtol=1e-4) and 0.8 <= ratio.min() and ratio.max() <= 1.25
    xn = np.concatenate([pts, pts / np.linalg.norm(pts, axis=2, keepdims=True)], axis=2)
    rn = du.rotate_point_cloud_with_normal(xn, rng=np.random.default_rng(13))
    assert np.allclose((rn[:, :, :3] * rn[:, :, 3:]).sum(2), (xn[:, :, :3] * xn[:, :, 3:]).sum(2), atol=1e-4)
    data, labels, idx = du.shuffle_data(pts, np.arange(6), rng=np.random.default_rng(14))
    assert np.array_equal(data, pts[idx]) and sorted(labels.tolist()) == list(range(6))


def test_block_overlap_evaluation(pkg):
    """evaluate_s3dis_with_overlap.py:245-318: resample until every inner point is covered, sum logits per point, score inner points"""
    ev, si = pkg.io.s3dis_eval, pkg.io.s3dis_input
    rng = np.random.default_rng(21)
    sizes, num_point, ncls = (300, 90, 150), 128, 5
    items = []
    for n in sizes:
        xyz, rgb = rng.random((n, 3), dtype=np.float32), rng.random((n, 3), dtype=np.float32)
        seg = rng.integers(0, ncls, n).astype(np.float32)
        inner = (rng.random(n) < 0.6).astype(np.float32)
        items.append(np.concatenate([xyz, rgb, seg[:, None], inner[:, None]], axis=1))
    padded = si.padded_batch(items)
    assert ev.block_lengths(padded).tolist() == list(sizes)
    lookup = {tuple(np.round(row[:6], 6).tolist()): int(row[6]) for it in items for row in it}
    calls = []

    def oracle_net(batch_input):                                # a "network" that knows the label of every point
        calls.append(batch_input.shape)
        out = np.zeros(batch_input.shape[:2] + (ncls,), dtype=np.float32)
        for b in range(batch_input.shape[0]):
            for j in range(batch_input.shape[1]):
                out[b, j, lookup[tuple(np.round(batch_input[b, j], 6).tolist())]] = 1.0
        return out

    summed, counts, rounds = ev.predict_blocks_with_overlap(padded, num_point, oracle_net, ncls, rng=np.random.default_rng(22))
    assert rounds == len(calls) >= 3 and all(c == (3, num_point, 6) for c in calls)   # 300 points need several draws of 128
    metrics = ev.SegmentationMetrics(ncls)
    for i, it in enumerate(items):
        inner = it[:, 7] == 1
        assert (counts[i][inner] > 0).all()                     # stop criterion: every inner point drawn at least once
        assert counts[i].sum() == rounds * num_point
        drawn = counts[i] > 0
        assert (summed[i][~drawn] == 0).all() and (summed[i][drawn].sum(1) > 0).all()
        metrics.update(summed[i], it[:, 6], it[:, 7])
    res = metrics.result()
    assert res["accuracy"] == 1.0 and np.allclose(res["iou"], 1.0) and res["mean_iou"] == 1.0
    wrong = ev.SegmentationMetrics(ncls)                         # everything predicted as class 0
    for i, it in enumerate(items):
        z = np.zeros_like(summed[i]); z[:, 0] = 1
        wrong.update(z, it[:, 6], it[:, 7])
    r = wrong.result()
    inner_gt = np.concatenate([it[it[:, 7] == 1, 6] for it in items])
    assert np.isclose(r["accuracy"], (inner_gt == 0).mean()) and np.isclose(r["iou"][0], (inner_gt == 0).mean())
    assert (r["iou"][1:] == 0).all()
    with pytest.raises(ValueError):
        ev.predict_blocks_with_overlap(padded, num_point, lambda x: np.zeros((3, num_point, ncls + 1)), ncls)


def test_shapenet_and_modelnet_records(pkg, tfr, tmp_path):
    rng = np.random.default_rng(41)
    shapes = [(rng.random((n, 3), dtype=np.float32), rng.integers(0, 50, n).astype(np.int32)) for n in (20, 35, 28)]
    p1 = str(tmp_path / "shapenet.tfrecord")
    tfr.write_records(p1, [tfr.make_example({"xyz_raw": x.tobytes(), "part_label": l.tobytes()}) for x, l in shapes])
    sn = pkg.io.shapenet_input
    batches = list(sn.input_fn([p1], batch_size=2, buffer_size=2, rng=np.random.default_rng(1)))      # train_shapenet.py:155-180
    assert [b.shape for b in batches] in ([(2, 35, 4), (1, 20, 4)], [(2, 35, 4), (1, 28, 4)], [(2, 28, 4), (1, 35, 4)])
    rows = np.concatenate([b[i][b[i, :, -1] >= 0] for b in batches for i in range(len(b))])
    want = np.concatenate([np.concatenate([x, l[:, None].astype(np.float32)], 1) for x, l in shapes])
    assert sorted(map(tuple, rows.tolist())) == sorted(map(tuple, want.tolist()))
    clouds = [(rng.random((64, 3), dtype=np.float32), int(rng.integers(0, 40))) for _ in range(5)]
    p2 = str(tmp_path / "modelnet.tfrecord")
    tfr.write_records(p2, [tfr.make_example({"xyz_raw": x.tobytes(), "label": np.array([l])}) for x, l in clouds])
    mn = pkg.io.modelnet_input
    got = list(mn.input_fn([p2], batch_size=2, buffer_size=10, rng=np.random.default_rng(2)))          # train_modelnet.py:118-138
    assert [x.shape for x, _ in got] == [(2, 64, 3), (2, 64, 3), (1, 64, 3)] and got[0][1].dtype == np.int32
    assert sorted(int(l) for _, ls in got for l in ls) == sorted(l for _, l in clouds)
    by_label = {l: x for x, l in clouds}
    for xs, ls in got:
        for x, l in zip(xs, ls):
            if sum(1 for _, l2 in clouds if l2 == int(l)) == 1:
                assert np.array_equal(x, by_label[int(l)])


def test_scene_merge_restates_the_matlab_script(pkg):
    """post-merging/s3dis_merge.m:37-81 written out literally (loops, 1-based -> 0-based) vs io/s3dis_merge.py"""
    mg = pkg.io.s3dis_merge
    rng = np.random.default_rng(5)
    P, C = 400, 13
    voxel_xyz = rng.random((P, 3))
    gt = rng.integers(0, C, P)
    blocks = []
    for k in range(5):                                           # overlapping blocks: random subsets of the scene
        n = int(rng.integers(120, 260))
        index = rng.choice(P, n, replace=False)
        inner = (rng.random(n) < 0.6).astype(np.int32)
        logits = (rng.standard_normal((n, C)) * rng.integers(1, 4)).astype(np.float32)     # summed over 1-3 draws
        blocks.append((logits, inner, index))
    # literal restatement
    want = np.zeros((P, C))
    for logits, inner, index in blocks:
        for row in range(len(index)):
            if inner[row] != 1:
                continue
            v = logits[row].astype(np.float64)
            v = v / np.sqrt(np.sum(v ** 2))
            v = np.exp(v) / np.sum(np.exp(v))
            want[index[row]] += v
    pred, label = mg.merge_scene(P, blocks, C)
    assert np.allclose(pred, want, rtol=1e-12, atol=0)
    assert (label == want.argmax(1)).all()
    assert np.allclose(mg.block_confidence(blocks[0][0]).sum(1), 1.0)
    assert (mg.block_confidence(np.zeros((2, C))) == 0).all()     # never-drawn rows contribute nothing (MATLAB: NaN)
    with pytest.raises(ValueError):
        mg.merge_scene(10, blocks, C)
    # nearest-voxel propagation to the full cloud + IoU totals
    full_xyz = voxel_xyz[rng.integers(0, P, 1500)] + rng.normal(0, 1e-4, (1500, 3))
    d = ((full_xyz[:, None, :] - voxel_xyz[None]) ** 2).sum(-1)
    assert (mg.propagate_to_full_cloud(voxel_xyz, label, full_xyz) == label[d.argmin(1)]).all()
    full_gt = gt[d.argmin(1)]
    full_pred = label[d.argmin(1)]
    m = mg.SceneIoU(C)
    m.update(full_pred[:700], full_gt[:700]); m.update(full_pred[700:], full_gt[700:])
    res = m.result()
    for c in range(C):
        inter, uni = ((full_pred == c) & (full_gt == c)).sum(), ((full_pred == c) | (full_gt == c)).sum()
        assert m.intersect[c] == inter and m.union[c] == uni and (uni == 0 or abs(res["iou"][c] - inter / uni) < 1e-12)
