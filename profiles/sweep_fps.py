"""FPS round latency: single-CTA kernel vs thread-block clusters with the sequence-tagged DSMEM handshake vs the same
clusters with a cluster barrier per round.   python profiles/sweep_fps.py [--out gpurun_out/r2_fps.json]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch

import sph3d_gcn_b200 as S


def env(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    S._lib.reload_tunables()


def timeit(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_fps.json"))
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = []
    g = torch.Generator().manual_seed(3)
    for (B, N, S_) in ((4, 65536, 16384), (1, 65536, 16384), (2, 20000, 5000), (32, 10000, 2500), (8, 8192, 2048), (1, 8192, 2048),
                       (8, 2048, 768), (16, 2048, 1024)):
        xyz = torch.rand(B, N, 3, generator=g).to(dev)
        base = None
        for name, e in (("default", {}), ("handshake_all_warps_poll", dict(SPH3D_FPS_HANDSHAKE=2)), ("cluster_barrier", dict(SPH3D_FPS_HANDSHAKE=0)),
                        ("clusters_from_1025", dict(SPH3D_FPS_CLUSTER_MIN_N=1025)),
                        ("clusters_from_4097", dict(SPH3D_FPS_CLUSTER_MIN_N=4097))):
            env(SPH3D_FPS_HANDSHAKE=None, SPH3D_FPS_CLUSTER_MIN_N=None)
            env(**e)
            out = S.tf_sample.farthest_point_sample(S_, xyz)
            if base is None:
                base = out.clone()
            same = bool(torch.equal(out, base))
            ms = timeit(lambda: S.tf_sample.farthest_point_sample(S_, xyz))
            rec = {"B": B, "N": N, "S": S_, "variant": name, "ms": round(ms, 4), "us_per_round": round(ms * 1e3 / S_, 4), "same_picks": same}
            rows.append(rec)
            print(json.dumps(rec), flush=True)
    env(SPH3D_FPS_HANDSHAKE=None, SPH3D_FPS_CLUSTER_MIN_N=None)
    json.dump({"device": torch.cuda.get_device_name(0), "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
