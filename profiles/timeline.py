"""GPU timeline of one training step (torch.profiler / CUPTI; no nsys in the image): per-kernel start/duration/stream,
the busy-time union, and the largest idle gaps with the kernels on either side.  Not a timing source (profiler
overhead inflates host-bound gaps) -- it shows WHERE the step idles.

    python profiles/timeline.py --model s3dis [--graph]      -> gpurun_out/timeline_<model>[_graph].json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "profiles")]

import torch
from torch.profiler import ProfilerActivity, profile

import bench_encoder as be


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="modelnet")
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--B", type=int, default=0, help="clouds per step (default: the config's batch); B=1 = one rank's share of S3DIS at 8 GPUs")
    a = ap.parse_args()
    B, N = be.DEFAULT_SHAPE[a.model]
    B = a.B or B
    step, cfg = be.make_step(B, N, model=a.model)
    run_step = be.S.utils.graph_step.GraphedStep(step, be.s3g_util.trainable_variables, warmup=3) if a.graph else step
    for _ in range(3):
        run_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run_step()
    e1.record()
    torch.cuda.synchronize()
    rec = {"ms_per_step": e0.elapsed_time(e1) / 5}
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        run_step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    last = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "stream", -1)) for e in evs), key=lambda t: t[0])
    t0, t1 = last[0][0], max(k[1] for k in last)
    busy, cur_end = 0.0, t0
    gaps = []
    prev = None
    for s, e, name, st in last:
        if s > cur_end:
            gaps.append((s - cur_end, prev, name))
            busy += e - s
            cur_end = e
        else:
            if e > cur_end:
                busy += e - cur_end
                cur_end = e
        prev = name if e >= cur_end else prev
    gaps.sort(key=lambda g: -g[0])
    by_name = {}
    for s, e, name, st in last:
        d = by_name.setdefault(name[:90], [0, 0.0])
        d[0] += 1
        d[1] += e - s
    out = {"model": a.model, "graph": a.graph, "ms_per_step_unprofiled_driver": rec["ms_per_step"], "kernels_in_step": len(last),
           "span_us": t1 - t0, "busy_union_us": busy, "idle_us": (t1 - t0) - busy, "sum_kernel_us": sum(k[1] - k[0] for k in last),
           "streams": sorted(set(k[3] for k in last)),
           "gap_histogram_us": {"<2": sum(1 for g in gaps if g[0] < 2), "2-5": sum(1 for g in gaps if 2 <= g[0] < 5),
                                "5-10": sum(1 for g in gaps if 5 <= g[0] < 10), "10-50": sum(1 for g in gaps if 10 <= g[0] < 50),
                                ">=50": sum(1 for g in gaps if g[0] >= 50)},
           "idle_in_gaps_ge_10us": sum(g[0] for g in gaps if g[0] >= 10),
           "top_gaps": [{"us": g[0], "after": (g[1] or "")[:80], "before": g[2][:80]} for g in gaps[:25]],
           "top_kernels": sorted(([n, c, round(t, 1)] for n, (c, t) in by_name.items()), key=lambda r: -r[2])[:25]}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out["B"] = B
    fn = os.path.join(ROOT, "gpurun_out", "timeline_%s%s%s.json" % (a.model, "_graph" if a.graph else "", "_B%d" % B if a.B else ""))
    json.dump(out, open(fn, "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("model", "graph", "kernels_in_step", "span_us", "busy_union_us", "idle_us", "sum_kernel_us",
                                          "gap_histogram_us", "idle_in_gaps_ge_10us")}))


if __name__ == "__main__":
    main()
