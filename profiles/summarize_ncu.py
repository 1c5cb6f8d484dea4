"""Turns an Nsight Compute report (.ncu-rep, captured on the GPU box with the command in its header) into
the tracked evidence under profiles/:  python profiles/summarize_ncu.py gpurun_out/X.ncu-rep r1_conv

  profiles/<tag>_ncu.md        one block of key metrics per captured kernel launch
  profiles/ncu_summary.json    {"kernels": {<key>: {"dram_bytes_per_launch": ...}}}  (read by bench.py -> roofline.traffic)
Runs on the CPU box (ncu -i needs no GPU).
"""
import csv
import json
import os
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("gpc__cycles_elapsed.avg.per_second", "SM clock"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps/cycle"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sectors.sum", "L2 sectors"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 global-load sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "L1 reduction sectors"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "L1 shared wavefronts"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long-scoreboard / issue"),
]


def short(name):
    n = name.replace("void ", "").replace("sph3d::", "")
    return n.split("(")[0]


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    here = os.path.dirname(os.path.abspath(__file__))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = ["# ncu summary `%s` (source: %s, `ncu --set full --clock-control none --import-source on`)\n" % (tag, os.path.basename(rep))]
    summary_path = os.path.join(here, "ncu_summary.json")
    summary = json.load(open(summary_path)) if os.path.exists(summary_path) else {"kernels": {}}
    for d in data:
        name = short(d[col["Kernel Name"]])
        out.append("## %s\n" % name)
        out.append("| metric | value |\n|---|---|")
        for m, label in METRICS:
            if m in col and d[col[m]] != "":
                out.append("| %s (`%s`) | %s %s |" % (label, m, d[col[m]], units[col[m]]))
        out.append("")
        try:
            def to_bytes(m):
                v, u = float(d[col[m]]), units[col[m]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            key = name.split("<")[0]
            summary["kernels"][key] = {
                "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
                "duration_ms_under_ncu": float(d[col["gpu__time_duration.sum"]]) * {"ms": 1, "us": 1e-3, "s": 1e3}.get(units[col["gpu__time_duration.sum"]].lower().replace("msecond", "ms").replace("usecond", "us").replace("second", "s"), 1),
                "kernel": name, "report": os.path.basename(rep), "tag": tag}
        except Exception as e:                                                   # keep the markdown even if a unit is new
            out.append("(summary json skipped: %r)\n" % (e,))
    # the backward op is several launches: aggregate them per op (bench.py roofline.traffic for "conv_bwd_op")
    BWD = ("transpose_edges_kernel", "scan_reduce_kernel", "scan_apply_kernel", "sort_segments_kernel", "scale_rows_kernel",
           "conv_bwd_t_kernel", "reduce_partials_kernel")
    ops = sum(1 for d in data if short(d[col["Kernel Name"]]).split("<")[0] == "conv_bwd_t_kernel")
    if ops:
        tot_b, tot_ms, n = 0.0, 0.0, 0
        for d in data:
            if short(d[col["Kernel Name"]]).split("<")[0] in BWD:
                ub = lambda m: float(d[col[m]]) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(units[col[m]].lower(), 1)
                tot_b += ub("dram__bytes_read.sum") + ub("dram__bytes_write.sum")
                tot_ms += float(d[col["gpu__time_duration.sum"]]) * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(
                    units[col["gpu__time_duration.sum"]].lower().replace("msecond", "ms").replace("usecond", "us").replace("nsecond", "ns").replace("second", "s"), 1)
                n += 1
        summary["kernels"]["conv_bwd_op"] = {"dram_bytes_per_launch": tot_b / ops, "duration_ms_under_ncu": tot_ms / ops,
                                             "kernel": "all %d kernels of one sph3d_depthwise_conv3d_grad call (transposed form)" % (n // ops),
                                             "report": os.path.basename(rep), "tag": tag}
        out.append("## conv backward op (sum over its %d kernels)\n\nDRAM bytes %.1f MB, device time %.3f ms (under ncu)\n" % (n // ops, tot_b / ops / 1e6, tot_ms / ops))
    open(os.path.join(here, "%s_ncu.md" % tag), "w").write("\n".join(out) + "\n")
    json.dump(summary, open(summary_path, "w"), indent=1, sort_keys=True)
    print("wrote", os.path.join(here, "%s_ncu.md" % tag), "and", summary_path)


if __name__ == "__main__":
    main()
