#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on a box without a GPU): key roofline metrics of the
profiled kernel, its SASS opcode mix and warp-stall reasons.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max",
]


def stamp():
    """Identity of the code the capture was taken from: the library's source hash (what
    bench.py prints as roofline.library_source_hash) and the git commit, when available."""
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    try:
        from equilib_b200 import build

        h = build.source_hash()
    except Exception:
        h = "?"
    try:
        g = subprocess.run(["git", "-C", root, "rev-parse", "--short", "HEAD"], capture_output=True,
                           text=True).stdout.strip()
    except Exception:
        g = "?"
    return f"library source hash {h} (tree at summary time), git {g or '?'}"


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True,
                         text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep = sys.argv[1]
    rows = ncu(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print("report:", rep)
    print("kernel:", d.get("Kernel Name", ("?",))[0])
    print("stamp:", stamp())
    for k in KEYS:
        if k in d:
            print(f"{k:72s} {d[k][0]:>18s} {d[k][1]}")
    rd = float(d["dram__bytes_read.sum"][0]) if "dram__bytes_read.sum" in d else 0
    wr = float(d["dram__bytes_write.sum"][0]) if "dram__bytes_write.sum" in d else 0
    print(f"dram traffic (read+write, units as above): {rd + wr:.6f}")

    src = ncu(rep, "source")
    hdr, data = src[1], src[2:]
    i_s, i_e, i_n = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    tot = sum(int(r[i_e]) for r in data)
    tot_s = sum(int(r[i_n]) for r in data) or 1
    ops, samp = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_s])
        op = m.group(2).split(".")[0] if m else "?"
        ops[op] += int(r[i_e])
        samp[op] += int(r[i_n])
    print(f"\nstatic SASS instructions {len(data)}, executed warp-instructions {tot}")
    print("opcode        exec%   stall-samples%")
    for op, c in ops.most_common(24):
        print(f"{op:12s} {c / tot * 100:6.1f} {samp[op] / tot_s * 100:8.1f}")
    cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.Counter()
    for r in data:
        for i in cols:
            agg[hdr[i]] += int(r[i])
    s = sum(agg.values()) or 1
    print("\nwarp stall reasons (all samples):")
    for k, v in agg.most_common(8):
        print(f"  {k:28s} {v / s * 100:5.1f} %")


if __name__ == "__main__":
    main()
