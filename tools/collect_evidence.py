#!/usr/bin/env python3
"""Turns what tools/gpu_evidence*.sh brought back under gpurun_out/ev/ into the tracked
files under profiles/ (round 2): ncu summaries (tools/ncu_summary.py), traffic.json (DRAM
bytes of the dominant kernel per headline workload, with kernel name and library source hash),
the launch list of the bench command restricted to this library's kernels, the SASS summary
and the jsonl records.  Run here (no GPU needed) after the GPU calls:

    python tools/collect_evidence.py
"""
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EV = os.path.join(ROOT, "gpurun_out", "ev")
PROF = os.path.join(ROOT, "profiles")

NCU = {  # report -> (summary file, traffic.json key)
    "prof_bf16.ncu-rep": ("r02_ncu_lean2_bf16_headline.txt", "e2e_4k_bf16_b32"),
    "prof_f32.ncu-rep": ("r02_ncu_lean2_fp32_exact.txt", "e2e_4k_f32_b16"),
    "prof_u8.ncu-rep": ("r02_ncu_lean_u8_headline.txt", "e2e_4k_u8_b32"),
    "prof_bicubic.ncu-rep": ("r02_ncu_lean2_bicubic_bf16.txt", None),
    "prof_p2e.ncu-rep": ("r02_ncu_lean2_pers2equi_u8_bicubic.txt", "p2e_8k_u8_bicubic_b8"),
    "prof_nearest.ncu-rep": ("r02_ncu_lean2_nearest_bf16.txt", None),
}
COPY = {  # gpurun_out/ev file -> profiles file
    "pytest_gpu.log": "r02_pytest_gpu.log",
    "parity_rates.jsonl": "r02_parity_rates.jsonl",
    "configs.jsonl": "r02_configs.jsonl",
    "rotation_sweep.jsonl": "r02_rotation_sweep.jsonl",
    "validate.jsonl": "r02_validate_vs_reference_cuda.jsonl",
    "round2b.jsonl": "r02_round2b.jsonl",
}


def num(text, key):
    m = re.search(r"^" + re.escape(key) + r"\s+([0-9.]+)\s+(\S+)", text, re.M)
    if not m:
        return None
    v = float(m.group(1))
    unit = m.group(2).lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)


def main():
    sys.path.insert(0, ROOT)
    from equilib_b200 import build

    digest = build.source_hash()
    traffic_path = os.path.join(PROF, "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for rep, (name, key) in NCU.items():
        src = os.path.join(EV, rep)
        if not os.path.exists(src):
            print("missing", rep)
            continue
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), src],
                             capture_output=True, text=True).stdout
        open(os.path.join(PROF, name), "w").write(out)
        if key:
            rd, wr = num(out, "dram__bytes_read.sum"), num(out, "dram__bytes_write.sum")
            kern = re.search(r"^kernel: (.*)$", out, re.M)
            if rd is not None and wr is not None:
                traffic[key] = {"bytes": int(rd + wr), "kernel": kern.group(1) if kern else "?",
                                "source_hash": digest, "capture": "profiles/" + name}
        print("wrote", name)
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    # launch list of the bench command: this library's kernels only
    lc = os.path.join(EV, "launches.csv")
    if os.path.exists(lc):
        rows = [r for r in csv.reader(l for l in open(lc) if not l.startswith("=="))]
        keep = [r for r in rows if r and (r[0] == "ID" or any("remap" in c or "minmax" in c
                                                                for c in r))]
        with open(os.path.join(PROF, "r02_launches_lean2_remap_only.csv"), "w", newline="") as f:
            csv.writer(f).writerows(keep)
        print("launch list:", len(keep) - 1, "launches")
    for a, b in COPY.items():
        if os.path.exists(os.path.join(EV, a)):
            shutil.copyfile(os.path.join(EV, a), os.path.join(PROF, b))
            print("copied", b)
    lines = []
    for f in ("bench_default.json", "bench_reference.json", "bench_workloads.jsonl"):
        p = os.path.join(EV, f)
        if os.path.exists(p):
            lines += [l for l in open(p).read().splitlines() if l.startswith("{")]
    if lines:
        open(os.path.join(PROF, "r02_bench_lines.jsonl"), "w").write("\n".join(lines) + "\n")
        print("bench lines:", len(lines))
    sass = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_summary.py")],
                          capture_output=True, text=True).stdout
    open(os.path.join(PROF, "r02_sass_summary.txt"), "w").write(sass)
    print("sass summary:", len(sass.splitlines()), "lines; library source hash", digest)


if __name__ == "__main__":
    main()
