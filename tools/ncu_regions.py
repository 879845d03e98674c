#!/usr/bin/env python3
"""Executed instructions of a kernel by region of its body (source line ranges of one file),
from an .ncu-rep compiled with -lineinfo.  Every SASS instruction is counted once, under the
line of `file` in its inline stack that lies inside the kernel body (>= body_start).

    python tools/ncu_regions.py rep.ncu-rep remap_lean2.cuh 241 "[(lo,hi,'name'),...]" pixels
"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep, file, body, reg, px = sys.argv[1], sys.argv[2], int(sys.argv[3]), eval(sys.argv[4]), float(sys.argv[5])
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source",
                          "sass,cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname, cur, hdr = "", None, None
    where = collections.defaultdict(list)
    info = {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) == 2:
            continue
        if r and r[0] == "Line No":
            hdr = r
            i_ex, i_sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
            i_wf = hdr.index("L1 Wavefronts Shared")
            continue
        if hdr is None or not r:
            continue
        if r[0] != "":
            cur = (fname, int(r[0]))
            continue
        a = r[2]
        where[a].append(cur)
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3])
        num = lambda v: int(v) if v.isdigit() else 0
        info[a] = (m.group(2).split(".")[0] if m else "?", num(r[i_ex]), num(r[i_sm]), num(r[i_wf]))
    agg = collections.Counter()
    smp = collections.Counter()
    wf = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    for a, locs in where.items():
        body_lines = [ln for f, ln in locs if f == file and ln >= body]
        name = "other"
        if body_lines:
            ln = max(body_lines)
            name = next((n for lo, hi, n in reg if lo <= ln <= hi), f"line {ln}")
        op, ex, sm, w = info[a]
        agg[name] += ex; smp[name] += sm; wf[name] += w
        ops[name][op] += ex
    tot, tots, totw = sum(agg.values()), sum(smp.values()) or 1, sum(wf.values()) or 1
    print(f"warp-instructions {tot} = {tot * 32 / px:.1f} thread-instr/px; shared wavefronts "
          f"{totw} = {totw * 128 / px:.1f} per 128 px")
    for k, v in agg.most_common():
        print(f"{k:14s} {v / tot * 100:5.1f}% {v * 32 / px:6.1f}/px  smp {smp[k] / tots * 100:5.1f}%  "
              f"wf/128px {wf[k] * 128 / px:5.1f} | "
              + ", ".join(f"{o} {c * 32 / px:.1f}" for o, c in ops[k].most_common(9)))


if __name__ == "__main__":
    main()
