#!/usr/bin/env python3
"""Per-source-line totals from an .ncu-rep compiled with -lineinfo: executed warp-instructions,
stall samples and shared-memory wavefronts per (file, line), heaviest first.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source",
                          "sass,cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname = ""
    hdr = None
    agg = {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) == 2:
            continue
        if r and r[0] == "Line No":
            hdr = r
            i_ex = hdr.index("Instructions Executed")
            i_sm = hdr.index("# Samples")
            i_wf = hdr.index("L1 Wavefronts Shared")
            i_wi = hdr.index("L1 Wavefronts Shared Ideal")
            continue
        if hdr is None or not r or r[0] == "":
            continue
        key = (fname, int(r[0]), r[1].strip()[:90])
        a = agg.setdefault(key, [0, 0, 0, 0])
        def num(v):
            try:
                return int(v)
            except ValueError:
                return 0
        a[0] += num(r[i_ex]); a[1] += num(r[i_sm]); a[2] += num(r[i_wf]); a[3] += num(r[i_wi])
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[1] for a in agg.values()) or 1
    totw = sum(a[2] for a in agg.values()) or 1
    print(f"total warp-instr {tot}, samples {tots}, shared wavefronts {totw} "
          f"(ideal {sum(a[3] for a in agg.values())})")
    print(f"{'file:line':28s} {'inst%':>6s} {'smp%':>6s} {'wf%':>6s} {'wf/ideal':>8s}  source")
    for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        ratio = a[2] / a[3] if a[3] else 0
        print(f"{f + ':' + str(ln):28s} {a[0] / tot * 100:6.2f} {a[1] / tots * 100:6.2f} "
              f"{a[2] / totw * 100:6.2f} {ratio:8.2f}  {src}")


if __name__ == "__main__":
    main()
