#!/usr/bin/env python3
"""SASS evidence for the shipped library: per kernel, the static count of the instructions
that show what hardware path it uses -- UTMALDG (TMA box loads), SYNCS (mbarrier), FFMA2 /
FMUL2 (packed fp32 pairs), LDS / LDG / STG, PRMT, MUFU -- from `cuobjdump -sass`.

    python tools/sass_summary.py [equilib_b200/libequilib_b200.so] > profiles/rNN_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTMALDG", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "LDS", "LDG", "STG", "STS", "PRMT",
       "MUFU", "TEX", "TLD4", "ATOMS", "REDUX", "CREDUX", "BAR"]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "equilib_b200", "libequilib_b200.so")
    sys.path.insert(0, ROOT)
    from equilib_b200 import build

    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    arch = set()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur["total"] += 1
            cur[m.group(1)] += 1
    names = subprocess.run(["cu++filt"], input="\n".join(kernels), capture_output=True,
                           text=True).stdout.splitlines()
    print(f"library: {os.path.relpath(so, ROOT)}  source hash (eqb_source_hash): {build.built_hash()}  "
          f"arch: {', '.join(sorted(arch))}")
    print(f"kernels: {len(kernels)}; totals: " + ", ".join(
        f"{op} {sum(c[op] for c in kernels.values())}" for op in OPS[:5]))
    print(f"{'total':>6s} " + " ".join(f"{op:>7s}" for op in OPS) + "  kernel")
    rows = sorted(zip(names, kernels.values()), key=lambda kv: kv[0])
    # one line per kernel family (template arguments collapsed) with the maximum over its
    # instantiations, then the instantiations of the headline kernels in full
    fam = collections.OrderedDict()
    for n, c in rows:
        f = re.sub(r"<.*", "", n.replace("void ", ""))
        e = fam.setdefault(f, [0, collections.Counter()])
        e[0] += 1
        for k, v in c.items():
            e[1][k] = max(e[1][k], v)
    for f, (cnt, c) in fam.items():
        print(f"{c['total']:6d} " + " ".join(f"{c[op]:7d}" for op in OPS) + f"  {f} (max over {cnt} instantiations)")
    print("\nheadline instantiations:")
    for n, c in rows:
        if re.search(r"remap_lean2_kernel<\(int\)0, (__nv_bfloat16, \(int\)3, \(bool\)0|float, \(int\)3, \(bool\)1)|"
                     r"remap_lean_kernel<\(int\)0, unsigned char, \(int\)0, \(int\)3>|"
                     r"remap_staged_kernel<\(int\)0, \(int\)2, __nv_bfloat16", n):
            print(f"{c['total']:6d} " + " ".join(f"{c[op]:7d}" for op in OPS) + f"  {n.replace('void ', '')[:90]}")


if __name__ == "__main__":
    main()
