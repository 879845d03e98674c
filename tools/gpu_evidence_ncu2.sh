#!/bin/bash
# ncu captures, call 2 of 4: the fp32 (exact chain) headline kernel
mkdir -p gpurun_out/ev
O=gpurun_out/ev
DTYPE=f32 B=16 N=3 python tools/prof_one.py >> $O/prof_plain.log 2>&1 && DTYPE=f32 B=16 ncu --set full --import-source on --clock-control none -k regex:remap_lean2 -s 2 -c 1 -o $O/prof_f32 -f python tools/prof_one.py > $O/ncu_f32.log 2>&1
ls -la $O
