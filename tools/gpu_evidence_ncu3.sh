#!/bin/bash
# ncu captures, call 3 of 4: the uint8 headline kernel (round-1 lean kernel)
mkdir -p gpurun_out/ev
O=gpurun_out/ev
DTYPE=u8 N=3 python tools/prof_one.py >> $O/prof_plain.log 2>&1 && DTYPE=u8 ncu --set full --import-source on --clock-control none -k regex:remap_lean -s 2 -c 1 -o $O/prof_u8 -f python tools/prof_one.py > $O/ncu_u8.log 2>&1
ls -la $O
