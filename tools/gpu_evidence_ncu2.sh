#!/bin/bash
# ncu --set full captures (at most two reports per gpurun call: 64 MiB come back)
mkdir -p gpurun_out/ev
O=gpurun_out/ev
DTYPE=u8 N=3 python tools/prof_one.py >> $O/prof_plain.log 2>&1 && DTYPE=u8 ncu --set full --import-source on --clock-control none -k regex:remap_lean -s 2 -c 1 -o $O/prof_u8 -f python tools/prof_one.py > $O/ncu_u8.log 2>&1
MODE=bicubic N=3 python tools/prof_one.py >> $O/prof_plain.log 2>&1 && MODE=bicubic ncu --set full --import-source on --clock-control none -k regex:remap_lean2 -s 2 -c 1 -o $O/prof_bicubic -f python tools/prof_one.py > $O/ncu_bicubic.log 2>&1
ls -la $O
