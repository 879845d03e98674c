#!/bin/bash
# ncu capture of the strip-node kernel on nearest (bf16 headline shape: interpolated coordinates,
# tie guard, deferred pixels)
mkdir -p gpurun_out/ev
O=gpurun_out/ev
MODE=nearest N=3 python tools/prof_one.py >> $O/prof_plain.log 2>&1 && MODE=nearest ncu --set full --import-source on --clock-control none -k regex:remap_lean2 -s 2 -c 1 -o $O/prof_nearest -f python tools/prof_one.py > $O/ncu_nearest.log 2>&1
ls -la $O
