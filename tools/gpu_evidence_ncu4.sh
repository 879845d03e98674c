#!/bin/bash
# ncu captures, call 4 of 4: the strip-node kernel on bicubic (bf16 headline shape) and on
# pers2equi (config 4)
mkdir -p gpurun_out/ev
O=gpurun_out/ev
MODE=bicubic N=3 python tools/prof_one.py >> $O/prof_plain.log 2>&1 && MODE=bicubic ncu --set full --import-source on --clock-control none -k regex:remap_lean2 -s 2 -c 1 -o $O/prof_bicubic -f python tools/prof_one.py > $O/ncu_bicubic.log 2>&1
python bench.py --workload p2e_8k_u8_bicubic_b8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/p2e_plain.json 2>&1 && ncu --set full --import-source on --clock-control none -k regex:remap_lean2 -s 3 -c 1 -o $O/prof_p2e -f python bench.py --workload p2e_8k_u8_bicubic_b8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_p2e.log 2>&1
ls -la $O
