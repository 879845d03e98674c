#!/bin/bash
# ncu captures, call 1 of 4 (64 MiB come back per gpurun call; a --set full report with sources
# is ~31 MB): the launch list of the bench command and the bf16 headline kernel
mkdir -p gpurun_out/ev
O=gpurun_out/ev
ncu --metrics gpu__time_duration.sum --clock-control none -c 460 --csv --log-file $O/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_launches.log 2>&1
N=3 python tools/prof_one.py > $O/prof_plain.log 2>&1 && ncu --set full --import-source on --clock-control none -k regex:remap_lean2 -s 2 -c 1 -o $O/prof_bf16 -f python tools/prof_one.py > $O/ncu_bf16.log 2>&1
ls -la $O
