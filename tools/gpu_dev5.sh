#!/bin/bash
mkdir -p gpurun_out/dev5
O=gpurun_out/dev5
python bench.py --workload p2e_8k_u8_bicubic_b8 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/plain.json 2> $O/plain.err
ncu --set full --import-source on --clock-control none -k regex:remap_lean2 -s 3 -c 1 -o $O/prof_p2e -f python bench.py --workload p2e_8k_u8_bicubic_b8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu.log 2>&1
tail -2 $O/ncu.log; ls -la $O
