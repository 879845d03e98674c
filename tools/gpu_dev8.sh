#!/bin/bash
mkdir -p gpurun_out/dev8
O=gpurun_out/dev8
for w in "" "--workload e2e_4k_f32_b16" "--workload e2p_multiview_f16_v256" "--workload e2p_8k_u8_bicubic_b8"; do
python bench.py $w --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $O/b.json 2>> $O/bench.err
python -c "
import json;d=json.load(open('$O/b.json'));print('$w', d['ms_per_step'], d['roofline']['frac'])"
done
