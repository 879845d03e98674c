#!/bin/bash
# the bench lines only (after traffic.json has been refreshed for the current library)
mkdir -p gpurun_out/ev
O=gpurun_out/ev
rm -f $O/bench_workloads.jsonl
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_reference.json 2>> $O/bench_default.err
for w in e2e_4k_f32_b16 e2e_4k_u8_b32 e2p_8k_u8_bicubic_b8 p2e_8k_u8_bicubic_b8 e2p_multiview_f16_v256; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline >> $O/bench_workloads.jsonl 2>> $O/bench_default.err
done
python bench.py --exact-coords --steps 20 --warmup 3 --no-cpu-baseline --no-e2e >> $O/bench_workloads.jsonl 2>> $O/bench_default.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
tail -c 400 $O/bench_default.json
