#!/bin/bash
# One gpurun call: GPU tests, parity rates, the bench lines.  Outputs under gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/parity_rates.jsonl
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
tail -c 1500 gpurun_out/bench_default.log
for w in e2e_4k_f32_b16 e2e_4k_u8_b32 e2p_8k_u8_bicubic_b8 p2e_8k_u8_bicubic_b8 e2p_multiview_f16_v256; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline >> gpurun_out/bench_workloads.log 2>> gpurun_out/bench_workloads.err
done
python bench.py --exact-coords --steps 20 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/bench_workloads.log 2>> gpurun_out/bench_workloads.err
