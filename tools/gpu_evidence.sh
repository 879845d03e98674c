#!/bin/bash
# One gpurun call: everything profiles/ needs, from the code as it is now.
mkdir -p gpurun_out/ev
O=gpurun_out/ev
rm -f gpurun_out/parity_rates.jsonl
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
cp gpurun_out/parity_rates.jsonl $O/parity_rates.jsonl 2>/dev/null
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_reference.json 2>> $O/bench_default.err
for w in e2e_4k_f32_b16 e2e_4k_u8_b32 e2p_8k_u8_bicubic_b8 p2e_8k_u8_bicubic_b8 e2p_multiview_f16_v256; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline >> $O/bench_workloads.jsonl 2>> $O/bench_default.err
done
python bench.py --exact-coords --steps 20 --warmup 3 --no-cpu-baseline --no-e2e >> $O/bench_workloads.jsonl 2>> $O/bench_default.err
python tools/bench_configs.py > $O/configs.jsonl 2> $O/configs.err
python tools/rotation_sweep.py > $O/rotation_sweep.jsonl 2> $O/rotation_sweep.err
python tools/bench_round2b.py > $O/round2b.jsonl 2> $O/round2b.err
python tools/validate_vs_reference_cuda.py --ref baseline/_ref > $O/validate.jsonl 2> $O/validate.err
# launch list of the bench command (share of the step per kernel), then full captures
tail -3 $O/pytest_gpu.log
