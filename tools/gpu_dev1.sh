#!/bin/bash
# dev call: new parity tests first, then A/B timings
mkdir -p gpurun_out
O=gpurun_out/dev1
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "lean2_bicubic or fused_cube or staged_and_direct or bicubic or 16bit" > $O/pytest_new.log 2>&1
echo "pytest rc=$?" >> $O/pytest_new.log
tail -30 $O/pytest_new.log
timeout 600 python tools/bench_round2b.py > $O/round2b.jsonl 2> $O/round2b.err
echo "bench rc=$?"
tail -5 $O/round2b.err
cat $O/round2b.jsonl | cut -c1-260
