#!/bin/bash
mkdir -p gpurun_out/dev3
O=gpurun_out/dev3
rm -f gpurun_out/parity_rates.jsonl
echo "== default (exact) ==" > $O/rates.txt
timeout 600 python -m pytest tests -m gpu -q -k "bicubic" > $O/pytest_exact.log 2>&1; tail -3 $O/pytest_exact.log >> $O/rates.txt
cat gpurun_out/parity_rates.jsonl >> $O/rates.txt; rm -f gpurun_out/parity_rates.jsonl
echo "== EQUILIB_B200_FAST_COORDS=1 ==" >> $O/rates.txt
EQUILIB_B200_FAST_COORDS=1 timeout 600 python -m pytest tests -m gpu -q -k "bicubic" > $O/pytest_fast.log 2>&1; tail -15 $O/pytest_fast.log >> $O/rates.txt
cat gpurun_out/parity_rates.jsonl >> $O/rates.txt
cat $O/rates.txt
