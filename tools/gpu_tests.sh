#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_rates.jsonl
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
