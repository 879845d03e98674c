#!/bin/bash
mkdir -p gpurun_out/dev10
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "lean2_pers2equi" > gpurun_out/dev10/pytest.log 2>&1; tail -15 gpurun_out/dev10/pytest.log
