#!/bin/bash
# 8-GPU box: the host-copy ceiling at 1/2/4/8 processes, then the bench at N = 8 and 1.
mkdir -p gpurun_out
python tools/host_copy_ceiling.py --procs 1 2 4 8 --mb 1024 --seconds 1.5 --out gpurun_out/host_copy_ceiling.json > gpurun_out/host_copy_ceiling.log 2>&1
nproc > gpurun_out/topology.txt; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" >> gpurun_out/topology.txt; nvidia-smi topo -m >> gpurun_out/topology.txt 2>&1
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 30 --warmup 5 >> gpurun_out/bench_scale.log 2>> gpurun_out/bench_scale.err
done
python bench.py --steps 30 --warmup 5 --no-cpu-baseline >> gpurun_out/bench_scale.log 2>> gpurun_out/bench_scale.err
for n in 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 --warmup 3 --workload e2p_multiview_f16_v256 >> gpurun_out/bench_scale_cfg.log 2>> gpurun_out/bench_scale.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --steps 20 --warmup 3 --workload e2p_8k_u8_bicubic_b8 >> gpurun_out/bench_scale_cfg.log 2>> gpurun_out/bench_scale.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $n --steps 20 --warmup 3 --workload p2e_8k_u8_bicubic_b8 >> gpurun_out/bench_scale_cfg.log 2>> gpurun_out/bench_scale.err
done
tail -c 600 gpurun_out/host_copy_ceiling.log
