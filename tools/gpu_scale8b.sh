#!/bin/bash
# 8-GPU box, reduced: the bench at N = 8 (headline + config 4 both directions + config 5)
mkdir -p gpurun_out/scale8b
O=gpurun_out/scale8b
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 >> $O/bench_scale.jsonl 2>> $O/bench_scale.err
p=29512
for w in e2p_multiview_f16_v256 e2p_8k_u8_bicubic_b8 p2e_8k_u8_bicubic_b8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $p bench.py --gpus 8 --steps 20 --warmup 3 --workload $w >> $O/bench_scale.jsonl 2>> $O/bench_scale.err
p=$((p+1))
done
tail -3 $O/bench_scale.err
python - <<'PY'
import json
for l in open("gpurun_out/scale8b/bench_scale.jsonl"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["config"]["workload"][:60], d["n_gpus"], round(d["ms_per_step"], 4), round(d["value"], 1), d.get("e2e", {}).get("value"))
PY
