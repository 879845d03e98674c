#!/bin/bash
mkdir -p gpurun_out/dev4
O=gpurun_out/dev4
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "lean2_bicubic or staged_and_direct or fused_cube" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
python bench.py --workload e2e_4k_f32_b16 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_f32.json 2> $O/bench.err
python -c "
import json;d=json.load(open('$O/bench_f32.json'));print('fp32 headline', d['ms_per_step'], d['roofline']['frac'])"
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_bf16.json 2>> $O/bench.err
python -c "
import json;d=json.load(open('$O/bench_bf16.json'));print('bf16 headline', d['ms_per_step'], d['roofline']['frac'])"
python bench.py --exact-coords --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_bf16_exact.json 2>> $O/bench.err
python -c "
import json;d=json.load(open('$O/bench_bf16_exact.json'));print('bf16 exact', d['ms_per_step'], d['roofline']['frac'])"
python - <<'PY'
import os, math, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import equilib_b200
dev = torch.device("cuda", 0)
def timed(fn, steps=5, warmup=2):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
for dt, b in ((torch.bfloat16, 32), (torch.float32, 16), (torch.uint8, 32)):
    src = (torch.rand((b, 3, 2048, 4096), device=dev) * (255 if dt == torch.uint8 else 1)).to(dt)
    r = [{"roll": 0.01 * i, "pitch": 0.4 * math.sin(i), "yaw": 0.05 * i} for i in range(b)]
    for fast in (None, False, True):
        ms = timed(lambda: equilib_b200.equi2equi(src, r, mode="bicubic", clip_output=False, fast_coords=fast))
        print(f"bicubic {dt} fast={fast}: {ms:.3f} ms", flush=True)
    del src
PY
