#!/bin/bash
mkdir -p gpurun_out/dev2
O=gpurun_out/dev2
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_cube or lean2_bicubic or staged_and_direct" > $O/pytest_new.log 2>&1
echo "pytest rc=$?" >> $O/pytest_new.log
tail -5 $O/pytest_new.log
python - > $O/rot_ab.txt 2>&1 <<'PY'
import os, math, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import equilib_b200
dev = torch.device("cuda", 0)
b = 32
src = torch.rand((b, 3, 2048, 4096), device=dev).to(torch.bfloat16)
rng = np.random.default_rng(0)
R = {
 "identity": [{"roll": 0.0, "pitch": 0.0, "yaw": 0.0}] * b,
 "bench": [{"roll": 0.01 * i, "pitch": 0.4 * math.sin(i), "yaw": 0.05 * i} for i in range(b)],
 "random": [{"roll": float(rng.uniform(-math.pi, math.pi)), "pitch": float(math.asin(rng.uniform(-1, 1))), "yaw": float(rng.uniform(-math.pi, math.pi))} for _ in range(b)],
}
def timed(fn, steps=5, warmup=2):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
for name, r in R.items():
    for fast in (None, True):
        for no in ("", "1"):
            if no: os.environ["EQUILIB_B200_NO_LEAN2"] = "1"
            else: os.environ.pop("EQUILIB_B200_NO_LEAN2", None)
            ms = timed(lambda: equilib_b200.equi2equi(src, r, mode="bicubic", clip_output=False, fast_coords=fast))
            print(f"{name:9s} fast={fast} no_lean2={no or 0}: {ms:.3f} ms", flush=True)
PY
cat $O/rot_ab.txt
MODE=bicubic B=8 N=3 python tools/prof_one.py > $O/prof_plain.log 2>&1 && MODE=bicubic B=8 ncu --set full --import-source on --clock-control none -k regex:remap_lean2 -s 2 -c 1 -o $O/prof_bicubic_bf16 -f python tools/prof_one.py > $O/ncu_bicubic.log 2>&1
tail -3 $O/ncu_bicubic.log
ls -la $O
