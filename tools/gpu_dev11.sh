#!/bin/bash
mkdir -p gpurun_out/dev11
O=gpurun_out/dev11
timeout 900 python -m pytest tests -m gpu -q -k "nearest or tie_guard" > $O/pytest.log 2>&1; tail -8 $O/pytest.log
python - <<'PY'
import os, math, sys, torch
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
    for no in ("", "1"):
        if no: os.environ["EQUILIB_B200_NO_LEAN2"] = "1"
        else: os.environ.pop("EQUILIB_B200_NO_LEAN2", None)
        ms = timed(lambda: equilib_b200.equi2equi(src, r, mode="nearest"))
        print(f"equi2equi nearest {dt} b{b} no_lean2={no or 0}: {ms:.3f} ms", flush=True)
        if b == 16 or dt == torch.uint8:
            ms = timed(lambda: equilib_b200.equi2cube(src[:16], r[:16], 1024, "dice", mode="nearest"))
            print(f"equi2cube dice nearest {dt} b16 no_lean2={no or 0}: {ms:.3f} ms", flush=True)
    os.environ.pop("EQUILIB_B200_NO_LEAN2", None)
    ry = [{"roll": 0.0, "pitch": 0.0, "yaw": 0.2 * i} for i in range(b)]
    ms = timed(lambda: equilib_b200.equi2equi(src, ry, mode="nearest"))
    ms_e = timed(lambda: equilib_b200.equi2equi(src, ry, mode="nearest", fast_coords=False))
    print(f"equi2equi nearest {dt} b{b} YAW ONLY: default {ms:.3f} ms, exact {ms_e:.3f} ms", flush=True)
    ms_e = timed(lambda: equilib_b200.equi2equi(src, r, mode="nearest", fast_coords=False))
    print(f"equi2equi nearest {dt} b{b} bench rotations, exact: {ms_e:.3f} ms", flush=True)
    del src
PY
