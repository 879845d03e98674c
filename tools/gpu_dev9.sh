#!/bin/bash
mkdir -p gpurun_out/dev9
O=gpurun_out/dev9
timeout 900 python -m pytest tests -m gpu -q -k "cube2equi or c2e or staged_and_direct or golden_case or config3 or fused_cube or formats" > $O/pytest.log 2>&1; tail -8 $O/pytest.log
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
r = [{"roll": 0.01 * i, "pitch": 0.4 * math.sin(i), "yaw": 0.05 * i} for i in range(16)]
for dt in (torch.float32, torch.bfloat16, torch.uint8):
    equi = (torch.rand((16, 3, 2048, 4096), device=dev) * (255 if dt == torch.uint8 else 1)).to(dt)
    cube = equilib_b200.equi2cube(equi, r, 1024, "horizon")
    del equi
    for mode in ("bilinear", "bicubic"):
        for no in ("", "1"):
            if no: os.environ["EQUILIB_B200_NO_LEAN2"] = "1"
            else: os.environ.pop("EQUILIB_B200_NO_LEAN2", None)
            ms = timed(lambda: equilib_b200.cube2equi(cube, "horizon", 2048, 4096, mode=mode, clip_output=False))
            print(f"cube2equi horizon {mode} {dt} no_lean2={no or 0}: {ms:.3f} ms", flush=True)
    os.environ.pop("EQUILIB_B200_NO_LEAN2", None)
    if dt == torch.uint8:
        os.environ["EQUILIB_B200_FORCE_LEAN2"] = "1"
        ms = timed(lambda: equilib_b200.cube2equi(cube, "horizon", 2048, 4096, mode="bilinear"))
        print(f"cube2equi horizon bilinear uint8 FORCE_LEAN2: {ms:.3f} ms", flush=True)
        os.environ.pop("EQUILIB_B200_FORCE_LEAN2", None)
    del cube
PY
