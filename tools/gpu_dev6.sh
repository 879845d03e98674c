#!/bin/bash
mkdir -p gpurun_out/dev6
O=gpurun_out/dev6
timeout 900 python -m pytest tests -m gpu -q -k "pers2equi or p2e or staged_and_direct or golden_case" > $O/pytest.log 2>&1; tail -8 $O/pytest.log
python bench.py --workload p2e_8k_u8_bicubic_b8 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/p2e.json 2> $O/p2e.err
python -c "
import json;d=json.load(open('$O/p2e.json'));print('p2e config4', d['ms_per_step'], d['value'], d['roofline']['frac'])"
EQUILIB_B200_NO_LEAN2=1 python bench.py --workload p2e_8k_u8_bicubic_b8 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/p2e_old.json 2>> $O/p2e.err
python -c "
import json;d=json.load(open('$O/p2e_old.json'));print('p2e config4 direct kernel', d['ms_per_step'], d['value'], d['roofline']['frac'])"
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
r = [{"roll": 0.01 * i, "pitch": 0.4 * math.sin(i), "yaw": 0.05 * i} for i in range(8)]
for dt in (torch.uint8, torch.float32, torch.bfloat16):
    pers = (torch.rand((8, 3, 1080, 1920), device=dev) * (255 if dt == torch.uint8 else 1)).to(dt)
    for fov in (90.0, 140.0):
        for no in ("", "1"):
            if no: os.environ["EQUILIB_B200_NO_LEAN2"] = "1"
            else: os.environ.pop("EQUILIB_B200_NO_LEAN2", None)
            ms = timed(lambda: equilib_b200.pers2equi(pers, r, 4096, 8192, fov, mode="bicubic"))
            print(f"pers2equi bicubic {dt} fov {fov} no_lean2={no or 0}: {ms:.3f} ms", flush=True)
        os.environ.pop("EQUILIB_B200_NO_LEAN2", None)
        for mode in ("nearest", "bilinear"):
            ms = timed(lambda: equilib_b200.pers2equi(pers, r, 4096, 8192, fov, mode=mode))
            print(f"pers2equi {mode} {dt} fov {fov}: {ms:.3f} ms", flush=True)
PY
