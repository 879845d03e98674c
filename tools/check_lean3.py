#!/usr/bin/env python3
"""A/B of the 32 x 32-tile kernel (remap_lean3.cuh) against the strip-node kernel
(remap_lean2.cuh, EQUILIB_B200_NO_LEAN3=1) and the exact-coordinate path on the headline
workload: value agreement and time, one JSON line per case."""
import json
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import equilib_b200  # noqa: E402

DEV = torch.device("cuda", 0)


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def with_env(name, val, fn):
    old = os.environ.get(name)
    os.environ[name] = val
    try:
        return fn()
    finally:
        if old is None:
            del os.environ[name]
        else:
            os.environ[name] = old


def band(b, c, h, w):
    x = torch.arange(w, device=DEV, dtype=torch.float32)[None, None, None, :]
    y = torch.arange(h, device=DEV, dtype=torch.float32)[None, None, :, None]
    ph = torch.arange(c, device=DEV, dtype=torch.float32)[None, :, None, None]
    bi = torch.arange(b, device=DEV, dtype=torch.float32)[:, None, None, None]
    return 0.5 + 0.5 * torch.sin(2 * math.pi * x / 512 + ph + 0.3 * bi) * torch.cos(2 * math.pi * y / 384)


def main():
    b = int(os.environ.get("B", "32"))
    H, W = int(os.environ.get("H", "2048")), int(os.environ.get("W", "4096"))
    rng = np.random.default_rng(0)
    fam = {
        "bench pattern": [{"roll": 0.01 * i, "pitch": 0.4 * math.sin(i), "yaw": 0.05 * i}
                          for i in range(b)],
        "uniform random": [{"roll": float(rng.uniform(-math.pi, math.pi)),
                            "pitch": float(math.asin(rng.uniform(-1, 1))),
                            "yaw": float(rng.uniform(-math.pi, math.pi))} for _ in range(b)],
        "identity": [{"roll": 0.0, "pitch": 0.0, "yaw": 0.0}] * b,
        "pitch 90 deg": [{"roll": 0.0, "pitch": math.pi / 2, "yaw": 0.2 * i} for i in range(b)],
    }
    dtypes = {"bf16": torch.bfloat16, "u8": torch.uint8, "f16": torch.float16, "f32": torch.float32}
    for dn in os.environ.get("DTYPES", "bf16,u8,f32").split(","):
        dt = dtypes[dn]
        bb = b // 2 if dn == "f32" else b
        img = band(bb, 3, H, W)
        src = (img * 255).to(torch.uint8) if dn == "u8" else img.to(dt)
        del img
        for name, rots in fam.items():
            rots = rots[:bb]
            new = equilib_b200.equi2equi(src, rots)
            old = with_env("EQUILIB_B200_NO_LEAN3", "1", lambda: equilib_b200.equi2equi(src, rots))
            exact = equilib_b200.equi2equi(src, rots, fast_coords=False)
            exact_old = with_env("EQUILIB_B200_NO_LEAN3", "1",
                                 lambda: equilib_b200.equi2equi(src, rots, fast_coords=False))
            torch.cuda.synchronize()

            def cmp(a, r):
                d = (a.float() - r.float()).abs()
                return {"same": round(float((d == 0).float().mean()), 6), "max": float(d.max())}

            rec = {"dtype": dn, "rotations": name, "frames": bb,
                   "new_vs_exact": cmp(new, exact), "old_vs_exact": cmp(old, exact),
                   "exact_new_vs_exact_old": cmp(exact, exact_old)}
            del new, old, exact, exact_old
            rec["ms_new"] = round(timed(lambda: equilib_b200.equi2equi(src, rots)), 4)
            rec["ms_old"] = round(with_env("EQUILIB_B200_NO_LEAN3", "1", lambda: timed(
                lambda: equilib_b200.equi2equi(src, rots))), 4)
            rec["ms_new_exact"] = round(timed(
                lambda: equilib_b200.equi2equi(src, rots, fast_coords=False), steps=5), 4)
            rec["ms_old_exact"] = round(with_env("EQUILIB_B200_NO_LEAN3", "1", lambda: timed(
                lambda: equilib_b200.equi2equi(src, rots, fast_coords=False), steps=5)), 4)
            print(json.dumps(rec), flush=True)
        del src
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
