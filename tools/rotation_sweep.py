#!/usr/bin/env python3
"""Timing of the headline workload (32 x 3 x 2048 x 4096 bf16, bilinear equi2equi) over
families of rotations: how the kernel time depends on the view (tiles whose source
footprint fits no shared-memory box gather from global memory).  One JSON line each."""
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


def main():
    b = 32
    src = torch.rand((b, 3, 2048, 4096), device=DEV).to(torch.bfloat16)
    rng = np.random.default_rng(0)
    fam = {
        "identity": [{"roll": 0.0, "pitch": 0.0, "yaw": 0.0}] * b,
        "yaw only": [{"roll": 0.0, "pitch": 0.0, "yaw": 0.2 * i} for i in range(b)],
        "bench pattern": [{"roll": 0.01 * i, "pitch": 0.4 * math.sin(i), "yaw": 0.05 * i}
                          for i in range(b)],
        "roll 45 deg": [{"roll": math.pi / 4, "pitch": 0.1 * math.sin(i), "yaw": 0.2 * i}
                        for i in range(b)],
        "roll 90 deg": [{"roll": math.pi / 2, "pitch": 0.1 * math.sin(i), "yaw": 0.2 * i}
                        for i in range(b)],
        "pitch 45 deg": [{"roll": 0.0, "pitch": math.pi / 4, "yaw": 0.2 * i} for i in range(b)],
        "pitch 90 deg (pole on the horizon)": [{"roll": 0.0, "pitch": math.pi / 2, "yaw": 0.2 * i}
                                              for i in range(b)],
        "uniform random": [{"roll": float(rng.uniform(-math.pi, math.pi)),
                            "pitch": float(math.asin(rng.uniform(-1, 1))),
                            "yaw": float(rng.uniform(-math.pi, math.pi))} for _ in range(b)],
    }
    for name, rots in fam.items():
        ms = timed(lambda: equilib_b200.equi2equi(src, rots))
        ms_exact = timed(lambda: equilib_b200.equi2equi(src, rots, fast_coords=False), steps=5)
        ms_cubic = timed(lambda: equilib_b200.equi2equi(src, rots, mode="bicubic",
                                                        clip_output=False), steps=5)
        print(json.dumps({"rotations": name, "ms": round(ms, 3),
                          "gpix_s": round(b * 2048 * 4096 / ms / 1e6, 1),
                          "ms_exact_coords": round(ms_exact, 3),
                          "ms_bicubic_no_clamp": round(ms_cubic, 3)}), flush=True)


if __name__ == "__main__":
    main()
