#!/usr/bin/env python3
"""Minimal driver for ncu: N calls of one workload (default: the headline 32 x 3 x 2048 x 4096
bf16 bilinear equi2equi on the bench rotations).  Env: DTYPE (bf16|f16|u8|f32), B, MODE,
ROTS (bench|random|identity), FAST (0|1|unset), N."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import equilib_b200  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    dn = os.environ.get("DTYPE", "bf16")
    b = int(os.environ.get("B", "32"))
    mode = os.environ.get("MODE", "bilinear")
    n = int(os.environ.get("N", "4"))
    fast = {"0": False, "1": True}.get(os.environ.get("FAST", ""), None)
    img = torch.rand((b, 3, 2048, 4096), device=dev)
    src = (img * 255).to(torch.uint8) if dn == "u8" else img.to(
        {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[dn])
    del img
    rng = np.random.default_rng(0)
    rots = {
        "bench": [{"roll": 0.01 * i, "pitch": 0.4 * math.sin(i), "yaw": 0.05 * i} for i in range(b)],
        "random": [{"roll": float(rng.uniform(-math.pi, math.pi)),
                    "pitch": float(math.asin(rng.uniform(-1, 1))),
                    "yaw": float(rng.uniform(-math.pi, math.pi))} for _ in range(b)],
        "identity": [{"roll": 0.0, "pitch": 0.0, "yaw": 0.0}] * b,
    }[os.environ.get("ROTS", "bench")]
    for _ in range(n):
        out = equilib_b200.equi2equi(src, rots, mode=mode, fast_coords=fast)
    torch.cuda.synchronize()
    print("ok", tuple(out.shape), out.dtype)


if __name__ == "__main__":
    main()
