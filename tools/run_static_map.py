#!/usr/bin/env python3
"""The static-camera path alone (for ncu): build the map of the bench rotations, apply it
to 32 x 3 x 2048 x 4096 bf16 frames a few times."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import equilib_b200  # noqa: E402

dev = torch.device("cuda", 0)
b = 32
src = torch.rand((b, 3, 2048, 4096), device=dev).to(torch.bfloat16)
rots = [{"roll": 0.01 * i, "pitch": 0.4 * math.sin(i), "yaw": 0.05 * i} for i in range(b)]
sr = equilib_b200.StaticRemap.equi2equi(rots, src.shape, src.dtype, dev)
for _ in range(3):
    out = sr(src)
torch.cuda.synchronize()
print("ok", float(out.float().mean()))
