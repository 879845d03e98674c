"""Performance port of the reference's torch-CPU path, for the CPU baseline only.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py): used by bench.py's
``cpu_baseline`` leg and ``--impl reference`` arm, and checked against
oracle/remap_oracle.py in tests/test_reference_port.py.

bench.py times the UNMODIFIED reference package from baseline/_ref whenever that directory
is present (it is git-ignored but travels with the repo snapshot; kind "reference").  This
file is the fallback for a checkout without it (kind "port"): it restates what the
reference's torch path *executes* on CPU tensors, op for op, so that its timing stands in
for the reference's:

  equilib/equi2equi/torch.py:143-198   grid prep (cached, as the class API does)
                                        -> matmul -> convert_grid -> native() -> clip
  equilib/grid_sample/torch/native.py:44-88  normalise + F.grid_sample(align_corners=
                                        True, padding_mode="reflection")

Unlike oracle/remap_oracle.py this file DOES call ``torch.nn.functional.grid_sample``
-- that library call is precisely what the reference spends its sampler time in.
It materialises the same full-size intermediates as the reference: the
(B, H, W, 3) ray tensor ``m`` (cached across calls like ``Equi2Equi._cache``), the
rotated rays ``M`` and the (B, 2, H, W) grid.
"""

from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import remap_oracle as O

PI = O.PI


class Equi2EquiCPU:
    """Stand-in for ``equilib.Equi2Equi(mode=...)`` on CPU float32 tensors."""

    def __init__(self, mode: str = "bilinear", z_down: bool = False, clip_output: bool = True):
        self.mode = mode
        self.z_down = z_down
        self.clip_output = clip_output
        self._cache: Dict = {}

    def _rays(self, bs: int, h: int, w: int) -> torch.Tensor:
        key = (bs, h, w)
        if key not in self._cache:
            # torch_utils/grid.py:63-116, tiled per batch item like the reference
            xs = torch.linspace(0, w - 1, w)
            ys = torch.linspace(0, h - 1, h)
            theta = xs * 2 * PI / w - PI
            phi = ys * PI / h - PI / 2
            phi, theta = torch.meshgrid([phi, theta], indexing="ij")
            m = torch.stack(
                (torch.cos(phi) * torch.cos(theta), torch.cos(phi) * torch.sin(theta),
                 torch.sin(phi)), dim=-1)
            self._cache.clear()
            self._cache[key] = torch.cat([m.unsqueeze(0)] * bs).unsqueeze(-1)
        return self._cache[key]

    def __call__(self, src: torch.Tensor, rots: List[Dict[str, float]]) -> torch.Tensor:
        assert src.dim() == 4 and src.dtype == torch.float32 and not src.is_cuda
        bs, _, h, w = src.shape
        m = self._rays(bs, h, w)
        R = O.rotation_matrices(rots, self.z_down)
        M = torch.matmul(R[:, None, None, ...], m).squeeze(-1)
        phi = torch.asin(M[..., 2] / torch.norm(M, dim=-1))
        theta = torch.atan2(M[..., 1], M[..., 0])
        ui = (theta - PI) * w / (2 * PI)
        uj = (phi + PI / 2) * h / PI
        ui += 0.5
        uj += 0.5
        ui %= w
        uj.clamp_(0, h - 1)
        grid = torch.stack((uj, ui), dim=-3).permute(0, 2, 3, 1)
        norm_uj = torch.clamp(2 * grid[..., 0] / (h - 1) - 1, -1, 1)
        norm_ui = torch.clamp(2 * grid[..., 1] / (w - 1) - 1, -1, 1)
        grid[..., 0] = norm_ui
        grid[..., 1] = norm_uj
        out = F.grid_sample(src, grid, mode=self.mode, align_corners=True,
                            padding_mode="reflection")
        if self.clip_output:
            out = torch.clip(out, torch.min(src), torch.max(src))
        return out
