"""grid_sample -- drop-in for equilib/grid_sample/torch/grid_sample.py:16-86 with the
``native`` backend (native.py:11-90), for callers that already hold a pixel-unit
(y, x) grid.  ``F.grid_sample`` is not called; unlike the reference the grid is not
mutated (quirk Q9)."""

from __future__ import annotations

from typing import Optional

import torch

from . import _common as K
from . import _lib

__all__ = ["grid_sample"]


def grid_sample(
    img: torch.Tensor,
    grid: torch.Tensor,
    out: Optional[torch.Tensor] = None,
    mode: str = "bilinear",
    backend: str = "b200",
) -> torch.Tensor:
    """img (B, C, H, W); grid (B, 2, Hout, Wout) in source pixel units, (y, x) order.
    Returns (B, C, Hout, Wout) in img's dtype (uint8 images are sampled in fp32 and
    truncated, like the transforms do)."""
    if backend not in ("b200", "native"):
        raise ValueError(f"ERR: {backend} is not supported")
    K.require_cuda_tensor(img, "img")
    K.require_cuda_tensor(grid, "grid")
    assert img.device == grid.device, (
        "ERR: the devices of `img` and `grid` need to be on the same device"
    )
    single = img.dim() == 3
    if single:
        img, grid = img[None], grid[None]
    img = K.check_image(img, "img")
    assert grid.dim() == 4 and grid.shape[1] == 2 and grid.shape[0] in (1, img.shape[0])
    g = grid.to(torch.float32).contiguous()
    bs, c, h, w = img.shape
    ho, wo = g.shape[2:]
    res = torch.empty((bs, c, ho, wo), dtype=img.dtype, device=img.device)
    K.launch(
        img, res, transform=_lib.GRID, mode=mode, flags=0, batch=bs, channels=c,
        src_hw=(h, w), dst_hw=(ho, wo), src_strides=img.stride()[:3],
        dst_strides=res.stride()[:3], grid=g,
        grid_stride_b=(g.stride(0) if g.shape[0] > 1 else 0),
    )
    if out is not None:
        out.copy_(res)
        res = out
    return res[0] if single else res
