"""Equi2Pers / equi2pers -- drop-in for equilib/equi2pers/base.py:19-164 (CUDA)."""

from __future__ import annotations

from typing import Dict, List, Union

import torch

import numpy as np

from . import _common as K
from . import _lib, _ops, _prep

__all__ = ["Equi2Pers", "equi2pers", "get_bounding_fov"]

Rot = Union[Dict[str, float], List[Dict[str, float]]]


class Equi2Pers(object):
    """
    params:
    - height, width (int): perspective size
    - fov_x (float): perspective image fov of x-axis (degrees)
    - skew (float), z_down (bool), mode (str), clip_output (bool)
    """

    def __init__(
        self,
        height: int,
        width: int,
        fov_x: float,
        skew: float = 0.0,
        z_down: bool = False,
        mode: str = "bilinear",
        clip_output: bool = True,
    ) -> None:
        self.height = height
        self.width = width
        self.fov_x = fov_x
        self.skew = skew
        self.mode = mode
        self.z_down = z_down
        self.clip_output = clip_output

    def __call__(self, equi: torch.Tensor, rots: Rot, **kwargs) -> torch.Tensor:
        return equi2pers(
            equi=equi,
            rots=rots,
            height=self.height,
            width=self.width,
            fov_x=self.fov_x,
            skew=self.skew,
            z_down=self.z_down,
            mode=self.mode,
            clip_output=self.clip_output,
            **kwargs,
        )


    def get_bounding_fov(self, equi: torch.Tensor, rots: Rot) -> np.ndarray:
        return get_bounding_fov(
            equi=equi,
            rots=rots,
            height=self.height,
            width=self.width,
            fov_x=self.fov_x,
            skew=self.skew,
            z_down=self.z_down,
        )


@K.accepts_float64
def equi2pers(
    equi: torch.Tensor,
    rots: Rot,
    height: int,
    width: int,
    fov_x: float,
    skew: float = 0.0,
    mode: str = "bilinear",
    z_down: bool = False,
    clip_output: bool = True,
    **kwargs,
) -> torch.Tensor:
    K.require_cuda_tensor(equi, "equi")
    K.check_backend(kwargs)
    opt = K.pop_options(kwargs)
    equi, rots, is_single = K.normalise_single(equi, rots)
    out = run(equi, rots, height=height, width=width, fov_x=fov_x, skew=skew, z_down=z_down,
              mode=mode, clip_output=clip_output, opt=opt)
    if is_single:
        out = out.squeeze(0)
    return out


def run(equi, rots, height, width, fov_x, skew, z_down, mode, clip_output=True,
        opt: K.Options = None) -> torch.Tensor:
    """Replaces equilib/equi2pers/torch.py:96-244 with one fused launch."""
    opt = opt or K.Options()
    equi = K.check_image(equi, "equi", allow_channels_last=True)
    assert len(equi) == len(rots), (
        f"ERR: length of equi and rot differs: {len(equi)} vs {len(rots)}"
    )
    K.mode_code(mode)
    bs, c, h_equi, w_equi = equi.shape
    dev = equi.device
    mats = _prep.mats_for("equi2pers", rots, z_down, dev,
                          (int(height), int(width), float(fov_x), float(skew)))
    clip = K.clip_tensor(equi, mode, clip_output, opt.exact_clip, group=opt.clip_group)
    return K.remap_new(
        equi, (bs, c, height, width), allow_channels_last=mode == "bilinear",
        transform=_lib.EQUI2PERS, mode=mode,
        flags=K.flags_of(equi.dtype, opt), batch=bs, channels=c,
        src_hw=(h_equi, w_equi), dst_hw=(height, width), mats=mats, clip=clip,
    )


def get_bounding_fov(
    equi: torch.Tensor,
    rots: Rot,
    height: int,
    width: int,
    fov_x: float,
    skew: float = 0.0,
    z_down: bool = False,
) -> np.ndarray:
    """Perimeter of the perspective view in equirectangular pixel coordinates,
    (B, 2*width + 2*(height-2), 2) int64 as (row, col), clockwise from the top-left
    corner -- drop-in for equilib/equi2pers/base.py:166-226 / torch.py:247-351.

    The coordinates come from the same device chain the remap uses
    (``eqb_remap_coords``); only the perimeter is copied back to the host."""
    K.require_cuda_tensor(equi, "equi")
    equi, rots, is_single = K.normalise_single(equi, rots)
    equi = K.check_image(equi, "equi", allow_channels_last=True)
    assert len(equi) == len(rots), (
        f"ERR: length of equi and rot differs: {len(equi)} vs {len(rots)}"
    )
    bs, _, h_equi, w_equi = equi.shape
    dev = equi.device
    mats = _prep.mats_for("equi2pers", rots, z_down, dev,
                          (int(height), int(width), float(fov_x), float(skew)))
    grid = torch.empty((bs, 2, height, width), dtype=torch.float32, device=dev)
    _ops.coords(grid, mats, None, None, None, None, _lib.EQUI2PERS, bs, h_equi, w_equi,
                height, width)
    top = grid[:, :, 0, :]
    right = grid[:, :, 1:, width - 1]
    bottom = grid[:, :, height - 1, 1:width - 1].flip(-1)
    left = grid[:, :, 1:, 0].flip(-1)
    bboxs = torch.cat([top, right, bottom, left], dim=-1).permute(0, 2, 1)
    assert bboxs.shape[1] == width * 2 + (height - 2) * 2
    out = np.rint(bboxs.cpu().numpy()).astype(np.int64)
    if is_single:
        out = out.squeeze(0)
    return out
