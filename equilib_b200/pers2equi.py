"""Pers2Equi / pers2equi -- drop-in for equilib/pers2equi/base.py:19-170 (CUDA)."""

from __future__ import annotations

from typing import Dict, List, Union

import torch

from . import _common as K
from . import _lib, _prep

__all__ = ["Pers2Equi", "pers2equi"]

Rot = Union[Dict[str, float], List[Dict[str, float]]]


class Pers2Equi(object):
    """
    params:
    - height, width (int): equirectangular size
    - z_down (bool), mode (str), clip_output (bool)
    """

    def __init__(
        self,
        height: int,
        width: int,
        z_down: bool = False,
        mode: str = "bilinear",
        clip_output: bool = True,
    ) -> None:
        self.height = height
        self.width = width
        self.mode = mode
        self.z_down = z_down
        self.clip_output = clip_output

    def __call__(self, pers: torch.Tensor, rots: Rot, fov_x: float, **kwargs) -> torch.Tensor:
        return pers2equi(
            pers=pers,
            rots=rots,
            height=self.height,
            width=self.width,
            fov_x=fov_x,
            z_down=self.z_down,
            mode=self.mode,
            clip_output=self.clip_output,
            **kwargs,
        )


@K.accepts_float64
def pers2equi(
    pers: torch.Tensor,
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
    K.require_cuda_tensor(pers, "pers")
    K.check_backend(kwargs)
    opt = K.pop_options(kwargs)
    pers, rots, is_single = K.normalise_single(pers, rots)
    out = run(pers, rots, height=height, width=width, fov_x=fov_x, skew=skew, z_down=z_down,
              mode=mode, clip_output=clip_output, opt=opt)
    if is_single:
        out = out.squeeze(0)
    return out


def run(pers, rots, height, width, fov_x, skew, z_down, mode, clip_output=True,
        opt: K.Options = None) -> torch.Tensor:
    """Replaces equilib/pers2equi/torch.py:97-253.  Pixels outside the field of view
    are zeroed and *then* clipped (lines 246-251), so with clip_output they end up at
    clamp(0, min(pers), max(pers)); hence the min/max pass always runs for floats."""
    opt = opt or K.Options()
    pers = K.check_image(pers, "pers")
    assert len(pers) == len(rots), (
        f"ERR: length of pers and rot differs: {len(pers)} vs {len(rots)}"
    )
    K.mode_code(mode)
    bs, c, h_pers, w_pers = pers.shape
    dev = pers.device
    mats = _prep.mats_for("pers2equi", rots, z_down, dev,
                          (int(h_pers), int(w_pers), float(fov_x), float(skew)))
    tab_x, tab_y = _prep.equirect_tables(height, width, dev)
    clip = K.clip_tensor(pers, mode, clip_output, exact_clip=True, always=True,
                         group=opt.clip_group)
    out = torch.empty((bs, c, height, width), dtype=pers.dtype, device=dev)
    K.launch(
        pers, out, transform=_lib.PERS2EQUI, mode=mode,
        flags=K.flags_of(pers.dtype, opt), batch=bs, channels=c,
        src_hw=(h_pers, w_pers), dst_hw=(height, width),
        src_strides=pers.stride()[:3], dst_strides=out.stride()[:3],
        mats=mats, tab_x=tab_x, tab_y=tab_y, clip=clip,
    )
    return out
