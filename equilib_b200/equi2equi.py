"""Equi2Equi / equi2equi -- drop-in for equilib/equi2equi/base.py:19-134 when the
input is a CUDA tensor.  Same names, argument order, defaults and shape rules."""

from __future__ import annotations

from typing import Dict, List, Optional, Union

import torch

from . import _common as K
from . import _lib, _prep

__all__ = ["Equi2Equi", "equi2equi"]

Rot = Union[Dict[str, float], List[Dict[str, float]]]


class Equi2Equi(object):
    """
    params:
    - height, width (optional int): stored but, as in the reference
      (equilib/equi2equi/base.py:42-43,51-60), not forwarded by ``__call__``
    - clip_output (bool), mode (str), z_down (bool)
    """

    def __init__(
        self,
        height: Optional[int] = None,
        width: Optional[int] = None,
        clip_output: bool = True,
        mode: str = "bilinear",
        z_down: bool = False,
    ) -> None:
        self.height = height
        self.width = width
        self.clip_output = clip_output
        self.mode = mode
        self.z_down = z_down

    def __call__(self, src: torch.Tensor, rots: Rot, **kwargs) -> torch.Tensor:
        return equi2equi(
            src=src,
            rots=rots,
            clip_output=self.clip_output,
            mode=self.mode,
            z_down=self.z_down,
            **kwargs,
        )


@K.accepts_float64
def equi2equi(
    src: torch.Tensor,
    rots: Rot,
    clip_output: bool = True,
    mode: str = "bilinear",
    z_down: bool = False,
    height: Optional[int] = None,
    width: Optional[int] = None,
    **kwargs,
) -> torch.Tensor:
    K.require_cuda_tensor(src, "src")
    K.check_backend(kwargs)
    opt = K.pop_options(kwargs)
    src, rots, is_single = K.normalise_single(src, rots)
    out = run(src, rots, z_down=z_down, mode=mode, height=height, width=width,
              clip_output=clip_output, opt=opt)
    if is_single:
        out = out.squeeze(0)
    return out


def run(src, rots, z_down, mode, height=None, width=None, clip_output=True,
        opt: K.Options = None) -> torch.Tensor:
    """Replaces equilib/equi2equi/torch.py:45-200 with one fused launch."""
    opt = opt or K.Options()
    src = K.check_image(src, "src", allow_channels_last=True)
    assert len(src) == len(rots), (
        f"ERR: length of `src` and `rot` differs: {len(src)} vs {len(rots)}"
    )
    K.mode_code(mode)
    bs, c, h_equi, w_equi = src.shape
    assert type(height) == type(width), (  # noqa: E721
        "ERR: `height` and `width` does not match types (maybe it was set separately?)"
    )
    if height is None and width is None:
        height, width = h_equi, w_equi
    else:
        assert isinstance(height, int) and isinstance(width, int)

    dev = src.device
    mats = _prep.mats_for("rot", rots, z_down, dev)
    tab_x, tab_y = _prep.equirect_tables(height, width, dev)
    clip = K.clip_tensor(src, mode, clip_output, opt.exact_clip, group=opt.clip_group)
    flags = K.flags_of(src.dtype, opt)
    if (mode == "nearest" and opt.fast_coords is None
            and all(r["roll"] == 0 and r["pitch"] == 0 for r in rots)):
        # identity / pure yaw: every row coordinate is y + 0.5, a rounding tie on every pixel --
        # interpolated coordinates could decide none of them (the kernel would notice per tile
        # and take its slower general path): ask for the exact chain up front
        flags = (flags | _lib.FLAG_EXACT_COORDS) & ~_lib.FLAG_FAST_COORDS
    return K.remap_new(
        src, (bs, c, height, width), allow_channels_last=mode == "bilinear",
        transform=_lib.EQUI2EQUI, mode=mode, flags=flags,
        batch=bs, channels=c, src_hw=(h_equi, w_equi), dst_hw=(height, width),
        mats=mats, tab_x=tab_x, tab_y=tab_y, clip=clip,
    )
