"""Equi2Cube2Equi / equi2cube2equi -- equirect -> cubemap -> equirect in one launch.

SURVEY 8f N2.  With the reference the chain is two calls,

    cube = equi2cube(equi, rots, w_face, cube_format, z_down, mode, clip_output)
    out = cube2equi(cube, cube_format, height, width, mode, clip_output)

(equilib/equi2cube/base.py:94-154, equilib/cube2equi/base.py:85-160), which write the
cubemap, re-format it (horizon <-> dice / list / dict copies, equilib/equi2cube/torch.py:17-76,
equilib/cube2equi/torch.py:15-132) and read it back.  ``equi2cube2equi`` returns the same
``out`` -- the cube format cannot matter: the reference converts every format to the horizon
strip before sampling -- from one kernel in which the cubemap never touches memory
(``eqb_equi2cube2equi``, csrc/remap_chain.cu): bit-equal to this package's own two calls with
``fast_coords=False``.

The fused kernel serves nearest and bilinear, every dtype of the package, up to 4 channels, no
clamp (for these modes ``clip_output`` cannot change a stored value beyond 2 ulp, see
``_common.clip_tensor``; ``exact_clip=True`` and bicubic always take the two launches).

``fused=None`` (default) picks by measurement (16 x 4K frames, w_face 1024, B200,
profiles/r02_round2b.jsonl): **nearest on 16/32-bit images** runs fused -- one chain and one
gather per output pixel, no cubemap: 1.43 ms instead of 2.02 (fp32), 1.44 instead of 1.66
(bf16) -- while uint8 nearest (1.48 fused, 1.37 - 1.54 as two launches) and **bilinear** run the
two launches: fused, every bilinear output pixel evaluates four cube pixels of four source taps
each (48 gathers per pixel for RGB instead of 12 + 12 staged ones), 5.5 ms against 2.5 (fp32) /
1.8 (bf16).  ``fused=True`` forces the single launch (no intermediate: 2.4 GB less memory for
16 fp32 frames), ``fused=False`` the two launches.
"""

from __future__ import annotations

from typing import Dict, List, Union

import torch

from . import _common as K
from . import _ops, _prep
from .cube2equi import run as _cube2equi_run
from .equi2cube import run as _equi2cube_run

__all__ = ["Equi2Cube2Equi", "equi2cube2equi"]

Rot = Union[Dict[str, float], List[Dict[str, float]]]


class Equi2Cube2Equi(object):
    """
    params:
    - w_face (int): width of the intermediate cube faces
    - height, width (int): output equirectangular size
    - z_down (bool), clip_output (bool), mode (str): as Equi2Cube / Cube2Equi
    """

    def __init__(self, w_face: int, height: int, width: int, z_down: bool = False,
                 clip_output: bool = True, mode: str = "bilinear") -> None:
        self.w_face = w_face
        self.height = height
        self.width = width
        self.z_down = z_down
        self.clip_output = clip_output
        self.mode = mode

    def __call__(self, equi: torch.Tensor, rots: Rot, **kwargs) -> torch.Tensor:
        return equi2cube2equi(equi=equi, rots=rots, w_face=self.w_face, height=self.height,
                              width=self.width, z_down=self.z_down,
                              clip_output=self.clip_output, mode=self.mode, **kwargs)


def fused_ok(equi: torch.Tensor, mode: str, clip_output: bool, opt: K.Options) -> bool:
    """Can the one-launch chain serve this call?  (Else: the two launches.)"""
    if mode not in ("nearest", "bilinear") or equi.shape[1] > 4:
        return False
    # a clamp is launched only where it can change a value (_common.clip_tensor)
    return not (equi.dtype != torch.uint8 and clip_output and opt.exact_clip)


@K.accepts_float64
def equi2cube2equi(
    equi: torch.Tensor,
    rots: Rot,
    w_face: int,
    height: int,
    width: int,
    z_down: bool = False,
    clip_output: bool = True,
    mode: str = "bilinear",
    **kwargs,
) -> torch.Tensor:
    K.require_cuda_tensor(equi, "equi")
    K.check_backend(kwargs)
    fused = kwargs.pop("fused", None)
    opt = K.pop_options(kwargs)
    equi, rots, is_single = K.normalise_single(equi, rots)
    equi = K.check_image(equi, "equi")
    assert len(equi) == len(rots), (
        f"ERR: length of equi and rot differs: {len(equi)} vs {len(rots)}"
    )
    K.mode_code(mode)
    # cube2equi's size rule (equilib/cube2equi/base.py:94)
    assert height % 8 == 0 and width % 8 == 0, "ERR: height and width must be multiples of 8"
    if fused is None:
        fused = mode == "nearest" and equi.dtype != torch.uint8
    if fused and fused_ok(equi, mode, clip_output, opt):
        out = run(equi, rots, int(w_face), int(height), int(width), z_down, mode, opt)
    else:
        cube = _equi2cube_run(equi, rots, w_face=w_face, cube_format="horizon", z_down=z_down,
                        mode=mode, clip_output=clip_output, opt=opt)
        out = _cube2equi_run(cube, "horizon", height=height, width=width, mode=mode,
                       clip_output=clip_output, opt=opt)
    # cube2equi squeezes whenever the batch is 1 (equilib/cube2equi/base.py:157-158)
    if is_single or out.shape[0] == 1:
        out = out.squeeze(0)
    return out


def run(equi, rots, w_face, height, width, z_down, mode, opt: K.Options) -> torch.Tensor:
    bs, c = equi.shape[:2]
    dev = equi.device
    if equi.stride(3) != 1:
        equi = equi.contiguous()
    # the reference flips z_down for equi2cube (equi2cube/torch.py:206-209)
    mats = _prep.mats_for("rot", rots, not z_down, dev)
    rng = _prep.cube_rng_table(w_face, dev)
    tab_x, tab_y, itab_x = _prep.cube2equi_tables(height, width, dev)
    out = torch.empty((bs, c, height, width), dtype=equi.dtype, device=dev)
    _ops.cube_chain(equi, out, mats, rng, tab_x, tab_y, itab_x, K.mode_code(mode),
                    K.flags_of(equi.dtype, opt), w_face)
    return out
