"""Equi2Cube / equi2cube -- drop-in for equilib/equi2cube/base.py:35-154 (CUDA).

The kernel writes each cube face straight into its final place: the horizon strip,
the 3x4 dice canvas (background zero-filled by the same launch) or a face-major
buffer whose slices are returned as the ``list`` / ``dict`` formats.  The
reference's horizon -> list/dict/dice re-formatting passes
(equilib/equi2cube/torch.py:17-76,242-251) do not exist here.
"""

from __future__ import annotations

from typing import Dict, List, Union

import torch

from . import _common as K
from . import _lib, _prep

__all__ = ["Equi2Cube", "equi2cube"]

Rot = Union[Dict[str, float], List[Dict[str, float]]]

FACE_KEYS = ("F", "R", "B", "L", "U", "D")
# dice cell (sx, sy) of F, R, B, L, U, D -- equilib/equi2cube/torch.py:69
DICE_CELLS = ((1, 1), (2, 1), (3, 1), (0, 1), (1, 0), (1, 2))
DICE_BACKGROUND = ((0, 0), (2, 0), (3, 0), (0, 2), (2, 2), (3, 2))
CUBE_FORMATS = ("horizon", "list", "dict", "dice")


class Equi2Cube(object):
    """
    params:
    - w_face (int): cube face width
    - cube_format (str): ("dice", "horizon", "dict", "list")
    - z_down (bool), clip_output (bool), mode (str)
    """

    def __init__(
        self,
        w_face: int,
        cube_format: str,
        z_down: bool = False,
        clip_output: bool = True,
        mode: str = "bilinear",
    ) -> None:
        self.w_face = w_face
        self.cube_format = cube_format
        self.z_down = z_down
        self.clip_output = clip_output
        self.mode = mode

    def __call__(self, equi: torch.Tensor, rots: Rot, **kwargs):
        return equi2cube(
            equi=equi,
            rots=rots,
            w_face=self.w_face,
            cube_format=self.cube_format,
            z_down=self.z_down,
            clip_output=self.clip_output,
            mode=self.mode,
            **kwargs,
        )


@K.accepts_float64
def equi2cube(
    equi: torch.Tensor,
    rots: Rot,
    w_face: int,
    cube_format: str,
    z_down: bool = False,
    clip_output: bool = True,
    mode: str = "bilinear",
    **kwargs,
):
    K.require_cuda_tensor(equi, "equi")
    K.check_backend(kwargs)
    opt = K.pop_options(kwargs)
    equi, rots, is_single = K.normalise_single(equi, rots)
    out = run(equi, rots, w_face=w_face, cube_format=cube_format, z_down=z_down, mode=mode,
              clip_output=clip_output, opt=opt)
    if is_single:
        out = out[0]
    return out


def run(equi, rots, w_face, cube_format, z_down, mode, clip_output=True, opt: K.Options = None):
    """Replaces equilib/equi2cube/torch.py:103-253 with one fused launch."""
    opt = opt or K.Options()
    equi = K.check_image(equi, "equi")
    assert len(equi) == len(rots), (
        f"ERR: length of equi and rot differs: {len(equi)} vs {len(rots)}"
    )
    K.mode_code(mode)
    if cube_format not in CUBE_FORMATS:
        raise NotImplementedError("{} is not supported".format(cube_format))
    bs, c, h_equi, w_equi = equi.shape
    dev = equi.device
    w = int(w_face)

    # the reference flips z_down for this transform (equi2cube/torch.py:206-209)
    mats = _prep.mats_for("rot", rots, not z_down, dev)
    rng = _prep.cube_rng_table(w, dev)
    clip = K.clip_tensor(equi, mode, clip_output, opt.exact_clip, group=opt.clip_group)

    if cube_format == "horizon":
        out = torch.empty((bs, c, w, 6 * w), dtype=equi.dtype, device=dev)
        strides = out.stride()[:3]
        cells = [f * w for f in range(6)]
    elif cube_format == "dice":
        out = torch.empty((bs, c, 3 * w, 4 * w), dtype=equi.dtype, device=dev)
        strides = out.stride()[:3]
        cells = [sy * w * 4 * w + sx * w for sx, sy in DICE_CELLS + DICE_BACKGROUND]
    else:  # list / dict: face-major buffer (B, 6, C, w, w)
        out = torch.empty((bs, 6, c, w, w), dtype=equi.dtype, device=dev)
        strides = (out.stride(0), out.stride(2), out.stride(3))
        cells = [f * out.stride(1) for f in range(6)]

    K.launch(
        equi, out, transform=_lib.EQUI2CUBE, mode=mode,
        flags=K.flags_of(equi.dtype, opt), batch=bs, channels=c,
        src_hw=(h_equi, w_equi), dst_hw=(w, len(cells) * w),
        src_strides=equi.stride()[:3], dst_strides=strides,
        mats=mats, tab_x=rng, clip=clip, w_face=w, n_cells=len(cells), cell_offset=cells,
    )

    if cube_format == "list":
        return [[out[b, f] for f in range(6)] for b in range(bs)]
    if cube_format == "dict":
        return [{k: out[b, f] for f, k in enumerate(FACE_KEYS)} for b in range(bs)]
    return out
