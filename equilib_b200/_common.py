"""Shared host logic of the five transforms: input checks, dtype and clip policy,
and the call into the ``equilib_b200::remap`` op."""

from __future__ import annotations

import functools
import os
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib, _ops

Rot = Union[Dict[str, float], List[Dict[str, float]]]

SUPPORTED = (torch.uint8, torch.float16, torch.bfloat16, torch.float32)


def check_backend(kwargs: dict) -> None:
    """The reference threads ``backend=`` ("native" | "pure") and ``cache=`` through
    **kwargs (e.g. equilib/equi2pers/base.py:59-72,155).  Here there is one sampler, the
    fused CUDA one, and its parity target IS the reference's "native" backend
    (F.grid_sample semantics, equilib/grid_sample/torch/native.py:80-88): ``backend="native"``
    and ``"b200"`` both select it, so reference call sites that pass ``backend="native"``
    explicitly keep working.  ``"pure"`` (the reference's index-gather sampler, which wraps
    the seam and has a known tap bug, bilinear.py:47) is a different function and is not
    emulated: it raises.  ``cache`` is accepted and ignored (tables are cached per device
    in _prep)."""
    backend = kwargs.pop("backend", None)
    kwargs.pop("cache", None)
    if backend not in (None, "b200", "native"):
        raise ValueError(
            f"ERR: {backend} is not supported by equilib_b200 ('b200' / 'native' select the "
            "fused CUDA sampler); use the reference package for the 'pure' backend"
        )


class Options:
    """Extensions over the reference API; all default to the reference's behaviour /
    the documented policy.

    exact_clip      run the global min/max pass also where it is a <= 2 ulp no-op
    saturate_uint8  saturate the uint8 store instead of wrapping like the reference
    clip_group      torch.distributed group over which the clip range is all-reduced
                    (sharded batches, equilib_b200/sharded.py)
    fast_coords     None = library default (on for uint8 / fp16 / bf16 bilinear and bicubic on
                    large launches, off for fp32); True / False force the interpolated / exact coordinate chain
    """

    __slots__ = ("exact_clip", "saturate_uint8", "clip_group", "fast_coords")

    def __init__(self, exact_clip=False, saturate_uint8=False, clip_group=None, fast_coords=None):
        self.exact_clip = exact_clip
        self.saturate_uint8 = saturate_uint8
        self.clip_group = clip_group
        self.fast_coords = fast_coords


def pop_options(kwargs: dict) -> Options:
    opt = Options(
        exact_clip=bool(kwargs.pop("exact_clip", _env(b"EQUILIB_B200_EXACT_CLIP") == b"1")),
        saturate_uint8=bool(kwargs.pop("saturate_uint8", False)),
        clip_group=kwargs.pop("clip_group", None),
        fast_coords=kwargs.pop("fast_coords", None),
    )
    if kwargs:
        raise TypeError(f"unexpected keyword arguments: {sorted(kwargs)}")
    return opt


def require_cuda_tensor(x, name: str) -> None:
    if isinstance(x, np.ndarray):
        raise ValueError(
            f"ERR: `{name}` is a numpy array; equilib_b200 accelerates CUDA torch tensors only "
            "(use the reference's numpy path)"
        )
    if not torch.is_tensor(x):
        raise ValueError(f"ERR: `{name}` is neither numpy nor torch")
    if not x.is_cuda:
        raise ValueError(
            f"ERR: `{name}` is on {x.device}; equilib_b200 has no CPU path (move it to CUDA or "
            "use the reference's torch-CPU path)"
        )


def check_image(img: torch.Tensor, name: str, allow_channels_last: bool = False) -> torch.Tensor:
    assert len(img.shape) == 4, (
        f"ERR: input `{name}` should be 4-dim (b, c, h, w), but got {len(img.shape)}"
    )
    assert img.dtype in SUPPORTED + (torch.float64,), (
        f"ERR: input {name} has dtype of {img.dtype} which is\n"
        f"incompatible: try {SUPPORTED}"
    )
    if img.dtype == torch.float64:  # (the public functions convert first: accepts_float64)
        raise NotImplementedError("float64 images reach the kernels as float32 (accepts_float64)")
    if img.stride(3) != 1 and not (allow_channels_last and is_channels_last_u8(img)):
        img = img.contiguous()
    return img


_IMAGE_ARGS = ("src", "equi", "cubemap", "pers")


def _cast_tree(x, src_dtype, dst_dtype):
    if torch.is_tensor(x):
        return x.to(dst_dtype) if x.dtype == src_dtype else x
    if isinstance(x, list):
        return [_cast_tree(v, src_dtype, dst_dtype) for v in x]
    if isinstance(x, dict):
        return {k: _cast_tree(v, src_dtype, dst_dtype) for k, v in x.items()}
    return x


def _has_dtype(x, dtype) -> bool:
    if torch.is_tensor(x):
        return x.dtype == dtype
    if isinstance(x, list):
        return any(_has_dtype(v, dtype) for v in x)
    if isinstance(x, dict):
        return any(_has_dtype(v, dtype) for v in x.values())
    return False


def accepts_float64(fn):
    """float64 images (which the reference accepts, e.g. equilib/equi2equi/torch.py:87-95, and
    processes in float64) are computed in float32 and returned as float64: the kernels'
    arithmetic is the reference's float32 chain, so the result agrees with the reference's
    float64 result to float32 accuracy (~1e-6 relative on band-limited images), not bit for
    bit.  Documented deviation; every other dtype keeps its own policy."""

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        first = args[0] if args else next((kwargs[k] for k in _IMAGE_ARGS if k in kwargs), None)
        if not _has_dtype(first, torch.float64):
            return fn(*args, **kwargs)
        if args:
            args = (_cast_tree(args[0], torch.float64, torch.float32),) + tuple(args[1:])
        else:
            kwargs = {k: (_cast_tree(v, torch.float64, torch.float32) if k in _IMAGE_ARGS else v)
                      for k, v in kwargs.items()}
        return _cast_tree(fn(*args, **kwargs), torch.float32, torch.float64)

    return wrapper


def is_channels_last_u8(img: torch.Tensor) -> bool:
    """uint8 RGB / RGBX in torch.channels_last memory format -- what video decoders hand
    over.  equi2equi / equi2pers bilinear read and write this layout in
    place; every other call converts to planar first (``remap_new``)."""
    return (img.dim() == 4 and img.dtype == torch.uint8 and img.shape[1] in (3, 4)
            and img.stride(1) == 1 and img.stride(3) == img.shape[1]
            and img.stride(2) >= img.shape[1] * img.shape[3])


def normalise_single(img: torch.Tensor, rots: Rot):
    """equilib/equi2pers/base.py:119-130 (same in every rotation-taking transform)."""
    is_single = False
    if len(img.shape) == 3 and isinstance(rots, dict):
        img = img[None, ...]
        rots = [rots]
        is_single = True
    elif len(img.shape) == 3:
        img = img[:, None, ...]
    assert isinstance(rots, list), "ERR: rots is not a list"
    return img, rots, is_single


def clip_tensor(src: torch.Tensor, mode: str, clip_output: bool, exact_clip: bool,
                always: bool = False, group=None) -> Optional[torch.Tensor]:
    """Device tensor {lo, hi} for ``torch.clip(out, torch.min(src), torch.max(src))``
    (e.g. equilib/equi2equi/torch.py:195-198), or None when no clamp is launched.

    uint8 and clip_output=False never clip (reference).  For float images the global
    min/max pass is skipped where the clamp cannot change the stored result: nearest
    (outputs are input values) and bilinear (a convex combination: fp32 rounding can
    overshoot the maximum by at most 2 ulp, invisible after a 16-bit rounding; for
    fp32 pass ``exact_clip=True`` to reproduce the reference's clamp exactly).
    ``always`` is used by pers2equi, whose fill value depends on the clamp."""
    if src.dtype == torch.uint8 or not clip_output:
        return None
    if not (always or exact_clip or mode == "bicubic"):
        return None
    out2 = torch.empty(2, dtype=torch.float32, device=src.device)
    _ops.minmax(src, out2)
    return reduce_clip(out2, group)


def reduce_clip(out2: torch.Tensor, group) -> torch.Tensor:
    """All-reduce a (min, max) pair over ``group`` (no-op without one)."""
    if group is None:
        return out2
    from .sharded import allreduce_minmax

    return allreduce_minmax(out2, None if group is True else group)


# Environment switches are read on every call (tests and experiments flip them at run time).
# os.environ.get() costs ~1.4 us per key (str -> bytes encoding + wrappers) and there are ten
# of them: a third of the host time of a small launch (BASELINE config 1).  On POSIX the
# mapping behind os.environ is a plain dict of bytes that putenv / monkeypatch keep in sync.
_ENV_DATA = getattr(os.environ, "_data", None)
if not (isinstance(_ENV_DATA, dict) and all(isinstance(k, bytes) for k in list(_ENV_DATA)[:4])):
    _ENV_DATA = None


def _env(name: bytes):
    """Value of an EQUILIB_B200_* switch as bytes, or None."""
    if _ENV_DATA is not None:
        return _ENV_DATA.get(name)
    v = os.environ.get(name.decode())
    return None if v is None else v.encode()


_ENV_FLAGS = (
    (b"EQUILIB_B200_NO_STAGE", _lib.FLAG_NO_STAGE),
    (b"EQUILIB_B200_NO_TMA", _lib.FLAG_NO_TMA),
    (b"EQUILIB_B200_FORCE_STAGE", _lib.FLAG_FORCE_STAGE),
    (b"EQUILIB_B200_NO_LEAN", _lib.FLAG_NO_LEAN),
    (b"EQUILIB_B200_NO_LEAN2", _lib.FLAG_NO_LEAN2),
    (b"EQUILIB_B200_FORCE_LEAN2", _lib.FLAG_FORCE_LEAN2),
)


def flags_of(dtype: torch.dtype, opt: "Options") -> int:
    flags = _lib.FLAG_SATURATE_U8 if (opt.saturate_uint8 and dtype == torch.uint8) else 0
    fast = opt.fast_coords
    if fast is None:
        e = _env(b"EQUILIB_B200_FAST_COORDS")
        if e == b"0" or e == b"1":
            fast = e == b"1"
    if fast is True:
        flags |= _lib.FLAG_FAST_COORDS
    elif fast is False:
        flags |= _lib.FLAG_EXACT_COORDS
    # experiments only: EQUILIB_B200_WARP_SHAPE = 1 (32x1), 2 (16x2), 3 (8x4) forces the
    # warp tile shape; unset/0 lets the library choose
    ws = _env(b"EQUILIB_B200_WARP_SHAPE")
    if ws:
        flags |= (int(ws) & 3) << _lib.FLAG_WARP_SHAPE_SHIFT
    for name, bit in _ENV_FLAGS:
        if _env(name) == b"1":
            flags |= bit
    return flags


def mode_code(mode: str) -> int:
    if mode not in _lib.MODES:
        raise ValueError(f"ERR: {mode} is not supported")
    return _lib.MODES[mode]


def remap_new(src: torch.Tensor, out_shape, *, allow_channels_last: bool = False,
              **kw) -> torch.Tensor:
    """Allocate the output and launch.  A channels-last uint8 source keeps its layout
    (output in torch.channels_last too) where the library serves it; otherwise it is
    converted to planar, like every input of the reference."""
    if src.stride(3) != 1:
        if allow_channels_last:
            out = torch.empty(out_shape, dtype=src.dtype, device=src.device,
                              memory_format=torch.channels_last)
            cl = dict(kw, flags=kw["flags"] | _lib.FLAG_CHANNELS_LAST,
                      src_strides=src.stride()[:3], dst_strides=out.stride()[:3])
            if launch(src, out, probe_channels_last=True, **cl):
                return out
        src = src.contiguous()
    out = torch.empty(out_shape, dtype=src.dtype, device=src.device)
    launch(src, out, src_strides=src.stride()[:3], dst_strides=out.stride()[:3], **kw)
    return out


def launch(src: torch.Tensor, out: torch.Tensor, *, transform: int, mode: str, flags: int,
           batch: int, channels: int, src_hw: Tuple[int, int], dst_hw: Tuple[int, int],
           src_strides: Sequence[int], dst_strides: Sequence[int], mats=None, tab_x=None,
           tab_y=None, itab_x=None, grid=None, grid_stride_b: int = 0, clip=None,
           w_face: int = 0, n_cells: int = 0, cell_offset: Optional[Sequence[int]] = None,
           face_ptrs=None, probe_channels_last: bool = False) -> bool:
    """One eqb_remap launch.  ``probe_channels_last``: ask the library first whether it
    serves this channels-last call; returns False without launching if it does not."""
    offs = list(cell_offset) if cell_offset is not None else []
    offs = offs + [0] * (12 - len(offs))
    args = (
        src, out, mats, tab_x, tab_y, itab_x, grid, clip, face_ptrs,
        transform, mode_code(mode), flags, batch, channels,
        src_hw[0], src_hw[1], dst_hw[0], dst_hw[1],
        [int(s) for s in src_strides], [int(s) for s in dst_strides],
        int(grid_stride_b), w_face, n_cells, offs,
    )
    if probe_channels_last and not _ops.supports_channels_last(*args):
        return False
    _ops.remap(*args)
    return True
