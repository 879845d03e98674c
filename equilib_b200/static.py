"""Static-camera remap: the sampling map of a fixed set of rotations, built once.

For video from a fixed rig the rotations of every call are the same; the reference
recomputes rotation, spherical mapping and normalisation per call anyway (its grid cache,
equilib/_cache.py:21-40, keeps only the rays) and lists "full-grid cache for the
fixed-rotation case" as future work (README.md:168).  ``StaticRemap`` is that cache for
the CUDA path: 8 bytes per output pixel (top-left tap and the two bilinear fractions) -- or 4
with ``compact=True`` (tap relative to the tile's box, 8-bit fractions) -- plus the
shared-memory box of every output tile.  Applying it runs none of the coordinate
chain -- the kernel reads the map, loads the box with TMA and blends.

    sr = StaticRemap.equi2equi(rots, src_shape=frames.shape, dtype=frames.dtype,
                               device=frames.device)
    out = sr(frames)            # same result as equi2equi(frames, rots, fast_coords=False)
                                # up to the 16-bit fractions of the map

Bilinear only; 1, 3 or 4 planar channels of uint8 / float16 / bfloat16, or uint8 RGB / RGBX
frames in torch.channels_last.
"""

from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _common as K
from . import _lib, _ops, _prep

__all__ = ["StaticRemap"]


class StaticRemap:
    def __init__(self, transform: int, mats, tab_x, tab_y, src_shape: Tuple[int, int, int, int],
                 dst_hw: Tuple[int, int], dtype: torch.dtype, device: torch.device,
                 compact: bool = False) -> None:
        b, c, sh, sw = (int(v) for v in src_shape)
        dh, dw = int(dst_hw[0]), int(dst_hw[1])
        if dtype not in (torch.uint8, torch.float16, torch.bfloat16):
            raise NotImplementedError("StaticRemap serves uint8, float16 and bfloat16 frames")
        if c not in (1, 3, 4):
            raise NotImplementedError("StaticRemap serves 1, 3 or 4 channels")
        if dw % 4:
            raise NotImplementedError("StaticRemap needs an output width that is a multiple of 4")
        self.transform, self.dtype, self.device = transform, dtype, device
        self.src_shape, self.dst_hw = (b, c, sh, sw), (dh, dw)
        self._mats, self._tab_x, self._tab_y = mats, tab_x, tab_y
        coords = torch.empty((b, 2, dh, dw), dtype=torch.float32, device=device)
        _ops.coords(coords, mats, tab_x, tab_y, None, None, transform, b, sh, sw, dh, dw)
        ty = (dh + _lib.MAP_TILE_H - 1) // _lib.MAP_TILE_H
        tx = (dw + _lib.MAP_TILE_W - 1) // _lib.MAP_TILE_W
        # 8 bytes per pixel (absolute tap, 16-bit fractions) or, compact, 4 (tap relative to the
        # tile's box, 8-bit fractions: coordinates to 1/510 pixel; tiles that fit no box have
        # no entries -- the kernel runs the exact chain there)
        self.compact = bool(compact)
        self._map_flags = _lib.FLAG_COMPACT_MAP if self.compact else 0
        self.map = torch.empty((b, dh, dw) if self.compact else (b, dh, dw, 2),
                               dtype=torch.int32, device=device)
        self.heads = torch.empty((b, ty, tx, 4), dtype=torch.int32, device=device)
        _ops.map_build(coords, self.map, self.heads, dtype, c, mats, tab_x, tab_y, None,
                       transform, b, sh, sw, dh, dw, flags=self._map_flags)

    @classmethod
    def equi2equi(cls, rots: Sequence[Dict[str, float]], src_shape, dtype, device,
                  height: Optional[int] = None, width: Optional[int] = None,
                  z_down: bool = False, compact: bool = False) -> "StaticRemap":
        """The map of ``equi2equi(src, rots, height=, width=, z_down=)``."""
        rots = _as_list(rots, src_shape)
        device = _device(device)
        dh = int(height) if height is not None else int(src_shape[2])
        dw = int(width) if width is not None else int(src_shape[3])
        mats = _prep.mats_for("rot", rots, z_down, device)
        tab_x, tab_y = _prep.equirect_tables(dh, dw, device)
        return cls(_lib.EQUI2EQUI, mats, tab_x, tab_y, src_shape, (dh, dw), dtype, device,
                   compact=compact)

    @classmethod
    def equi2pers(cls, rots: Sequence[Dict[str, float]], src_shape, dtype, device, height: int,
                  width: int, fov_x: float, skew: float = 0.0,
                  z_down: bool = False, compact: bool = False) -> "StaticRemap":
        """The map of ``equi2pers(equi, rots, height, width, fov_x, skew=, z_down=)``."""
        rots = _as_list(rots, src_shape)
        device = _device(device)
        mats = _prep.mats_for("equi2pers", rots, z_down, device,
                              (int(height), int(width), float(fov_x), float(skew)))
        return cls(_lib.EQUI2PERS, mats, None, None, src_shape, (height, width), dtype, device,
                   compact=compact)

    def __call__(self, src: torch.Tensor, out: Optional[torch.Tensor] = None,
                 saturate_uint8: bool = False) -> torch.Tensor:
        K.require_cuda_tensor(src, "src")
        if tuple(src.shape) != self.src_shape or src.dtype != self.dtype or \
                src.device != self.device:
            raise ValueError(
                f"StaticRemap was built for {self.src_shape} {self.dtype} on {self.device}, got "
                f"{tuple(src.shape)} {src.dtype} on {src.device}")
        b, c = self.src_shape[:2]
        flags = _lib.FLAG_SATURATE_U8 if (saturate_uint8 and src.dtype == torch.uint8) else 0
        flags |= self._map_flags
        # uint8 RGB / RGBX frames in torch.channels_last (the decoder's layout) are read and
        # written in place, like in equi2equi / equi2pers
        cl = src.stride(3) != 1 and K.is_channels_last_u8(src) and src.stride(2) % 16 == 0 \
            and src.data_ptr() % 16 == 0
        if cl:
            flags |= _lib.FLAG_CHANNELS_LAST
            if out is None:
                out = torch.empty((b, c) + self.dst_hw, dtype=src.dtype, device=src.device,
                                  memory_format=torch.channels_last)
        elif src.stride(3) != 1:
            src = src.contiguous()
        if out is None:
            out = torch.empty((b, c) + self.dst_hw, dtype=src.dtype, device=src.device)
        _ops.remap_mapped(src, out, self.map, self.heads, self._mats, self._tab_x, self._tab_y,
                          None, self.transform, flags)
        return out


def _device(device) -> torch.device:
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def _as_list(rots, src_shape) -> List[Dict[str, float]]:
    if isinstance(rots, dict):
        rots = [rots]
    rots = list(rots)
    if len(rots) != int(src_shape[0]):
        raise ValueError(f"{len(rots)} rotations for a batch of {int(src_shape[0])}")
    return rots
