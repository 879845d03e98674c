"""torch custom ops over the C ABI: ``equilib_b200::remap``, ``::minmax``, ``::coords``.

The ops only marshal pointers, sizes and strides into an ``eqb_remap_desc`` and
launch on ``torch.cuda.current_stream()`` of the tensor's device.  They never
allocate, never synchronise and have no CPU implementation.
"""

from __future__ import annotations

import ctypes as C
import threading
from typing import List, Optional

import torch
from torch import Tensor

from . import _lib

_tls = threading.local()  # per-thread cache of filled descriptors (_remap_impl)

_DTYPES = {
    torch.uint8: _lib.U8,
    torch.float16: _lib.F16,
    torch.bfloat16: _lib.BF16,
    torch.float32: _lib.F32,
}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise NotImplementedError(
            f"equilib_b200 supports uint8, float16, bfloat16 and float32 images, got {dt}"
        ) from None


def _raw_stream(index: int) -> int:
    """cudaStream_t of torch's current stream on device ``index`` (what
    ``torch.cuda.current_stream(dev).cuda_stream`` returns, without building a Stream object:
    0.3 us instead of 7 us per launch)."""
    return _get_raw_stream(index)


try:
    _get_raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:  # (a torch build without the private accessor)
    def _get_raw_stream(index: int) -> int:
        return torch.cuda.current_stream(torch.device("cuda", index)).cuda_stream


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(t: Tensor, name: str) -> None:
    if not t.is_cuda:
        raise _lib.EquilibB200Error(
            f"{name} must be a CUDA tensor: equilib_b200 has no CPU path (got {t.device})"
        )


def _f32_on(t: Optional[Tensor], dev: torch.device, name: str) -> None:
    if t is None:
        return
    if t.device != dev or not t.is_contiguous():
        raise _lib.EquilibB200Error(f"{name} must be contiguous and on {dev}")


def _fill_desc(src, out, mats, tab_x, tab_y, itab_x, grid, clip, transform, mode, flags, batch,
               channels, src_h, src_w, dst_h, dst_w, src_strides, dst_strides, grid_stride_b,
               w_face, n_cells, cell_offset, face_ptrs) -> _lib.RemapDesc:
    d = _lib.RemapDesc()
    d.transform, d.mode, d.flags = transform, mode, flags
    d.batch, d.channels = batch, channels
    d.src_h, d.src_w, d.dst_h, d.dst_w = src_h, src_w, dst_h, dst_w
    if src is not None:
        d.dtype = dtype_code(src.dtype)
        d.src = src.data_ptr()
        d.src_stride_b, d.src_stride_c, d.src_stride_h = src_strides
    if out is not None:
        d.dst = out.data_ptr()
        d.dst_stride_b, d.dst_stride_c, d.dst_stride_h = dst_strides
    d.mats, d.tab_x, d.tab_y = _ptr(mats), _ptr(tab_x), _ptr(tab_y)
    d.itab_x, d.grid, d.clip = _ptr(itab_x), _ptr(grid), _ptr(clip)
    d.grid_stride_b = grid_stride_b
    d.w_face, d.n_cells = w_face, n_cells
    for i, v in enumerate(cell_offset):
        d.cell_offset[i] = v
    d.face_ptrs = _ptr(face_ptrs)
    return d


def _remap_impl(
    src: Tensor,
    out: Tensor,
    mats: Optional[Tensor],
    tab_x: Optional[Tensor],
    tab_y: Optional[Tensor],
    itab_x: Optional[Tensor],
    grid: Optional[Tensor],
    clip: Optional[Tensor],
    face_ptrs: Optional[Tensor],
    transform: int,
    mode: int,
    flags: int,
    batch: int,
    channels: int,
    src_h: int,
    src_w: int,
    dst_h: int,
    dst_w: int,
    src_strides: List[int],
    dst_strides: List[int],
    grid_stride_b: int,
    w_face: int,
    n_cells: int,
    cell_offset: List[int],
) -> None:
    """One fused remap launch (eqb_remap).  ``out`` is written in place.

    Repeated calls with the same geometry (video frames, BASELINE config 1's 6 us kernel) reuse
    a filled descriptor and only refresh its pointers: the per-call host cost is what bounds
    small launches."""
    dev = src.device
    key = (transform, mode, flags, batch, channels, src_h, src_w, dst_h, dst_w,
           tuple(src_strides), tuple(dst_strides), grid_stride_b, w_face, n_cells,
           tuple(cell_offset), src.dtype, dev.index)
    cache = _tls.__dict__.setdefault("descs", {})
    d = cache.get(key)
    if d is None:
        _require_cuda(src, "src")
        _require_cuda(out, "out")
        d = _fill_desc(src, out, mats, tab_x, tab_y, itab_x, grid, clip, transform, mode, flags,
                       batch, channels, src_h, src_w, dst_h, dst_w, src_strides, dst_strides,
                       grid_stride_b, w_face, n_cells, cell_offset, face_ptrs)
        if len(cache) >= 64:
            cache.clear()
        cache[key] = d
    if not src.is_cuda or out.device != dev or out.dtype != src.dtype:
        raise _lib.EquilibB200Error("src and out must be CUDA tensors sharing device and dtype")
    for t, n in ((mats, "mats"), (tab_x, "tab_x"), (tab_y, "tab_y"), (itab_x, "itab_x"),
                 (grid, "grid"), (clip, "clip"), (face_ptrs, "face_ptrs")):
        if t is not None and (t.device != dev or not t.is_contiguous()):
            raise _lib.EquilibB200Error(f"{n} must be contiguous and on {dev}")
    d.src, d.dst = src.data_ptr(), out.data_ptr()
    d.mats, d.tab_x, d.tab_y = _ptr(mats), _ptr(tab_x), _ptr(tab_y)
    d.itab_x, d.grid, d.clip = _ptr(itab_x), _ptr(grid), _ptr(clip)
    d.face_ptrs = _ptr(face_ptrs)
    if torch.cuda.current_device() == dev.index:
        status = _lib.lib().eqb_remap(C.byref(d), _raw_stream(dev.index))
    else:
        with torch.cuda.device(dev):
            status = _lib.lib().eqb_remap(C.byref(d), _raw_stream(dev.index))
    if status != 0:
        _lib.check(status, "eqb_remap")


def supports_channels_last(src, out, mats, tab_x, tab_y, itab_x, grid, clip, face_ptrs,
                           transform, mode, flags, batch, channels, src_h, src_w, dst_h, dst_w,
                           src_strides, dst_strides, grid_stride_b, w_face, n_cells,
                           cell_offset) -> bool:
    """eqb_remap_supports_channels_last for the arguments of ``remap`` (host only)."""
    d = _fill_desc(src, out, mats, tab_x, tab_y, itab_x, grid, clip, transform, mode, flags,
                   batch, channels, src_h, src_w, dst_h, dst_w, src_strides, dst_strides,
                   grid_stride_b, w_face, n_cells, cell_offset, face_ptrs)
    return bool(_lib.lib().eqb_remap_supports_channels_last(C.byref(d)))


# The registered op (torch.ops.equilib_b200.remap) is what graph capture / torch.compile
# see; eager callers go straight to the implementation and skip ~50 us of dispatcher
# overhead per call, which dominates small launches (BASELINE config 1 is a 5 us kernel).
remap_op = torch.library.custom_op("equilib_b200::remap", _remap_impl, mutates_args=("out",),
                                   device_types="cuda")


def remap(*args) -> None:
    if torch.compiler.is_compiling():
        remap_op(*args)
    else:
        _remap_impl(*args)


@torch.library.custom_op("equilib_b200::minmax", mutates_args=("out2",), device_types="cuda")
def minmax(img: Tensor, out2: Tensor) -> None:
    """Global min/max of a (B, C, H, W) image batch into a 2-float device tensor."""
    _require_cuda(img, "img")
    if img.dim() != 4 or img.stride(3) != 1:
        raise _lib.EquilibB200Error("minmax expects a 4-d tensor with unit inner stride")
    if out2.device != img.device or out2.dtype != torch.float32 or out2.numel() != 2:
        raise _lib.EquilibB200Error("out2 must be 2 float32 values on the image's device")
    im = _lib.Image()
    im.data = img.data_ptr()
    im.dtype = dtype_code(img.dtype)
    im.batch, im.channels, im.height, im.width = img.shape
    im.stride_b, im.stride_c, im.stride_h = img.stride(0), img.stride(1), img.stride(2)
    with torch.cuda.device(img.device):
        stream = torch.cuda.current_stream(img.device).cuda_stream
        _lib.check(
            _lib.lib().eqb_minmax(C.byref(im), C.c_void_p(out2.data_ptr()), C.c_void_p(stream)),
            "eqb_minmax",
        )


def coords(coords_out: Tensor, mats, tab_x, tab_y, itab_x, grid, transform: int, batch: int,
           src_h: int, src_w: int, dst_h: int, dst_w: int, grid_stride_b: int = 0, w_face: int = 0,
           n_cells: int = 0, flags: int = 0) -> None:
    """eqb_remap_coords: the transform's sampling grid (B, 2, dst_h, dst_w), for tests."""
    _require_cuda(coords_out, "coords_out")
    d = _fill_desc(None, None, mats, tab_x, tab_y, itab_x, grid, None, transform, 0, flags, batch, 1,
                   src_h, src_w, dst_h, dst_w, (0, 0, 0), (0, 0, 0), grid_stride_b, w_face,
                   n_cells, [0] * 12, None)
    with torch.cuda.device(coords_out.device):
        stream = torch.cuda.current_stream(coords_out.device).cuda_stream
        _lib.check(
            _lib.lib().eqb_remap_coords(C.byref(d), C.c_void_p(coords_out.data_ptr()),
                                        C.c_void_p(stream)),
            "eqb_remap_coords",
        )


def map_build(coords_t: Tensor, map_t: Tensor, heads_t: Tensor, dtype: torch.dtype, channels: int,
              mats, tab_x, tab_y, itab_x, transform: int, batch: int, src_h: int, src_w: int,
              dst_h: int, dst_w: int, flags: int = 0) -> None:
    """eqb_map_build: pack the coordinates of a fixed set of rotations into a sampling map."""
    _require_cuda(coords_t, "coords")
    d = _fill_desc(None, None, mats, tab_x, tab_y, itab_x, None, None, transform, 1, flags, batch,
                   channels, src_h, src_w, dst_h, dst_w, (0, 0, 0), (0, 0, 0), 0, 0, 0,
                   [0] * 12, None)
    d.dtype = dtype_code(dtype)
    with torch.cuda.device(coords_t.device):
        stream = torch.cuda.current_stream(coords_t.device).cuda_stream
        _lib.check(
            _lib.lib().eqb_map_build(C.byref(d), C.c_void_p(coords_t.data_ptr()),
                                     C.c_void_p(map_t.data_ptr()),
                                     C.c_void_p(heads_t.data_ptr()), C.c_void_p(stream)),
            "eqb_map_build",
        )


def remap_mapped(src: Tensor, out: Tensor, map_t: Tensor, heads_t: Tensor, mats, tab_x, tab_y,
                 itab_x, transform: int, flags: int) -> None:
    """eqb_remap_mapped: ``out`` <- ``src`` sampled through a map from ``map_build``."""
    _require_cuda(src, "src")
    _require_cuda(out, "out")
    if out.device != src.device or out.dtype != src.dtype:
        raise _lib.EquilibB200Error("src and out must share device and dtype")
    b, c, sh, sw = src.shape
    d = _fill_desc(src, out, mats, tab_x, tab_y, itab_x, None, None, transform, 1, flags, b, c,
                   sh, sw, out.shape[2], out.shape[3], src.stride()[:3], out.stride()[:3], 0, 0,
                   0, [0] * 12, None)
    with torch.cuda.device(src.device):
        stream = torch.cuda.current_stream(src.device).cuda_stream
        _lib.check(
            _lib.lib().eqb_remap_mapped(C.byref(d), C.c_void_p(map_t.data_ptr()),
                                        C.c_void_p(heads_t.data_ptr()), C.c_void_p(stream)),
            "eqb_remap_mapped",
        )


def cube_chain(equi: Tensor, out: Tensor, mats: Tensor, rng: Tensor, tab_x: Tensor, tab_y: Tensor,
               itab_x: Tensor, mode: int, flags: int, w_face: int) -> None:
    """eqb_equi2cube2equi: ``out`` <- cube2equi(equi2cube(``equi``)) in one launch; the two
    descriptors are exactly those of the two eqb_remap launches it replaces."""
    _require_cuda(equi, "equi")
    _require_cuda(out, "out")
    if out.device != equi.device or out.dtype != equi.dtype:
        raise _lib.EquilibB200Error("equi and out must share device and dtype")
    dev = equi.device
    for t, n in ((mats, "mats"), (rng, "rng"), (tab_x, "tab_x"), (tab_y, "tab_y"),
                 (itab_x, "itab_x")):
        if t.device != dev or not t.is_contiguous():
            raise _lib.EquilibB200Error(f"{n} must be contiguous and on {dev}")
    b, c, h, w = equi.shape
    cells = [f * w_face for f in range(6)] + [0] * 6
    to_cube = _fill_desc(equi, None, mats, rng, None, None, None, None, _lib.EQUI2CUBE, mode,
                         flags, b, c, h, w, w_face, 6 * w_face, equi.stride()[:3], (0, 0, 0), 0,
                         w_face, 6, cells, None)
    to_equi = _fill_desc(None, out, None, tab_x, tab_y, itab_x, None, None, _lib.CUBE2EQUI, mode,
                         flags, b, c, w_face, 6 * w_face, out.shape[2], out.shape[3], (0, 0, 0),
                         out.stride()[:3], 0, w_face, 6, cells, None)
    to_equi.dtype = to_cube.dtype
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(
            _lib.lib().eqb_equi2cube2equi(C.byref(to_cube), C.byref(to_equi), C.c_void_p(stream)),
            "eqb_equi2cube2equi",
        )
