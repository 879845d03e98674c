"""ctypes binding of libequilib_b200.so (C ABI in include/equilib_b200.h).

There is no fallback: if the shared library is missing or its ABI version does not
match, importing the operators raises.  Build it with ``python -m equilib_b200.build``
(or ``__graft_entry__.build()``).
"""

from __future__ import annotations

import ctypes as C
import os
import threading

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libequilib_b200.so")

ABI_VERSION = 1

# enums (include/equilib_b200.h)
EQUI2EQUI, EQUI2PERS, EQUI2CUBE, CUBE2EQUI, PERS2EQUI, GRID = range(6)
NEAREST, BILINEAR, BICUBIC = range(3)
U8, F16, BF16, F32 = range(4)
FLAG_SATURATE_U8 = 1
FLAG_LIBM = 4
FLAG_NO_STAGE = 8
FLAG_EXACT_COORDS = 16
FLAG_FAST_COORDS = 32
FLAG_NO_TMA = 64
FLAG_FORCE_STAGE = 128
FLAG_NO_LEAN = 1024
FLAG_CHANNELS_LAST = 2048
FLAG_NO_LEAN2 = 4096
FLAG_FORCE_LEAN2 = 8192
FLAG_COMPACT_MAP = 32768
MAP_TILE_W, MAP_TILE_H = 64, 8
FLAG_WARP_SHAPE_SHIFT = 8

MODES = {"nearest": NEAREST, "bilinear": BILINEAR, "bicubic": BICUBIC}


class RemapDesc(C.Structure):
    _fields_ = [
        ("transform", C.c_int32),
        ("mode", C.c_int32),
        ("dtype", C.c_int32),
        ("flags", C.c_uint32),
        ("batch", C.c_int32),
        ("channels", C.c_int32),
        ("src_h", C.c_int32),
        ("src_w", C.c_int32),
        ("dst_h", C.c_int32),
        ("dst_w", C.c_int32),
        ("src", C.c_void_p),
        ("src_stride_b", C.c_int64),
        ("src_stride_c", C.c_int64),
        ("src_stride_h", C.c_int64),
        ("dst", C.c_void_p),
        ("dst_stride_b", C.c_int64),
        ("dst_stride_c", C.c_int64),
        ("dst_stride_h", C.c_int64),
        ("mats", C.c_void_p),
        ("tab_x", C.c_void_p),
        ("tab_y", C.c_void_p),
        ("itab_x", C.c_void_p),
        ("grid", C.c_void_p),
        ("grid_stride_b", C.c_int64),
        ("clip", C.c_void_p),
        ("w_face", C.c_int32),
        ("n_cells", C.c_int32),
        ("cell_offset", C.c_int64 * 12),
        ("face_ptrs", C.c_void_p),
    ]


class Image(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("dtype", C.c_int32),
        ("batch", C.c_int32),
        ("channels", C.c_int32),
        ("height", C.c_int32),
        ("width", C.c_int32),
        ("stride_b", C.c_int64),
        ("stride_c", C.c_int64),
        ("stride_h", C.c_int64),
    ]


EXPORTS = {
    "eqb_abi_version": (C.c_int32, []),
    "eqb_last_error": (C.c_char_p, []),
    "eqb_remap": (C.c_int32, [C.POINTER(RemapDesc), C.c_void_p]),
    "eqb_remap_supports_channels_last": (C.c_int32, [C.POINTER(RemapDesc)]),
    "eqb_equi2cube2equi": (C.c_int32, [C.POINTER(RemapDesc), C.POINTER(RemapDesc), C.c_void_p]),
    "eqb_remap_coords": (C.c_int32, [C.POINTER(RemapDesc), C.c_void_p, C.c_void_p]),
    "eqb_map_build": (C.c_int32, [C.POINTER(RemapDesc), C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "eqb_remap_mapped": (C.c_int32, [C.POINTER(RemapDesc), C.c_void_p, C.c_void_p, C.c_void_p]),
    "eqb_minmax": (C.c_int32, [C.POINTER(Image), C.c_void_p, C.c_void_p]),
    "eqb_rotation_compose": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p]),
    "eqb_launch_count": (C.c_int64, []),
    "eqb_source_hash": (C.c_char_p, []),
}

_lock = threading.Lock()
_lib = None


class EquilibB200Error(RuntimeError):
    pass


def lib() -> C.CDLL:
    """The loaded library; loads it on first use.  Raises if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise EquilibB200Error(
                f"{LIB_PATH} is missing: build it with `python -m equilib_b200.build`. "
                "equilib_b200 has no CPU or PyTorch fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)  # AttributeError if a symbol is missing
            fn.restype = res
            fn.argtypes = args
        got = handle.eqb_abi_version()
        if got != ABI_VERSION:
            raise EquilibB200Error(f"ABI version mismatch: library {got}, binding {ABI_VERSION}")
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().eqb_last_error().decode("utf-8", "replace")
        raise EquilibB200Error(f"{what} failed (status {status}): {msg}")


def source_hash() -> str:
    """Source hash compiled into the loaded library (equilib_b200/build.py::source_hash)."""
    return lib().eqb_source_hash().decode("ascii", "replace")


def launch_count() -> int:
    return int(lib().eqb_launch_count())
