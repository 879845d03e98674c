"""Host-side preparation for the fused kernels: 3x3 matrices and 1-D tables.

Everything a kernel needs besides the image is tiny: ``B x 9`` floats per call and
a few 1-D tables per output size.  They are built on the host with the same torch
CPU operations the reference uses for its full-size grids (so the values are
bit-identical to the reference's), uploaded once and cached per device.

Reference code this replaces / follows:
  rotation matrices   equilib/torch_utils/rotation.py:27-78,134-154
  intrinsics, G       equilib/torch_utils/intrinsic.py:5-39,
                      equilib/equi2pers/torch.py:21-40, pers2equi/torch.py:19-38
  unit-sphere rays    equilib/torch_utils/grid.py:92-101  (as 1-D tables)
  cube-face rays      equilib/torch_utils/grid.py:127-132 (the ``rng`` vector)
  cube2equi grid      equilib/cube2equi/torch.py:135-245  (as 1-D tables)
  grid-prep cache     equilib/_cache.py:21-40 (bounded FIFO, here per device)
"""

from __future__ import annotations

import ctypes as C
import math
import threading
from collections import OrderedDict
from typing import Dict, Hashable, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib

F32 = torch.float32
_PI = torch.tensor([3.14159265358979323846], dtype=F32)  # intrinsic.py:5

TABLE_CACHE_MAXSIZE = 16  # same bound as equilib/_cache.py:18
MATS_CACHE_MAXSIZE = 64


def _tensors(value):
    if torch.is_tensor(value):
        yield value
    elif isinstance(value, (tuple, list)):
        for v in value:
            yield from _tensors(v)


class _BoundedCache:
    """Thread-safe FIFO-bounded dict (the reference's cached_grid policy).

    Entries are device tensors that launches on ANY stream read.  Two rules keep that safe
    without per-use events: (1) an entry is published only after its upload has completed
    (one host synchronisation of the uploading stream, once per entry); (2) an evicted
    entry is parked in a second FIFO of the same size instead of being freed, so its memory
    cannot be handed out again while a launch enqueued before the eviction may still read
    it (a block would have to survive 2 x maxsize newer entries' worth of calls)."""

    def __init__(self, maxsize: int):
        self.maxsize = maxsize
        self._d: "OrderedDict[Hashable, object]" = OrderedDict()
        self._parked: List[object] = []
        self._lock = threading.Lock()

    def get(self, key, build):
        with self._lock:
            if key in self._d:
                return self._d[key]
        value = build()
        for t in _tensors(value):
            if t.is_cuda:
                torch.cuda.current_stream(t.device).synchronize()
                break
        with self._lock:
            if key not in self._d:
                if len(self._d) >= self.maxsize:
                    self._parked.append(self._d.popitem(last=False)[1])
                    del self._parked[:-self.maxsize]
                self._d[key] = value
            return self._d[key]

    def __len__(self):
        return len(self._d)

    def clear(self):
        with self._lock:
            self._d.clear()
            self._parked.clear()


_tables = _BoundedCache(TABLE_CACHE_MAXSIZE)
_mats = _BoundedCache(MATS_CACHE_MAXSIZE)


def clear_caches() -> None:
    _tables.clear()
    _mats.clear()


# --------------------------------------------------------------------------
# rotation matrices
# --------------------------------------------------------------------------


def rotation_matrices_host(rots: Sequence[Dict[str, float]], z_down: bool) -> torch.Tensor:
    """(B, 3, 3) fp32 on the host, bit-identical to create_rotation_matrices.

    The reference builds each matrix from float64 numpy trig rounded to fp32 and two
    fp32 2-D matmuls (an FMA chain) inside a per-rotation python loop (~30 us each).
    Here the trig is one vectorised numpy call and the products are one C call.
    """
    n = len(rots)
    ang = np.empty((n, 3), dtype=np.float64)
    for i, r in enumerate(rots):
        ang[i, 0] = r["roll"]
        ang[i, 1] = r["pitch"]
        ang[i, 2] = r["yaw"]
    if not z_down:
        ang[:, 1:] = -ang[:, 1:]
    trig = np.empty((n, 6), dtype=np.float32)
    trig[:, 0::2] = np.cos(ang)
    trig[:, 1::2] = np.sin(ang)
    out = np.empty((n, 9), dtype=np.float32)
    _lib.check(
        _lib.lib().eqb_rotation_compose(
            trig.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p)
        ),
        "eqb_rotation_compose",
    )
    return torch.from_numpy(out).view(n, 3, 3)


def _g2c() -> torch.Tensor:
    r_xy = torch.tensor([[0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]], dtype=F32)
    r_yz = torch.tensor([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, 1.0, 0.0]], dtype=F32)
    return r_xy @ r_yz


def intrinsic_matrix(height: int, width: int, fov_x: float, skew: float) -> torch.Tensor:
    rad = torch.tensor(fov_x, dtype=F32) * _PI / 180.0
    f = width / (2 * torch.tan(rad / 2))
    return torch.tensor(
        [[f, skew, width / 2], [0.0, f, height / 2], [0.0, 0.0, 1.0]], dtype=F32
    )


def _rots_key(rots: Sequence[Dict[str, float]]) -> Tuple:
    return tuple((float(r["roll"]), float(r["pitch"]), float(r["yaw"])) for r in rots)


def mats_for(kind: str, rots: Sequence[Dict[str, float]], z_down: bool, device: torch.device,
             extra: Tuple = ()) -> torch.Tensor:
    """Device tensor (B, 9) fp32 of the per-item matrix A, cached per rots batch.

    kind "rot":       A = R                      (equi2equi, equi2cube)
    kind "equi2pers": A = R @ (g2c @ K^-1)       extra = (height, width, fov_x, skew)
    kind "pers2equi": A = (K @ g2c^T) @ R        extra = (h_pers, w_pers, fov_x, skew)
    """
    key = (kind, _rots_key(rots), bool(z_down), extra, device.type, device.index)

    def build():
        R = rotation_matrices_host(rots, z_down)
        if kind == "rot":
            A = R
        elif kind == "equi2pers":
            h, w, fov_x, skew = extra
            G = _g2c() @ intrinsic_matrix(h, w, fov_x, skew).inverse()
            A = torch.matmul(R, G)
        elif kind == "pers2equi":
            h, w, fov_x, skew = extra
            G = intrinsic_matrix(h, w, fov_x, skew) @ _g2c().T
            A = torch.matmul(G, R)
        else:
            raise ValueError(kind)
        return A.reshape(len(rots), 9).contiguous().to(device)

    return _mats.get(key, build)


# --------------------------------------------------------------------------
# 1-D tables
# --------------------------------------------------------------------------


def equirect_tables(height: int, width: int, device: torch.device):
    """(tab_x, tab_y) for an equirect *output*: W interleaved (cos(theta), sin(theta))
    pairs and [cos(phi)|sin(phi)] (2H), theta = x*2*pi/W - pi, phi = y*pi/H - pi/2."""
    key = ("equirect", height, width, device.type, device.index)

    def build():
        xs = torch.linspace(0, width - 1, width, dtype=F32)
        ys = torch.linspace(0, height - 1, height, dtype=F32)
        theta = xs * 2 * _PI / width - _PI
        phi = ys * _PI / height - _PI / 2
        tx = torch.stack([torch.cos(theta), torch.sin(theta)], dim=1).reshape(-1).contiguous()
        ty = torch.cat([1 * torch.cos(phi), 1 * torch.sin(phi)])
        return tx.to(device), ty.to(device)

    return _tables.get(key, build)


def cube_rng_table(w_face: int, device: torch.device) -> torch.Tensor:
    key = ("rng", w_face, device.type, device.index)

    def build():
        ratio = (w_face - 1) / w_face
        return torch.linspace(-0.5 * ratio, 0.5 * ratio, w_face, dtype=F32).to(device)

    return _tables.get(key, build)


def cube2equi_tables(height: int, width: int, device: torch.device):
    """(tab_x [4W], tab_y [3H], itab_x [2W]) -- see include/equilib_b200.h."""
    key = ("cube2equi", height, width, device.type, device.index)

    def build():
        w_ratio = (width - 1) / width
        h_ratio = (height - 1) / height
        theta = torch.linspace(-(math.pi * w_ratio), math.pi * w_ratio, steps=width, dtype=F32)
        phi = torch.linspace((math.pi * h_ratio) / 2, -(math.pi * h_ratio) / 2, steps=height,
                             dtype=F32)
        x = torch.arange(width)
        face = ((x - (3 * width // 8)) % width) // (width // 4)
        idx = torch.linspace(-(math.pi * w_ratio), math.pi * w_ratio, width // 4) / 4
        idx = height // 2 - torch.round(
            torch.atan(torch.cos(idx)) * height / (math.pi * h_ratio)
        )
        thr = torch.roll(torch.cat([idx.type(torch.int64)] * 4), 3 * width // 8)
        shift = torch.zeros(width, dtype=F32)
        for i in range(4):
            m = face == i
            shift[m] = theta[m] - math.pi * i / 2
        tx = torch.cat([0.5 * torch.tan(shift), torch.cos(shift), torch.sin(theta),
                        torch.cos(theta)])
        ty = torch.cat([-0.5 * torch.tan(phi), 0.5 * torch.tan(math.pi / 2 - phi),
                        0.5 * torch.tan(math.pi / 2 - torch.abs(phi))])
        itx = torch.cat([face, thr]).to(torch.int32)
        return tx.to(device), ty.to(device), itx.to(device)

    return _tables.get(key, build)
