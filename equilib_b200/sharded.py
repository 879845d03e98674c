"""Multi-GPU use of the remap path: one process per GPU, frames sharded, no
inter-GPU traffic on the hot path.

Every batch item (frame, or view) of every transform is independent
(``len(img) == len(rots)``, e.g. equilib/equi2pers/torch.py:131-133), so the batch is
split contiguously across ranks and each rank runs the ordinary single-GPU call on
its shard.  The one cross-item quantity in the reference is the *global* min/max
behind ``clip_output`` (e.g. equilib/equi2equi/torch.py:197): pass
``clip_group=<process group>`` to any transform and the two scalars are all-reduced
(8 bytes) so sharded results equal the single-device result exactly.

``scatter_run_gather`` is the convenience mode for callers who hold the whole batch
on one device and want the whole result back there: NCCL point-to-point over
NVLink moves 1/world of the input to every rank and the outputs back.  It cannot
scale (the root's NVLink port carries (world-1)/world of all bytes in both
directions, several times slower than running the batch locally from HBM); it exists
for functionality, and bench.py reports scaling for the pre-sharded case only.

The reference has no counterpart: it contains no multi-device code at all
("Parallelism strategies: none").
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def partition(n_items: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) ranges, sizes differing by at most one, in rank order."""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    base, rem = divmod(n_items, world_size)
    out, start = [], 0
    for r in range(world_size):
        size = base + (1 if r < rem else 0)
        out.append((start, start + size))
        start += size
    return out


def local_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    return partition(n_items, world_size)[rank]


def allreduce_minmax(lo_hi: torch.Tensor, group=None) -> torch.Tensor:
    """In-place global (min, max) of per-rank (min, max) pairs with ONE all-reduce:
    (-lo, hi) reduced with MAX."""
    if not (dist.is_available() and dist.is_initialized()):
        return lo_hi
    t = torch.stack([-lo_hi[0], lo_hi[1]])
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    lo_hi[0] = -t[0]
    lo_hi[1] = t[1]
    return lo_hi


def run_presharded(fn: Callable, img_local, rots_local, group=None, **kwargs):
    """The scaling path: this rank's shard in, this rank's shard out.  Only the clip
    range (when one is needed) crosses GPUs."""
    return fn(img_local, rots_local, clip_group=group if group is not None else True, **kwargs)


def scatter_run_gather(fn: Callable, img: Optional[torch.Tensor], rots: Optional[Sequence],
                       *, root: int = 0, group=None, device: Optional[torch.device] = None,
                       **kwargs) -> Optional[torch.Tensor]:
    """``fn(img, rots, **kwargs)`` for a 4-d batch held by ``root``, computed by all
    ranks of ``group``.  Returns the full result on ``root`` and ``None`` elsewhere.
    Non-root ranks pass ``img=None, rots=None``."""
    # `rank`, `root` and the partition index are GROUP ranks; point-to-point calls and
    # broadcast_object_list address processes by GLOBAL rank
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)

    def g(r: int) -> int:
        return dist.get_global_rank(group, r) if group is not None else r

    meta = [None]
    if rank == root:
        assert img is not None and img.dim() == 4 and rots is not None
        assert len(img) == len(rots)
        meta = [(tuple(img.shape), img.dtype, list(rots))]
    dist.broadcast_object_list(meta, src=g(root), group=group)
    shape, dtype, all_rots = meta[0]
    if shape[0] < world:
        raise ValueError(f"batch {shape[0]} is smaller than the group ({world} ranks)")
    parts = partition(shape[0], world)
    lo, hi = parts[rank]
    dev = device if device is not None else (img.device if img is not None else None)
    if dev is None:
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
            else torch.device("cpu")

    # scatter: root -> every rank, one contiguous slice each
    if rank == root:
        reqs = []
        for r, (a, b) in enumerate(parts):
            if r != root and b > a:
                reqs.append(dist.isend(img[a:b].contiguous(), dst=g(r), group=group))
        local = img[lo:hi]
        for q in reqs:
            q.wait()
    else:
        local = torch.empty((hi - lo,) + tuple(shape[1:]), dtype=dtype, device=dev)
        if hi > lo:
            dist.recv(local, src=g(root), group=group)

    out_local = fn(local, all_rots[lo:hi], clip_group=group if group is not None else True,
                   **kwargs)

    # gather: every rank -> root
    out_meta = [None] * world
    dist.all_gather_object(out_meta, None if out_local is None else
                           (tuple(out_local.shape), out_local.dtype), group=group)
    if rank == root:
        ref = next(m for m in out_meta if m is not None)
        full = torch.empty((shape[0],) + tuple(ref[0][1:]), dtype=ref[1], device=dev)
        for r, (a, b) in enumerate(parts):
            if b <= a:
                continue
            if r == root:
                full[a:b] = out_local
            else:
                dist.recv(full[a:b], src=g(r), group=group)
        return full
    if out_local is not None:
        dist.send(out_local.contiguous(), dst=g(root), group=group)
    return None
