"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo groups.

The product has no CPU kernels, so the per-shard transform is replaced by a stand-in
with the same call shape ``fn(img, rots, clip_group=..., **kw)``; what is tested is the
partitioning, the scatter / gather plumbing, item order, and the clip-range
all-reduce that makes sharded results equal single-device results."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from equilib_b200 import sharded


def test_partition_is_contiguous_balanced_and_complete():
    for n in (1, 2, 7, 8, 9, 32, 256):
        for w in (1, 2, 3, 4, 8):
            parts = sharded.partition(n, w)
            assert len(parts) == w
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
            assert sharded.local_range(n, w, w - 1) == parts[-1]
    with pytest.raises(ValueError):
        sharded.partition(4, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_transform(img, rots, clip_group=None, clip_output=True, scale=1.0):
    """Stand-in for a remap: per-item scaling by the item's yaw, then the reference's
    global clamp to [min(img), max(img)] with the range all-reduced over the group."""
    yaw = torch.tensor([r["yaw"] for r in rots], dtype=img.dtype).view(-1, 1, 1, 1)
    out = img * yaw * scale
    if clip_output:
        lo_hi = torch.stack([img.min(), img.max()]).to(torch.float32)
        lo_hi = sharded.allreduce_minmax(lo_hi, None if clip_group is True else clip_group)
        out = out.clamp(lo_hi[0].to(img.dtype), lo_hi[1].to(img.dtype))
    return out


def _worker(rank, world, port, n_items, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full = torch.rand((n_items, 3, 4, 5), generator=g) * 4 - 1
        rots = [{"roll": 0.0, "pitch": 0.0, "yaw": 0.25 * (i + 1)} for i in range(n_items)]
        want = _fake_transform(full, rots, clip_group=None)  # no group: single device

        # pre-sharded path: each rank owns its slice, only the clip range is shared
        a, b = sharded.local_range(n_items, world, rank)
        mine = sharded.run_presharded(_fake_transform, full[a:b], rots[a:b], scale=1.0)
        assert torch.equal(mine, want[a:b]), "pre-sharded result differs from single-device"

        # scatter / gather path: root holds everything, result lands on root in order
        got = sharded.scatter_run_gather(
            _fake_transform, full if rank == 0 else None, rots if rank == 0 else None,
            root=0, device=torch.device("cpu"))
        if rank == 0:
            assert got is not None and torch.equal(got, want)
            torch.save(got, result_path)
        else:
            assert got is None

        # one all-reduce gives the global (min, max)
        lo_hi = torch.tensor([float(rank), float(rank) + 10.0])
        sharded.allreduce_minmax(lo_hi)
        assert lo_hi.tolist() == [0.0, float(world - 1) + 10.0]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [2, 5])
def test_world_size_two_gloo(tmp_path, n_items):
    port = _free_port()
    path = str(tmp_path / "out.pt")
    mp.spawn(_worker, args=(2, port, n_items, path), nprocs=2, join=True)
    assert os.path.exists(path)


def _worker_small_batch(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        img = torch.zeros((1, 3, 4, 5)) if rank == 0 else None
        rots = [{"roll": 0.0, "pitch": 0.0, "yaw": 1.0}] if rank == 0 else None
        try:
            sharded.scatter_run_gather(_fake_transform, img, rots, device=torch.device("cpu"))
        except ValueError:
            return
        raise AssertionError("a batch smaller than the group must be rejected")
    finally:
        dist.destroy_process_group()


def test_batch_smaller_than_group_is_rejected():
    mp.spawn(_worker_small_batch, args=(2, _free_port()), nprocs=2, join=True)


def _worker_subgroup(rank, world, port):
    """A strict sub-group {1, 2} of a world of 3: group rank 0 is global rank 1, so any
    point-to-point call that confuses the two addresses the wrong process."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sub = dist.new_group(ranks=[1, 2])
        if rank == 0:
            return
        g = torch.Generator().manual_seed(1)
        full = torch.rand((5, 3, 4, 5), generator=g) * 4 - 1
        rots = [{"roll": 0.0, "pitch": 0.0, "yaw": 0.25 * (i + 1)} for i in range(5)]
        yaw = torch.tensor([r["yaw"] for r in rots]).view(-1, 1, 1, 1)
        want = (full * yaw).clamp(full.min(), full.max())  # single-device result
        is_root = dist.get_rank(sub) == 0  # global rank 1
        got = sharded.scatter_run_gather(
            _fake_transform, full if is_root else None, rots if is_root else None,
            root=0, group=sub, device=torch.device("cpu"))
        if is_root:
            assert got is not None and torch.equal(got, want)
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


def test_scatter_gather_over_a_strict_subgroup():
    mp.spawn(_worker_subgroup, args=(3, _free_port()), nprocs=3, join=True)
