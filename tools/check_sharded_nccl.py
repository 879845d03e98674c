#!/usr/bin/env python3
"""Two (or more) GPUs, NCCL: the sharded entry points against the single-GPU result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29531 tools/check_sharded_nccl.py

Checks (rank 0 prints one JSON line):
  * scatter_run_gather(equi2equi bicubic, clip_output=True): NCCL point-to-point scatter of
    the frames, local remap with the clip range all-reduced, gather -> equals the
    single-device call bit for bit;
  * run_presharded: per-rank shards with clip_group -> equal the slices of that result.
"""
import json
import math
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import equilib_b200  # noqa: E402
from equilib_b200 import sharded  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    b = 2 * world + 1  # uneven split
    g = torch.Generator().manual_seed(7)
    img = torch.rand((b, 3, 256, 512), generator=g)
    # frames with different ranges: per-shard min/max would differ from the global one
    img = img * torch.linspace(0.2, 1.0, b).view(b, 1, 1, 1)
    rots = [{"roll": 0.1 * i, "pitch": 0.3 * math.sin(i), "yaw": 0.7 * i} for i in range(b)]
    fn = equilib_b200.equi2equi
    full = fn(img.to(dev), rots, mode="bicubic")  # every rank computes the reference result
    out = sharded.scatter_run_gather(fn, img.to(dev) if rank == 0 else None,
                                     rots if rank == 0 else None, mode="bicubic")
    lo, hi = sharded.local_range(b, world, rank)
    mine = sharded.run_presharded(fn, img[lo:hi].to(dev), rots[lo:hi], mode="bicubic")
    ok_local = torch.equal(mine, full[lo:hi])
    flags = torch.tensor([int(ok_local)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "batch": b,
                          "scatter_run_gather_equal": bool(torch.equal(out, full)),
                          "presharded_equal_all_ranks": bool(flags.item())}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
