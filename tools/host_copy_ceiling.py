#!/usr/bin/env python3
"""The box's pinned host <-> device copy ceiling with N processes copying at once.

    python tools/host_copy_ceiling.py --procs 1 2 4 8 [--mb 1600] [--out gpurun_out/host_copy_ceiling.json]

Process r binds to GPU r, pins two buffers of --mb megabytes and runs, on two streams,
plain cudaMemcpyAsync H2D and D2H copies of 64 MB pieces concurrently for a few seconds
(one plain async copy per piece, no batched-copy API).  All processes start together (a
file barrier); each reports its own GB/s per direction; the parent prints per-N aggregate
and per-process figures.  This is what bench.py's `e2e` number is bounded by: the remap
kernel itself moves a 4K frame 20x faster than PCIe does.

Copy profiles/host_copy_ceiling.json from the output to let bench.py print the ceiling
beside `e2e`.
"""
import argparse
import json
import os
import subprocess
import sys
import time


def worker(rank, procs, mb, seconds, sync_dir, bind):
    import torch

    if bind:
        # distinct cores per rank before pinning (first-touch places the pinned pages)
        n = os.cpu_count() or 1
        per = max(1, n // procs)
        try:
            os.sched_setaffinity(0, set(range(rank * per, min(n, (rank + 1) * per))))
        except OSError:
            pass
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    nbytes = mb << 20
    hin = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    hout = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    hin.fill_(1)
    hout.fill_(0)
    din = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dout = torch.ones(nbytes, dtype=torch.uint8, device=dev)
    piece = 64 << 20
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    open(os.path.join(sync_dir, f"ready{rank}"), "w").close()
    while len([f for f in os.listdir(sync_dir) if f.startswith("ready")]) < procs:
        time.sleep(0.001)

    res = {}
    for mode in ("both", "h2d", "d2h"):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        moved = 0
        t_end = time.perf_counter() + seconds
        e[0].record(s_in)
        e[2].record(s_out)
        while time.perf_counter() < t_end:
            for off in range(0, nbytes, piece):
                if mode != "d2h":
                    with torch.cuda.stream(s_in):
                        din[off:off + piece].copy_(hin[off:off + piece], non_blocking=True)
                if mode != "h2d":
                    with torch.cuda.stream(s_out):
                        hout[off:off + piece].copy_(dout[off:off + piece], non_blocking=True)
            moved += nbytes
            s_in.synchronize()
            s_out.synchronize()
        e[1].record(s_in)
        e[3].record(s_out)
        torch.cuda.synchronize()
        ms = max(e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3]))
        res[mode] = moved / (ms * 1e-3) / 1e9  # GB/s per direction
    print(json.dumps({"rank": rank, **res}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, nargs="+", default=[1])
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--out", default="")
    ap.add_argument("--bind", action="store_true")
    ap.add_argument("--worker", type=int, default=-1)
    ap.add_argument("--nprocs", type=int, default=1)
    ap.add_argument("--sync-dir", default="")
    a = ap.parse_args()
    if a.worker >= 0:
        return worker(a.worker, a.nprocs, a.mb, a.seconds, a.sync_dir, a.bind)
    import tempfile

    table = {}
    for n in a.procs:
        for bind in (False, True):
            with tempfile.TemporaryDirectory() as d:
                ps = [subprocess.Popen([sys.executable, __file__, "--worker", str(r), "--nprocs",
                                        str(n), "--mb", str(a.mb), "--seconds", str(a.seconds),
                                        "--sync-dir", d] + (["--bind"] if bind else []),
                                       stdout=subprocess.PIPE, text=True) for r in range(n)]
                outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in ps]
            agg = {m: sum(o[m] for o in outs) for m in ("both", "h2d", "d2h")}
            key = str(n) if not bind else f"{n}_bound"
            table[key] = {
                "procs": n, "cores_bound": bind, "cpu_count": os.cpu_count(),
                "aggregate_gbs_per_direction_concurrent": round(agg["both"], 2),
                "aggregate_gbs_h2d_alone": round(agg["h2d"], 2),
                "aggregate_gbs_d2h_alone": round(agg["d2h"], 2),
                "per_proc_concurrent": [round(o["both"], 2) for o in outs],
            }
            print(json.dumps({key: table[key]}), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(table, f, indent=1)


if __name__ == "__main__":
    main()
