#!/usr/bin/env python3
"""Benchmark of the remap hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload e2e_4k_bf16_b32|...]

One "step" = one ``equi2equi`` call over one batch of synthetic frames.  The default
workload is BASELINE.json configs[1]: batch 32 of 3x2048x4096 bf16, bilinear, one
distinct rotation per frame.  Under torchrun (N > 1) every rank owns one GPU and
its own batch (frames are independent: weak scaling, no collective on the data
path); the time is the max over ranks.  Rank 0 prints ONE JSON line.

Keys beyond the base contract: ``roofline`` (algorithmic bytes / measured time vs
the measured HBM copy peak), ``cpu_baseline`` (the reference's torch-CPU algorithm,
oracle/reference_port.py, timed on this box's host cores on a bounded sample),
``e2e`` (same call with pinned host buffers, H2D + D2H inside the timed region),
``gpu_launches``, ``clocks``.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "equi2equi 4K bilinear Mpix/s at 1/2/4/8 B200; % of HBM roofline"

WORKLOADS = {
    # BASELINE.json configs[1] -- the config the metric is quoted on
    "e2e_4k_bf16_b32": dict(transform="equi2equi", batch=32, c=3, h=2048, w=4096, dtype="bfloat16",
                            mode="bilinear",
                            desc="equi2equi rotation, batch 32 per GPU, 3x2048x4096 bf16, bilinear"),
    "e2e_4k_f32_b16": dict(transform="equi2equi", batch=16, c=3, h=2048, w=4096, dtype="float32",
                           mode="bilinear",
                           desc="equi2equi rotation, batch 16 per GPU, 3x2048x4096 fp32, bilinear"),
    "e2e_4k_u8_b32": dict(transform="equi2equi", batch=32, c=3, h=2048, w=4096, dtype="uint8",
                          mode="bilinear",
                          desc="equi2equi rotation, batch 32 per GPU, 3x2048x4096 uint8, bilinear"),
}
TORCH_DTYPES = {"bfloat16": torch.bfloat16, "float32": torch.float32, "uint8": torch.uint8,
                "float16": torch.float16}
ELT = {"bfloat16": 2, "float16": 2, "float32": 4, "uint8": 1}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bench_rotations(n, offset=0, rank=0):
    """n distinct rotations (roll 0.01 i, pitch 0.4 sin i, yaw 0.05 i: the pattern of the
    reference's benchmarks/cache_benchmark.py:37-40).  Every rank of a multi-GPU run gets
    the same roll / pitch pattern under its own yaw shift: rotations stay distinct across
    the job while the per-GPU work is the same on every rank (weak scaling), instead of
    later ranks getting ever more strongly rolled -- and more expensive -- views."""
    return [{"roll": 0.01 * (i + offset), "pitch": 0.4 * math.sin(i + offset),
             "yaw": 0.05 * (i + offset) + 0.61 * rank} for i in range(n)]


def synth_frames(wl, device, seed):
    """Band-limited pattern + noise, generated on the device (synthetic data)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    b, c, h, w = wl["batch"], wl["c"], wl["h"], wl["w"]
    x = torch.arange(w, device=device, dtype=torch.float32)[None, None, None, :]
    y = torch.arange(h, device=device, dtype=torch.float32)[None, None, :, None]
    ph = torch.rand((b, c, 1, 1), device=device, generator=g) * 6.28318
    out = torch.empty((b, c, h, w), dtype=TORCH_DTYPES[wl["dtype"]], device=device)
    for i in range(b):  # frame by frame: keeps the fp32 temporaries small
        v = 0.5 + 0.4 * torch.sin(x * (6.28318 / 512) + ph[i:i + 1]) * torch.cos(y * (6.28318 / 384))
        v = v + 0.1 * (torch.rand((1, c, h, w), device=device, generator=g) - 0.5)
        if wl["dtype"] == "uint8":
            v = (v.clamp(0, 1) * 255).round()
        out[i:i + 1] = v.to(out.dtype)
    return out


class ClockSampler:
    """SM clock / power / throttle reasons sampled every few ms DURING the timed region
    (NVML from a thread; the timed region is tens of ms, too short for `nvidia-smi -lms`)."""

    def __init__(self, index, period_s=0.004):
        self.index = index
        self.period = period_s
        self.samples = []
        self._stop = False
        self._thread = None
        self._nvml = None

    def start(self):
        import threading

        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        except Exception:
            self._nvml = None
            return
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def _loop(self):
        n = self._nvml
        while not self._stop:
            try:
                sm = n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)
                try:
                    reasons = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    reasons = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                pw = 0.0
                if len(self.samples) % 4 == 0:  # power is the slow query
                    pw = n.nvmlDeviceGetPowerUsage(self._h) / 1000.0
                self.samples.append((sm, reasons, pw))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self._stop = True
        if self._thread is not None:
            self._thread.join(timeout=2)
        n = self._nvml
        if n is None or not self.samples:
            return out
        try:
            out["sm_max_mhz"] = float(n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM))
        except Exception:
            pass
        out["sm_mhz"] = float(statistics.median(s[0] for s in self.samples))
        out["power_w_max"] = max(s[2] for s in self.samples)
        out["samples"] = len(self.samples)
        bits = 0
        for s_ in self.samples:
            bits |= int(s_[1])
        names = {
            "hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        out["reasons"] = [k for k, v in names.items() if bits & int(v)]
        return out


def cpu_reference_run(wl, frames, steps, warmup):
    """The reference's torch-CPU algorithm on `frames` frames per step; returns Mpix/s.
    fp32 stand-in for 16-bit workloads: the reference rejects bf16, and fp16 on CPU
    (equilib/equi2equi/torch.py:87-95,109-114)."""
    from oracle.reference_port import Equi2EquiCPU

    torch.set_num_threads(os.cpu_count() or 1)
    h, w, c = wl["h"], wl["w"], wl["c"]
    g = torch.Generator().manual_seed(0)
    if wl["dtype"] == "uint8":
        src = torch.randint(0, 256, (frames, c, h, w), generator=g).float()
    else:
        src = torch.rand((frames, c, h, w), generator=g)
    op = Equi2EquiCPU(mode=wl["mode"])
    times = []
    for i in range(warmup + steps):
        rots = bench_rotations(frames, offset=i)
        t0 = time.perf_counter()
        out = op(src, rots)
        if wl["dtype"] == "uint8":
            out = out.type(torch.uint8)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    px = frames * h * w
    return px / (sum(times) / len(times)) / 1e6, sum(times) / len(times)


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = 1
    steps = max(1, min(args.steps, 12))
    warmup = max(1, min(args.warmup, 2))
    mpix, sec = cpu_reference_run(wl, frames, steps, warmup)
    cores = torch.get_num_threads()
    sample = (f"{frames} frame(s) of {wl['c']}x{wl['h']}x{wl['w']} per step, fp32 on CPU"
              f" ({'uint8 in/out' if wl['dtype'] == 'uint8' else 'stand-in for ' + wl['dtype']}),"
              f" {steps} timed steps")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "device": "host CPU",
                   "path": "oracle/reference_port.py (reference torch-CPU op sequence)"},
        "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="e2e_4k_bf16_b32", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stage", action="store_true",
                    help="experiments: force the direct-gather kernel")
    ap.add_argument("--exact-coords", action="store_true",
                    help="experiments: exact coordinate chain on every pixel")
    ap.add_argument("--warp-shape", type=int, default=0,
                    help="experiments: force the warp tile (1=32x1, 2=16x2, 3=8x4 lanes)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.warp_shape:
        os.environ["EQUILIB_B200_WARP_SHAPE"] = str(args.warp_shape)
    if args.no_stage:
        os.environ["EQUILIB_B200_NO_STAGE"] = "1"
    if args.exact_coords:
        os.environ["EQUILIB_B200_FAST_COORDS"] = "0"
    if args.impl == "reference":
        return run_reference(args, wl)

    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import equilib_b200
    from equilib_b200 import _lib

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    b, c, h, w = wl["batch"], wl["c"], wl["h"], wl["w"]
    src = synth_frames(wl, dev, seed=rank)
    rots = bench_rotations(b, rank=rank)
    op = equilib_b200.Equi2Equi(mode=wl["mode"])

    for _ in range(args.warmup):
        out = op(src, rots)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = op(src, rots)
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_per_step = ms_max / args.steps
    px_per_step = b * h * w * world
    value = px_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- end to end: pinned host frames in, pinned host frames out, per step, through
    # the public streaming helper (H2D | remap | D2H overlapped on three streams)
    e2e = None
    if not args.no_e2e:
        from equilib_b200.pipeline import HostPipeline

        hin = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
        hin.copy_(src)
        hout = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
        pipe = HostPipeline(op, chunk=4, device=dev)
        steps2 = max(2, min(args.steps, 5))
        pipe(hin, rots, out=hout)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps2):
            pipe(hin, rots, out=hout)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t2 = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": px_per_step / (float(t2.item()) / steps2) / 1e6, "unit": "Mpix/s",
               "h2d_bytes_per_step": hin.numel() * hin.element_size(),
               "d2h_bytes_per_step": hout.numel() * hout.element_size(),
               "steps": steps2,
               "path": "pinned host -> equilib_b200.pipeline.HostPipeline(Equi2Equi, chunk=4): "
                       "H2D | remap | D2H on three streams -> pinned host"}
        del hin, hout

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    elt = ELT[wl["dtype"]]
    alg_bytes = b * c * h * w * elt * 2  # whole source read once + whole output written once
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9  # per GPU: one launch per step per GPU
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload)
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "storage_dtype": wl["dtype"],
                   "global_batch": b * world, "parallelism": f"frames sharded over {world} GPU(s), "
                   "no data-path collective",
                   "l2": f"no flush: per-step input {alg_bytes // 2 / 1e6:.0f} MB + output "
                         f"{alg_bytes // 2 / 1e6:.0f} MB per GPU exceed the 126 MB L2",
                   "clip": "bilinear: min/max pass skipped (<=2 ulp no-op, DESIGN.md Q3)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "kernel": ("eqb::remap_staged_kernel<EQUI2EQUI,BILINEAR,float,PX=2> "
                                "(TMA-staged, exact coordinates)"
                                if wl["dtype"] == "float32" else
                                "eqb::remap_lean_kernel<EQUI2EQUI,%s> (TMA-staged, fast "
                                "coordinates, 4-warp CTAs)" % wl["dtype"])},
        "gpu_launches": int(launches) * world,
        "clocks": clocks,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        frames = 1
        mpix, sec = cpu_reference_run(wl, frames, steps=3, warmup=1)
        line["cpu_baseline"] = {
            "value": mpix, "unit": "Mpix/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{frames} frame of {c}x{h}x{w} fp32 per step, 3 timed steps "
                      f"({sec * 1e3:.0f} ms/frame), oracle/reference_port.py"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
