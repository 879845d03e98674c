#!/usr/bin/env python3
"""Benchmark of the remap hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload e2e_4k_bf16_b32|...]

One "step" = one call of the workload's transform over one batch of synthetic frames.  The
default workload is BASELINE.json configs[1] (the config the metric is quoted on): batch 32
of 3x2048x4096 bf16, bilinear equi2equi, one distinct rotation per frame.  Under torchrun
(N > 1) every rank owns one GPU; frames are independent, so there is no collective on the
data path.  "weak" workloads give every rank its own batch; the "strong" one (config 5: one
equirect x 256 views) splits the views over the ranks.  The time is the max over ranks;
rank 0 prints ONE JSON line.

Other workloads (--workload): config 4 both directions (8K uint8 bicubic, 8 frames per GPU)
and config 5; fp32 / uint8 variants of the headline.

Keys beyond the base contract: ``roofline`` (algorithmic bytes / measured time vs the
measured HBM copy peak; ``traffic`` = DRAM bytes of the dominant kernel from the last ncu
capture, ``traffic_stale`` says whether the library has changed since), ``cpu_baseline``
(the reference's torch-CPU path -- the UNMODIFIED package from baseline/_ref when present,
kind "reference", else oracle/reference_port.py, kind "port" -- plus the reference's numpy
path, timed on this box's host cores on a bounded sample), ``e2e`` (same call with pinned
host buffers, H2D + D2H inside the timed region), ``gpu_launches``, ``clocks``.

``--impl reference`` times the reference's CPU implementation alone (rank 0; other ranks
exit): one frame per step, the driver's --steps / --warmup honoured unless the run would
exceed ~4 minutes (then fewer steps, stated in ``config``).
"""

from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "equi2equi 4K bilinear Mpix/s at 1/2/4/8 B200; % of HBM roofline"

WORKLOADS = {
    # BASELINE.json configs[1] -- the config the metric is quoted on
    "e2e_4k_bf16_b32": dict(transform="equi2equi", batch=32, c=3, h=2048, w=4096, dtype="bfloat16",
                            mode="bilinear", scaling="weak",
                            desc="equi2equi rotation, batch 32 per GPU, 3x2048x4096 bf16, bilinear"),
    "e2e_4k_f32_b16": dict(transform="equi2equi", batch=16, c=3, h=2048, w=4096, dtype="float32",
                           mode="bilinear", scaling="weak",
                           desc="equi2equi rotation, batch 16 per GPU, 3x2048x4096 fp32, bilinear"),
    "e2e_4k_u8_b32": dict(transform="equi2equi", batch=32, c=3, h=2048, w=4096, dtype="uint8",
                          mode="bilinear", scaling="weak",
                          desc="equi2equi rotation, batch 32 per GPU, 3x2048x4096 uint8, bilinear"),
    # BASELINE.json configs[3]: 64 frames of 8K uint8, bicubic, 8 frames per GPU
    "e2p_8k_u8_bicubic_b8": dict(transform="equi2pers", batch=8, c=3, h=4096, w=8192,
                                 out_h=1080, out_w=1920, fov_x=90.0, dtype="uint8",
                                 mode="bicubic", scaling="weak",
                                 desc="equi2pers bicubic, 8 frames per GPU, 3x4096x8192 uint8 -> "
                                      "3x1080x1920, fov_x=90"),
    "p2e_8k_u8_bicubic_b8": dict(transform="pers2equi", batch=8, c=3, h=1080, w=1920,
                                 out_h=4096, out_w=8192, fov_x=90.0, dtype="uint8",
                                 mode="bicubic", scaling="weak",
                                 desc="pers2equi bicubic, 8 frames per GPU, 3x1080x1920 uint8 -> "
                                      "3x4096x8192, fov_x=90"),
    # BASELINE.json configs[4]: one equirect x 256 rotations, split over the GPUs
    "e2p_multiview_f16_v256": dict(transform="equi2pers", batch=256, c=3, h=2048, w=4096,
                                   out_h=1024, out_w=1024, fov_x=90.0, dtype="float16",
                                   mode="bilinear", scaling="strong", one_source=True,
                                   desc="equi2pers multi-view sweep, 1 equirect 3x2048x4096 fp16 x "
                                        "256 rotations -> 3x1024x1024, views split over the GPUs"),
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



def out_hw(wl):
    return wl.get("out_h", wl["h"]), wl.get("out_w", wl["w"])


def rotations_for(wl, n, offset=0, rank=0):
    if wl.get("one_source"):
        # yaw / pitch sweep of one panorama (config 5)
        return [{"roll": 0.0, "pitch": 0.35 * math.sin(0.7 * (i + offset)),
                 "yaw": 2 * math.pi * (i + offset) / 256.0} for i in range(n)]
    return bench_rotations(n, offset=offset, rank=rank)


def synth_frames(wl, n, device, seed):
    """Band-limited pattern + noise, generated on the device (synthetic data)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    c, h, w = wl["c"], wl["h"], wl["w"]
    x = torch.arange(w, device=device, dtype=torch.float32)[None, None, None, :]
    y = torch.arange(h, device=device, dtype=torch.float32)[None, None, :, None]
    ph = torch.rand((n, c, 1, 1), device=device, generator=g) * 6.28318
    out = torch.empty((n, c, h, w), dtype=TORCH_DTYPES[wl["dtype"]], device=device)
    for i in range(n):  # frame by frame: keeps the fp32 temporaries small
        v = 0.5 + 0.4 * torch.sin(x * (6.28318 / 512) + ph[i:i + 1]) * torch.cos(y * (6.28318 / 384))
        v = v + 0.1 * (torch.rand((1, c, h, w), device=device, generator=g) - 0.5)
        if wl["dtype"] == "uint8":
            v = (v.clamp(0, 1) * 255).round()
        out[i:i + 1] = v.to(out.dtype)
    return out


def make_op(pkg, wl):
    """The workload's transform bound to its configuration: fn(img, rots) -> out.  `pkg` is
    equilib_b200 or the reference package -- same class API."""
    t, oh, ow = wl["transform"], *out_hw(wl)
    if t == "equi2equi":
        return pkg.Equi2Equi(mode=wl["mode"])
    if t == "equi2pers":
        return pkg.Equi2Pers(oh, ow, wl["fov_x"], mode=wl["mode"])
    if t == "pers2equi":
        op = pkg.Pers2Equi(oh, ow, mode=wl["mode"])
        return lambda img, rots, **kw: op(img, rots, wl["fov_x"], **kw)
    raise ValueError(t)


def footprint_px(wl, rots, device):
    """Unique source pixels the taps of an equi2pers workload touch ("unique
    source = the sampled footprint"), counted exactly: the product's own coordinates
    (eqb_remap_coords) -> every tap of every output pixel marked in a boolean image.
    Outside the timed region; one frame at a time."""
    from equilib_b200 import _lib, _ops, _prep

    oh, ow = out_hw(wl)
    h, w = wl["h"], wl["w"]
    taps = range(-1, 3) if wl["mode"] == "bicubic" else range(0, 2)
    total = 0
    for r in rots:
        mats = _prep.mats_for("equi2pers", [r], False, device, (oh, ow, float(wl["fov_x"]), 0.0))
        co = torch.empty((1, 2, oh, ow), dtype=torch.float32, device=device)
        _ops.coords(co, mats, None, None, None, None, _lib.EQUI2PERS, 1, h, w, oh, ow)
        y0 = co[0, 0].floor().long()
        x0 = co[0, 1].floor().long()
        mask = torch.zeros(h * w, dtype=torch.bool, device=device)
        for dy in taps:
            for dx in taps:
                yy = (y0 + dy).clamp(0, h - 1)
                xx = (x0 + dx).clamp(0, w - 1)
                mask[(yy * w + xx).reshape(-1)] = True
        total += int(mask.sum())
    return total


def algorithmic_bytes(wl, n_items, rots, device):
    """unique source elements that must be read + output elements written, per
    launch (one rank).  -> (bytes, note)"""
    elt, c = ELT[wl["dtype"]], wl["c"]
    oh, ow = out_hw(wl)
    out_b = n_items * c * oh * ow * elt
    t = wl["transform"]
    if t == "equi2pers" and wl.get("one_source"):
        return c * wl["h"] * wl["w"] * elt + out_b, "one source frame read once + all views written"
    if t == "equi2pers":
        fp = footprint_px(wl, rots, device)
        return fp * c * elt + out_b, (f"sampled footprint ({fp / (n_items * wl['h'] * wl['w']):.3f} "
                                      "of the source pixels, counted exactly) + output written; "
                                      f"write-only lower bound {out_b}")
    return n_items * c * wl["h"] * wl["w"] * elt + out_b, "whole source read once + whole output written once"


# ---------------------------------------------------------------------------------------
# CPU legs: the reference itself (baseline/_ref) or its port (oracle/reference_port.py)
# ---------------------------------------------------------------------------------------


def load_reference():
    """The unmodified reference package from baseline/_ref (git-ignored; it travels with
    the repo snapshot), or None.  Imported only by the CPU legs below."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "equilib")):
        return None
    sys.path.insert(0, ref_dir)
    try:
        import equilib

        return equilib
    except Exception:
        sys.path.remove(ref_dir)
        return None


def cpu_frames(wl, frames, numpy=False):
    g = torch.Generator().manual_seed(0)
    shape = (frames, wl["c"], wl["h"], wl["w"])
    if wl["dtype"] == "uint8":
        src = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)
    else:
        src = torch.rand(shape, generator=g)  # fp32 stand-in for fp16 / bf16 (the reference
        # rejects bf16 and fp16-on-CPU, equilib/equi2equi/torch.py:87-95,109-114)
    return src.numpy() if numpy else src


def cpu_reference_run(wl, frames, steps, warmup, budget_s=240.0, numpy=False):
    """The reference's CPU path on `frames` frames per step.  -> dict(mpix, sec, steps, kind,
    path).  Stops early (never below one timed step) when `budget_s` would be exceeded."""
    torch.set_num_threads(os.cpu_count() or 1)
    ref = load_reference()
    oh, ow = out_hw(wl)
    if ref is not None:
        op = make_op(ref, wl)
        kind = "reference"
        path = ("baseline/_ref/equilib (unmodified reference v%s), %s path, class API (cached "
                "ray grid)" % (getattr(ref, "__version__", "?"), "numpy" if numpy else "torch-CPU"))
    else:
        if numpy or wl["transform"] != "equi2equi":
            return None
        from oracle.reference_port import Equi2EquiCPU

        port = Equi2EquiCPU(mode=wl["mode"])

        def op(src, rots):
            out = port(src.float(), rots)
            return out.type(torch.uint8) if src.dtype == torch.uint8 else out

        kind = "port"
        path = "oracle/reference_port.py (the reference's torch-CPU op sequence; baseline/_ref absent)"
    src = cpu_frames(wl, frames, numpy=numpy)
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        rots = rotations_for(wl, frames, offset=i)
        t0 = time.perf_counter()
        op(src, rots)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        spent = time.perf_counter() - t_start
        if times and spent + dt > budget_s:
            break
        if not times and i + 1 >= warmup:
            pass
        elif not times and spent + 2 * dt > budget_s:
            warmup = i + 1  # cut the warm-up short, keep at least one timed step
    sec = sum(times) / len(times)
    return {"mpix": frames * oh * ow / sec / 1e6, "sec": sec, "steps": len(times), "kind": kind,
            "path": path, "threads": torch.get_num_threads()}


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = 1
    r = cpu_reference_run(wl, frames, max(1, args.steps), max(0, args.warmup))
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable":
                          "baseline/_ref absent and no port of this transform"}), flush=True)
        return
    c, h, w = wl["c"], wl["h"], wl["w"]
    sample = (f"{frames} frame of {c}x{h}x{w} per step ({wl['dtype']} workload: "
              f"{'uint8 in/out' if wl['dtype'] == 'uint8' else 'fp32 on CPU'}), "
              f"{r['steps']} timed steps, {r['sec'] * 1e3:.0f} ms per step")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": r["mpix"], "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["sec"] * 1e3,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "device": "host CPU", "path": r["path"],
                   "sample": "one frame per step (a bounded sample of the workload's batch; "
                             "the rate is per pixel)",
                   "steps_requested": args.steps,
                   "steps_note": ("as requested" if r["steps"] == args.steps else
                                  "fewer than requested: the run is capped at ~4 minutes")},
        "cpu_baseline": {"value": r["mpix"], "unit": "Mpix/s", "cores": r["threads"],
                         "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["mpix"], "unit": "Mpix/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


KERNELS = {
    ("equi2equi", "bilinear", 2): "eqb::remap_lean2_kernel<EQUI2EQUI,T,3,EXACT=0> (TMA boxes, "
                                  "strip-node coordinates)",
    ("equi2equi", "bilinear", 1): "eqb::remap_lean_kernel<EQUI2EQUI,uint8_t> (TMA boxes, "
                                  "interpolated coordinates)",
    ("equi2equi", "bilinear", 4): "eqb::remap_lean2_kernel<EQUI2EQUI,float,3,EXACT=1> (TMA boxes, "
                                  "exact coordinates)",
    ("equi2pers", "bilinear", 2): "eqb::remap_lean2_kernel<EQUI2PERS,T,3,EXACT=0>",
    ("equi2pers", "bicubic", 1): "eqb::remap_lean2_kernel<EQUI2PERS,uint8_t,3,EXACT=0,BICUBIC> "
                                 "(4x4 taps from TMA boxes, strip-node coordinates)",
    ("pers2equi", "bicubic", 1): "eqb::remap_lean2_kernel<PERS2EQUI,uint8_t,3,EXACT=1,BICUBIC> "
                                 "(fill strips from tile corners, in-view tiles from TMA boxes)",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="e2e_4k_bf16_b32", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunk", type=int, default=0, help="e2e pipeline chunk (0 = default)")
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

    c, h, w = wl["c"], wl["h"], wl["w"]
    oh, ow = out_hw(wl)
    strong = wl["scaling"] == "strong"
    if strong:
        if wl["batch"] % world:
            raise SystemExit(f"{wl['batch']} views do not split over {world} ranks")
        n_local = wl["batch"] // world
        offset = rank * n_local
    else:
        n_local, offset = wl["batch"], 0
    if wl.get("one_source"):
        base = synth_frames(wl, 1, dev, seed=0)        # every rank holds the same panorama
        src = base.expand(n_local, -1, -1, -1)         # stride-0 batch: nothing is replicated
    else:
        src = synth_frames(wl, n_local, dev, seed=rank)
    rots = rotations_for(wl, n_local, offset=offset, rank=rank)
    op = make_op(equilib_b200, wl)

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
    px_per_step = n_local * oh * ow * world
    value = px_per_step / (ms_per_step * 1e-3) / 1e6
    out_shape, out_dtype = tuple(out.shape), out.dtype
    del out

    # ---- end to end: pinned host frames in, pinned host frames out, per step, through
    # the public streaming helpers (H2D | remap | D2H overlapped)
    e2e = None
    if not args.no_e2e:
        from equilib_b200.pipeline import HostPipeline, HostViewsPipeline

        hout = torch.empty(out_shape, dtype=out_dtype, pin_memory=True)
        if wl.get("one_source"):
            hin = torch.empty(base.shape[1:], dtype=base.dtype, pin_memory=True)
            hin.copy_(base[0])
            chunk = args.chunk or 32
            pipe = HostViewsPipeline(op, chunk=chunk, device=dev)
            how = (f"pinned host equirect -> equilib_b200.pipeline.HostViewsPipeline(Equi2Pers, "
                   f"chunk={chunk}): one H2D | views | D2H -> pinned host")
        else:
            hin = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            hin.copy_(src)
            # chunks of one frame where a frame is large (4K: 25-100 MB): the pipeline's fill
            # and drain (one chunk each way with nothing to overlap) then cost least --
            # measured on 32 x 4K bf16: chunk 1 / 2 / 4 / 8 = 7.64 / 7.49 / 6.84 / 6.12 Gpix/s
            frame_bytes = c * h * w * ELT[wl["dtype"]]
            chunk = args.chunk or (1 if frame_bytes >= (8 << 20) else 4)
            pipe = HostPipeline(op, chunk=chunk, device=dev)
            how = (f"pinned host -> equilib_b200.pipeline.HostPipeline({wl['transform']}, "
                   f"chunk={chunk}, depth=3): H2D | remap | D2H on three streams -> pinned host; "
                   "the call returns when the last D2H copy has finished")
        steps2 = max(2, min(args.steps, 5))
        pipe(hin, rots, out=hout)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps2):
            pipe(hin, rots, out=hout)   # (synchronises the last D2H copy itself)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t2 = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": px_per_step / (float(t2.item()) / steps2) / 1e6, "unit": "Mpix/s",
               "h2d_bytes_per_step": hin.numel() * hin.element_size() * world,
               "d2h_bytes_per_step": hout.numel() * hout.element_size() * world,
               "steps": steps2, "path": how}
        ceiling = host_copy_ceiling(world)
        if ceiling is not None:
            e2e["host_copy_ceiling"] = ceiling
        del hin, hout

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    elt = ELT[wl["dtype"]]
    alg_bytes, alg_note = algorithmic_bytes(wl, n_local, rots, dev)
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9  # per GPU: one launch per step per GPU
    traffic, traffic_stale, traffic_kernel = None, None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            ent = json.load(open(tp)).get(args.workload)
            if isinstance(ent, dict):
                traffic, traffic_kernel = ent.get("bytes"), ent.get("kernel")
                traffic_stale = ent.get("source_hash") != _lib.source_hash()
            elif ent is not None:
                traffic, traffic_stale = ent, True
        except Exception:
            traffic = None
    in_mb = (c * h * w * elt if wl.get("one_source") else n_local * c * h * w * elt) / 1e6
    out_mb = n_local * c * oh * ow * elt / 1e6
    line = {
        "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "storage_dtype": wl["dtype"],
                   "global_batch": n_local * world,
                   "parallelism": f"{'views' if strong else 'frames'} sharded over {world} GPU(s), "
                                  "no data-path collective",
                   "l2": (f"no flush: per-step input {in_mb:.0f} MB + output {out_mb:.0f} MB per "
                          "GPU" + (" exceed the 126 MB L2" if in_mb + out_mb > 126 else
                                   " (the source is meant to stay L2-resident; the output alone "
                                   "exceeds L2)" if out_mb > 126 else " -- fits L2, latency-bound")),
                   "clip": ("bilinear: min/max pass skipped (<=2 ulp no-op, DESIGN.md Q3)"
                            if wl["mode"] == "bilinear" else
                            "uint8: no clamp (reference: truncating cast)"
                            if wl["dtype"] == "uint8" else "min/max pass + clamp")},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_stale": traffic_stale,
                     "traffic_kernel": traffic_kernel, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_bytes_note": alg_note,
                     "library_source_hash": _lib.source_hash(),
                     "kernel": KERNELS.get((wl["transform"], wl["mode"], elt), "see DESIGN.md 4")},
        "gpu_launches": int(launches) * world,
        "clocks": clocks,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(wl, 1, steps=3, warmup=1, budget_s=60.0)
        if r is not None:
            line["cpu_baseline"] = {
                "value": r["mpix"], "unit": "Mpix/s", "cores": r["threads"], "kind": r["kind"],
                "sample": f"1 frame of {c}x{h}x{w} per step, {r['steps']} timed steps "
                          f"({r['sec'] * 1e3:.0f} ms per frame), {r['path']}"}
            rn = cpu_reference_run(wl, 1, steps=1, warmup=0, budget_s=60.0, numpy=True)
            if rn is not None:
                line["cpu_baseline"]["numpy"] = {
                    "value": rn["mpix"], "unit": "Mpix/s",
                    "sample": f"1 frame, 1 timed step ({rn['sec'] * 1e3:.0f} ms), {rn['path']}"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def host_copy_ceiling(world):
    """The box's measured pinned-copy ceiling at this process count, recorded by
    tools/host_copy_ceiling.py (profiles/host_copy_ceiling.json), if present."""
    p = os.path.join(ROOT, "profiles", "host_copy_ceiling.json")
    try:
        ent = json.load(open(p)).get(str(world))
    except Exception:
        return None
    return ent


if __name__ == "__main__":
    main()
