#!/usr/bin/env python3
"""Device-resident timing of every BASELINE.json config (CUDA events, one GPU).

    python tools/bench_configs.py > gpurun_out/configs.jsonl

Not the graded bench (that is bench.py on configs[1]); these lines document the other
configs in DESIGN.md.  Mpix/s counts output pixels; GB/s counts algorithmic bytes
(whole source read once + output written once; for equi2pers the source term is the
upper bound "whole frame").
"""

import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import equilib_b200  # noqa: E402
from equilib_b200 import _lib, _ops, _prep  # noqa: E402

DEV = torch.device("cuda", 0)


def rots(n):
    return [{"roll": 0.01 * i, "pitch": 0.4 * math.sin(i), "yaw": 0.05 * i} for i in range(n)]


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def frames(b, c, h, w, dtype):
    if dtype == torch.uint8:
        return torch.randint(0, 256, (b, c, h, w), dtype=torch.uint8, device=DEV)
    return torch.rand((b, c, h, w), device=DEV).to(dtype)


def report(name, ms, out_px, alg_bytes, **kw):
    rec = {"config": name, "ms": round(ms, 4), "mpix_s": round(out_px / ms / 1e3, 1),
           "alg_gb_s": round(alg_bytes / ms / 1e6, 1),
           "frac_hbm_6553": round(alg_bytes / ms / 1e6 / 6553.3, 4)}
    rec.update(kw)
    print(json.dumps(rec), flush=True)


def main():
    # config 1: equi2pers bilinear, batch 1, 3x2048x4096 fp32 -> 480x640, fov 90
    equi = frames(1, 3, 2048, 4096, torch.float32)
    rot = {"roll": 0.1, "pitch": 0.3, "yaw": -0.5}
    op = equilib_b200.Equi2Pers(480, 640, 90.0)
    ms = timed(lambda: op(equi[0], rot), steps=50, warmup=5)
    report("1: equi2pers fp32 4K -> 480x640 (device time per call, back to back)", ms, 480 * 640,
           480 * 640 * 3 * 4 * 2, note="latency-bound; includes host dispatch when it dominates")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        op(equi[0], rot)
    torch.cuda.synchronize()
    report("1: same, wall clock per call incl. python + custom-op dispatch",
           (time.perf_counter() - t0) / 200 * 1e3, 480 * 640, 480 * 640 * 3 * 4 * 2)
    # the same call captured in a CUDA graph and replayed (no python on the replay path)
    e4 = equi.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        op(e4, [rot])
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        op(e4, [rot])
    g.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(500):
        g.replay()
    torch.cuda.synchronize()
    report("1: same, one CUDA-graph replay per call (wall clock)",
           (time.perf_counter() - t0) / 500 * 1e3, 480 * 640, 480 * 640 * 3 * 4 * 2)
    del equi, e4, g

    # config 2 family: equi2equi batch 32 bf16 / fp32 / uint8, three modes
    for dtype, b in ((torch.bfloat16, 32), (torch.float32, 16), (torch.uint8, 32)):
        src = frames(b, 3, 2048, 4096, dtype)
        r = rots(b)
        for mode in ("nearest", "bilinear", "bicubic"):
            ms = timed(lambda: equilib_b200.equi2equi(src, r, mode=mode))
            px = b * 2048 * 4096
            report(f"2: equi2equi {str(dtype)[6:]} b{b} 4K {mode}", ms, px,
                   2 * px * 3 * src.element_size())
        if dtype == torch.bfloat16:
            # the same views rolled by a further 55-73 degrees: the 64-pixel side of a tile
            # then runs down the source columns (tall boxes)
            rr = [dict(x, roll=x["roll"] + 0.96) for x in r]
            ms = timed(lambda: equilib_b200.equi2equi(src, rr))
            report("2: equi2equi bfloat16 b32 4K bilinear, views rolled by a further 0.96 rad", ms,
                   b * 2048 * 4096, 2 * b * 2048 * 4096 * 3 * 2)
        if dtype != torch.float32:
            ms = timed(lambda: equilib_b200.equi2equi(src, r, mode="bicubic", fast_coords=False))
            report(f"2: equi2equi {str(dtype)[6:]} b{b} 4K bicubic fast_coords=False (exact chain)",
                   ms, b * 2048 * 4096, 2 * b * 2048 * 4096 * 3 * src.element_size())
        if dtype == torch.float32:
            ms = timed(lambda: equilib_b200.equi2equi(src, r, mode="bicubic", fast_coords=True))
            report("2: equi2equi float32 b16 4K bicubic fast_coords=True (opt-in)", ms,
                   b * 2048 * 4096, 2 * b * 2048 * 4096 * 3 * 4)
            ms = timed(lambda: equilib_b200.equi2equi(src, r, fast_coords=True))
            report("2: equi2equi float32 b16 4K bilinear fast_coords=True (opt-in)", ms,
                   b * 2048 * 4096, 2 * b * 2048 * 4096 * 3 * 4)
        del src

    # config 2 comparator: torch.nn.functional.grid_sample on the same GPU, fed a
    # pre-built on-device normalised grid of the same rotations (grid build time excluded:
    # the reference builds it on the CPU, ~0.1 s per 4K frame).  Benchmark context only; the
    # product never calls grid_sample.
    import torch.nn.functional as F
    for dtype, b in ((torch.bfloat16, 32), (torch.float32, 16)):
        src = frames(b, 3, 2048, 4096, dtype)
        r = rots(b)
        tab_x, tab_y = _prep.equirect_tables(2048, 4096, DEV)
        coords = torch.empty((b, 2, 2048, 4096), dtype=torch.float32, device=DEV)
        _ops.coords(coords, _prep.mats_for("rot", r, False, DEV), tab_x, tab_y, None, None,
                    _lib.EQUI2EQUI, b, 2048, 4096, 2048, 4096)
        gx = (2 * coords[:, 1] / (4096 - 1) - 1)
        gy = (2 * coords[:, 0] / (2048 - 1) - 1)
        grid = torch.stack((gx, gy), dim=-1).to(dtype).contiguous()
        del coords, gx, gy
        ms = timed(lambda: F.grid_sample(src, grid, mode="bilinear", padding_mode="reflection",
                                         align_corners=True), steps=5)
        px = b * 2048 * 4096
        report(f"2: comparator torch F.grid_sample {str(dtype)[6:]} b{b} 4K bilinear, grid pre-built "
               "on device", ms, px, 2 * px * 3 * src.element_size(),
               note="reads an extra %d B/px of grid" % (2 * src.element_size()))
        del src, grid

    # uint8 RGB / RGBX in channels-last layout (what decoders hand over), read
    # and written in place, next to the planar layout of the same frames
    for c in (3, 4):
        src = frames(32, c, 2048, 4096, torch.uint8)
        r = rots(32)
        px = 32 * 2048 * 4096
        ms = timed(lambda: equilib_b200.equi2equi(src, r))
        report(f"N1: equi2equi uint8 b32 4K bilinear, {c} channels, planar", ms, px, 2 * px * c)
        src_cl = src.contiguous(memory_format=torch.channels_last)
        ms = timed(lambda: equilib_b200.equi2equi(src_cl, r))
        report(f"N1: equi2equi uint8 b32 4K bilinear, {c} channels, channels-last", ms, px,
               2 * px * c)
        del src, src_cl

    # static camera -- the sampling map of fixed rotations built once (8 B/px),
    # then applied per batch of frames; "alg bytes" here include the map read
    for dtype in (torch.bfloat16, torch.uint8):
        src = frames(32, 3, 2048, 4096, dtype)
        r = rots(32)
        px = 32 * 2048 * 4096
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sr = equilib_b200.StaticRemap.equi2equi(r, src.shape, dtype, DEV)
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        ms = timed(lambda: sr(src))
        alg = 2 * px * 3 * src.element_size()  # source + output, the map is overhead
        report(f"N3: equi2equi {str(dtype)[6:]} b32 4K bilinear through a static map", ms, px,
               alg + 8 * px, map_build_ms=round(build_ms, 2),
               frac_hbm_on_8d_bytes=round(alg / ms / 1e6 / 6553.3, 4),
               note="alg_gb_s / frac_hbm_6553 count the 8 B/px map read too")
        sr4 = equilib_b200.StaticRemap.equi2equi(r, src.shape, dtype, DEV, compact=True)
        ms4 = timed(lambda: sr4(src))
        report(f"N3: equi2equi {str(dtype)[6:]} b32 4K bilinear through a COMPACT static map "
               "(4 B/px)", ms4, px, alg + 4 * px,
               frac_hbm_on_8d_bytes=round(alg / ms4 / 1e6 / 6553.3, 4),
               note="alg_gb_s / frac_hbm_6553 count the 4 B/px map read too")
        del sr4
        if dtype == torch.uint8:
            src_cl = src.contiguous(memory_format=torch.channels_last)
            ms = timed(lambda: sr(src_cl))
            report("N3+N1: equi2equi uint8 RGB b32 4K bilinear, channels-last frames through a "
                   "static map", ms, px, 2 * px * 3 + 8 * px)
            del src_cl
        del src, sr

    # config 3: equi2cube + cube2equi, batch 16, w_face 1024, fp32 and uint8
    for dtype in (torch.float32, torch.uint8):
        equi = frames(16, 3, 2048, 4096, dtype)
        r = rots(16)
        es = equi.element_size()
        for fmt in ("dice", "horizon", "dict"):
            for mode in ("nearest", "bilinear"):
                ms = timed(lambda: equilib_b200.equi2cube(equi, r, 1024, fmt, mode=mode), steps=5)
                out_px = 16 * 6 * 1024 * 1024
                extra = 16 * 6 * 1024 * 1024 * 3 * es if fmt == "dice" else 0
                report(f"3: equi2cube {str(dtype)[6:]} b16 4K -> w1024 {fmt} {mode}", ms, out_px,
                       16 * 3 * (2048 * 4096 + 6 * 1024 * 1024) * es + extra)
            cube = equilib_b200.equi2cube(equi, r, 1024, fmt, mode="bilinear")
            for mode in ("nearest", "bilinear"):
                ms = timed(lambda: equilib_b200.cube2equi(cube, fmt, 2048, 4096, mode=mode), steps=5)
                report(f"3: cube2equi {str(dtype)[6:]} b16 w1024 {fmt} -> 4K {mode}", ms,
                       16 * 2048 * 4096, 16 * 3 * (2048 * 4096 + 6 * 1024 * 1024) * es)
            del cube
        # the round trip as ONE call: nearest runs the fused kernel (no cubemap
        # in memory), bilinear the two launches (equi2cube2equi picks by measurement)
        for mode in ("nearest", "bilinear"):
            alg = 2 * 16 * 3 * 2048 * 4096 * es
            ms = timed(lambda: equilib_b200.equi2cube2equi(equi, r, 1024, 2048, 4096, mode=mode),
                       steps=5)
            report(f"3+N2: equi -> cube(1024) -> equi {str(dtype)[6:]} b16 4K {mode}, "
                   "equi2cube2equi (default)", ms, 16 * 2048 * 4096, alg)
            ms = timed(lambda: equilib_b200.equi2cube2equi(equi, r, 1024, 2048, 4096, mode=mode,
                                                           fused=True), steps=5)
            report(f"3+N2: equi -> cube(1024) -> equi {str(dtype)[6:]} b16 4K {mode}, fused=True "
                   "(one launch, no cubemap)", ms, 16 * 2048 * 4096, alg)
            ms = timed(lambda: equilib_b200.equi2cube2equi(equi, r, 1024, 2048, 4096, mode=mode,
                                                           fused=False), steps=5)
            report(f"3+N2: equi -> cube(1024) -> equi {str(dtype)[6:]} b16 4K {mode}, fused=False "
                   "(two launches)", ms, 16 * 2048 * 4096, alg)
        del equi

    # config 4 (per GPU share: 8 frames of 3x4096x8192 uint8, bicubic)
    equi = frames(8, 3, 4096, 8192, torch.uint8)
    r = rots(8)
    ms = timed(lambda: equilib_b200.equi2pers(equi, r, 1080, 1920, 90.0, mode="bicubic"), steps=5)
    report("4: equi2pers uint8 bicubic 8K -> 1080x1920, 8 frames", ms, 8 * 1080 * 1920,
           8 * 1080 * 1920 * 3 * 2, note="alg bytes = write + equal-size footprint lower bound")
    pers = frames(8, 3, 1080, 1920, torch.uint8)
    ms = timed(lambda: equilib_b200.pers2equi(pers, r, 4096, 8192, 90.0, mode="bicubic"), steps=5)
    report("4: pers2equi uint8 bicubic 1080x1920 -> 8K, 8 frames", ms, 8 * 4096 * 8192,
           8 * 3 * (4096 * 8192 + 1080 * 1920))
    del equi, pers

    # config 5: one fp16 equirect x 256 rotations -> 3x1024x1024 (stride-0 batch)
    equi = frames(1, 3, 2048, 4096, torch.float16).expand(256, -1, -1, -1)
    r = [{"roll": 0.0, "pitch": 0.5 * math.sin(0.1 * i), "yaw": 2 * math.pi * i / 256}
         for i in range(256)]
    ms = timed(lambda: equilib_b200.equi2pers(equi, r, 1024, 1024, 90.0), steps=5)
    report("5: equi2pers fp16 1 x 256 views -> 1024x1024 (expanded batch)", ms, 256 * 1024 * 1024,
           256 * 1024 * 1024 * 3 * 2 + 3 * 2048 * 4096 * 2)
    r2 = [dict(x, roll=0.001) for x in r]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    equilib_b200.equi2pers(equi, r2, 1024, 1024, 90.0)
    torch.cuda.synchronize()
    report("5: same, first call with NEW rotations (host matrix prep + upload + launch), wall",
           (time.perf_counter() - t0) * 1e3, 256 * 1024 * 1024,
           256 * 1024 * 1024 * 3 * 2 + 3 * 2048 * 4096 * 2)


if __name__ == "__main__":
    main()
