#!/usr/bin/env python3
"""A/B timing of what the second half of round 2 added (CUDA events, one GPU):

* bicubic and nearest through the strip-node kernel (remap_lean2_kernel<..., BICUBIC / NEAREST>)
  against the kernels they replace (EQUILIB_B200_NO_LEAN2=1: generic staged / direct), headline
  shapes + config 4 equi2pers;
* the fused equi -> cube -> equi chain (eqb_equi2cube2equi) against the two launches it
  replaces, config 3 shapes.

    python tools/bench_round2b.py > gpurun_out/round2b.jsonl
"""

import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import equilib_b200  # noqa: E402
from equilib_b200 import _lib  # noqa: E402

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
           "frac_hbm_6553": round(alg_bytes / ms / 1e6 / 6553.3, 4),
           "library_source_hash": _lib.source_hash()}
    rec.update(kw)
    print(json.dumps(rec), flush=True)


def ab(name, fn, out_px, alg_bytes, **kw):
    os.environ.pop("EQUILIB_B200_NO_LEAN2", None)
    report(name + " [strip-node kernel]", timed(fn), out_px, alg_bytes, **kw)
    os.environ["EQUILIB_B200_NO_LEAN2"] = "1"
    report(name + " [previous kernel, EQUILIB_B200_NO_LEAN2=1]", timed(fn), out_px,
           alg_bytes, **kw)
    os.environ.pop("EQUILIB_B200_NO_LEAN2", None)


def main():
    h, w = 2048, 4096
    for dtype, b in ((torch.bfloat16, 32), (torch.float32, 16), (torch.uint8, 32),
                     (torch.float16, 32)):
        x = frames(b, 3, h, w, dtype)
        r = rots(b)
        es = x.element_size()
        name = f"2: equi2equi bicubic {str(dtype)[6:]} {b}x3x{h}x{w}"
        ab(name + " (with the clamp's min/max pass)",
           lambda: equilib_b200.equi2equi(x, r, mode="bicubic"), b * h * w, 2 * x.numel() * es)
        ab(name + " clip_output=False",
           lambda: equilib_b200.equi2equi(x, r, mode="bicubic", clip_output=False), b * h * w,
           2 * x.numel() * es)
        nname = f"2: equi2equi nearest {str(dtype)[6:]} {b}x3x{h}x{w}"
        ab(nname, lambda: equilib_b200.equi2equi(x, r, mode="nearest"), b * h * w,
           2 * x.numel() * es)
        if dtype != torch.float32:
            report(nname + " fast_coords=False [strip-node kernel, exact chain]",
                   timed(lambda: equilib_b200.equi2equi(x, r, mode="nearest", fast_coords=False)),
                   b * h * w, 2 * x.numel() * es)
        if dtype != torch.float16:
            ab(name + " fast_coords=True clip_output=False",
               lambda: equilib_b200.equi2equi(x, r, mode="bicubic", clip_output=False,
                                              fast_coords=True), b * h * w, 2 * x.numel() * es)
        del x
    # config 3 stage: equi2cube bicubic is not a BASELINE row; config 4 equi2pers is
    equi = frames(8, 3, 4096, 8192, torch.uint8)
    r = rots(8)
    ab("4: equi2pers uint8 bicubic 8K -> 1080x1920, 8 frames",
       lambda: equilib_b200.equi2pers(equi, r, 1080, 1920, 90.0, mode="bicubic"),
       8 * 1080 * 1920, 8 * 3 * 1080 * 1920 * 2, note="bytes: output written + as many read")
    del equi
    x = frames(64, 3, 2048, 4096, torch.float16)[:1].expand(64, -1, -1, -1)
    r = rots(64)
    ab("5-like: equi2pers fp16 bicubic, 1 equirect x 64 views 1024^2",
       lambda: equilib_b200.equi2pers(x, r, 1024, 1024, 90.0, mode="bicubic", clip_output=False),
       64 * 1024 * 1024, 64 * 3 * 1024 * 1024 * 2)
    del x

    # fused chain, config 3: 16 x 3 x 2048 x 4096, w_face 1024
    for dtype in (torch.float32, torch.uint8, torch.bfloat16):
        b = 16
        x = frames(b, 3, h, w, dtype)
        r = rots(b)
        es = x.element_size()
        for mode in ("nearest", "bilinear"):
            def two(fast=None):
                cube = equilib_b200.equi2cube(x, r, 1024, "horizon", mode=mode, fast_coords=fast)
                return equilib_b200.cube2equi(cube, "horizon", h, w, mode=mode)
            def one():
                return equilib_b200.equi2cube2equi(x, r, 1024, h, w, mode=mode, fused=True)
            alg = 2 * x.numel() * es
            name = f"N2: equi->cube(1024)->equi {mode} {str(dtype)[6:]} {b}x3x{h}x{w}"
            report(name + " [two launches, default coordinates]", timed(two), b * h * w, alg)
            report(name + " [two launches, fast_coords=False]", timed(lambda: two(False)),
                   b * h * w, alg)
            report(name + " [fused, one launch]", timed(one), b * h * w, alg,
                   cubemap_bytes_not_moved=2 * b * 3 * 6 * 1024 * 1024 * es)
        del x


if __name__ == "__main__":
    main()
