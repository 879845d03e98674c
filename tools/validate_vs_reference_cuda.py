#!/usr/bin/env python3
"""One-off validation on a GPU box: equilib_b200 vs the UNMODIFIED reference's own
CUDA path (CPU grid -> H2D -> CUDA F.grid_sample), which is the parity target of
SURVEY.md 8(c).  Needs a copy of the reference package under --ref (git-ignored
baseline/_ref; never imported by the product, the tests or bench.py).

    python tools/validate_vs_reference_cuda.py --ref baseline/_ref > gpurun_out/validate.log
"""

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import cases  # noqa: E402


def stats(name, out, ref, extra=None):
    o = out.float().cpu().numpy().astype(np.float64)
    r = ref.float().cpu().numpy().astype(np.float64)
    d = np.abs(o - r)
    rec = {
        "case": name,
        "max_abs": float(d.max()),
        "mean_abs": float(d.mean()),
        "frac_equal": float(np.mean(o == r)),
        "frac_1e5": float(np.mean(d <= 1e-5 + 1e-5 * np.abs(r))),
        "frac_1e4": float(np.mean(d <= 1e-4 + 1e-4 * np.abs(r))),
    }
    if extra:
        rec.update(extra)
    print(json.dumps(rec), flush=True)
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.path.join(ROOT, "baseline", "_ref"))
    ap.add_argument("--size", type=int, default=1024)
    args = ap.parse_args()
    sys.path.insert(0, args.ref)
    import equilib
    import equilib_b200

    dev = torch.device("cuda", 0)
    h, w = args.size, 2 * args.size
    rots = cases.bench_rotations(4)[2:]
    imgs = {
        "band": torch.from_numpy(cases.band_limited_image((2, 3, h, w), "float32")).to(dev),
        "noise": torch.from_numpy(cases.noise_image((2, 3, h, w), "float32", 0)).to(dev),
    }
    # does the CUDA reference normalise with a reciprocal multiply?
    g = torch.rand(1000, device=dev) * 4000
    a = 2 * g / 4095 - 1
    recip = (2 * g) * (torch.tensor(1.0) / torch.tensor(4095.0)).item() - 1
    inv32 = np.float32(1.0) / np.float32(4095.0)
    b = (2 * g) * float(inv32) - 1
    truediv = torch.from_numpy((2 * g.cpu().numpy() / np.float32(4095) - 1).astype(np.float32)).to(dev)
    print(json.dumps({"probe": "cuda tensor / python int",
                      "equals_reciprocal_mul": float((a == b).float().mean()),
                      "equals_true_div": float((a == truediv).float().mean())}), flush=True)

    for kind, img in imgs.items():
        for mode in ("nearest", "bilinear", "bicubic"):
            ref = equilib.equi2equi(img.clone(), rots, mode=mode)
            out = equilib_b200.equi2equi(img, rots, mode=mode, exact_clip=True)
            stats(f"equi2equi/{kind}/{mode}", out, ref)
            ref = equilib.equi2pers(img.clone(), rots, height=480, width=640, fov_x=90.0, mode=mode)
            out = equilib_b200.equi2pers(img, rots, 480, 640, 90.0, mode=mode, exact_clip=True)
            stats(f"equi2pers/{kind}/{mode}", out, ref)
            ref = equilib.equi2cube(img.clone(), rots, w_face=h // 2, cube_format="dice", mode=mode)
            out = equilib_b200.equi2cube(img, rots, h // 2, "dice", mode=mode, exact_clip=True)
            stats(f"equi2cube/{kind}/{mode}", out, ref)
            cube = ref
            ref2 = equilib.cube2equi(cube.clone(), "dice", height=h, width=w, mode=mode)
            out2 = equilib_b200.cube2equi(cube, "dice", h, w, mode=mode, exact_clip=True)
            stats(f"cube2equi/{kind}/{mode}", out2, ref2)
            pers = img[:, :, : h // 2, : w // 2].contiguous()
            ref = equilib.pers2equi(pers.clone(), rots, height=h // 2, width=w // 2, fov_x=90.0,
                                    mode=mode)
            out = equilib_b200.pers2equi(pers, rots, h // 2, w // 2, 90.0, mode=mode)
            stats(f"pers2equi/{kind}/{mode}", out, ref)
    # uint8 and fp16
    u8 = torch.from_numpy(cases.noise_image((2, 3, h, w), "uint8", 1)).to(dev)
    for mode in ("nearest", "bilinear", "bicubic"):
        ref = equilib.equi2equi(u8.clone(), rots, mode=mode)
        out = equilib_b200.equi2equi(u8, rots, mode=mode)
        d = (out.int() - ref.int()).abs()
        d = torch.minimum(d, 256 - d)
        print(json.dumps({"case": f"equi2equi/u8/{mode}", "max": int(d.max()),
                          "frac_equal": float((d == 0).float().mean())}), flush=True)
    # channels-last uint8 against the reference fed the planar frames
    for ch in (3, 4):
        x = torch.from_numpy(cases.noise_image((2, ch, h, w), "uint8", 2)).to(dev)
        ref = equilib.equi2equi(x.clone(), rots, mode="bilinear")
        out = equilib_b200.equi2equi(x.contiguous(memory_format=torch.channels_last), rots)
        d = (out.int() - ref.int()).abs()
        print(json.dumps({"case": f"equi2equi/u8 channels-last C={ch}/bilinear", "max": int(d.max()),
                          "frac_equal": float((d == 0).float().mean()),
                          "channels_last_out": bool(out.is_contiguous(
                              memory_format=torch.channels_last))}), flush=True)
    # bf16 (rejected by the reference): against the reference's fp32 result rounded once
    bf = imgs["band"].bfloat16()
    ref32 = equilib.equi2equi(bf.float(), rots, mode="bilinear")
    out = equilib_b200.equi2equi(bf, rots, mode="bilinear")
    stats("equi2equi/bf16 vs reference-fp32 rounded to bf16", out, ref32.bfloat16())
    f16 = imgs["band"].half()
    ref16 = equilib.equi2equi(f16.clone(), rots, mode="bilinear")
    ref32 = equilib.equi2equi(imgs["band"].clone(), rots, mode="bilinear")
    out16 = equilib_b200.equi2equi(f16, rots, mode="bilinear")
    stats("equi2equi/f16 vs reference-fp16 (grid quantised to fp16)", out16, ref16)
    stats("equi2equi/f16 vs reference-fp32", out16, ref32)
    stats("reference-fp16 vs reference-fp32", ref16, ref32)

    # timing of the reference's own CUDA path for context (end to end, wall clock)
    img = imgs["band"]
    for name, fn in (("reference", lambda: equilib.equi2equi(img, rots, mode="bilinear")),
                     ("equilib_b200", lambda: equilib_b200.equi2equi(img, rots, mode="bilinear"))):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(json.dumps({"timing": name, "ms_per_call": dt * 1e3,
                          "mpix_s": 2 * h * w / dt / 1e6, "shape": [2, 3, h, w]}), flush=True)


if __name__ == "__main__":
    main()
