#!/usr/bin/env python3
"""CPU model of the shared-memory wavefronts of the strip-node kernel's tap loads (design aid):
for real coordinates (oracle chain) and the kernel's box selection, how many wavefronts a tap
LDS.32 takes under different lane shapes and box pitches.  A wavefront serves one 32-bit word
per bank; lanes reading the same word share it.

    python tools/sim_banks.py [--rots bench|random] [--n 3]
"""
import argparse
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import remap_oracle as O  # noqa: E402

H, W = 2048, 4096
ELT = int(os.environ.get('ELT', '2'))
NC = 3
BUDGET = (12 if ELT == 2 else 18) * 1024 // ELT // NC  # pixels per channel in one box buffer


def boxes(widths):
    out = []
    for i, w in enumerate(widths):
        h = min(BUDGET // w, 128)
        if i == 0:
            h = min(h, 24)
        if i == 1:
            h = min(h, 32)
        out.append((w, h))
    return out


def wavefronts(words):
    """words: (n_instr, 32) int array of word addresses -> wavefronts per instruction."""
    banks = words % 32
    wf = np.zeros(len(words), dtype=np.int64)
    for i in range(len(words)):
        u = np.unique(words[i])
        wf[i] = np.bincount(u % 32, minlength=32).max()
    return wf


def tile_words(ui, uj, ty, tx, shape, pitch, gx0, gy0):
    """Word addresses of the 'left tap, upper row' load of every (warp, pixel slot) of tile
    (ty, tx) for a lane shape.  shape: '16x2' | '32x1' | '8x4'."""
    x0 = tx * 32
    y0 = ty * 16
    out = []
    for warp in range(4):
        for j in range(4):
            lanes = np.arange(32)
            if shape == "16x2":
                lx, ly = lanes & 15, lanes >> 4
                row = warp * 4 + ly + 2 * (j >> 1)
                col = lx + 16 * (j & 1)
            elif shape == "16x2 rows r,r+2":
                lx, ly = lanes & 15, lanes >> 4
                row = warp * 4 + 2 * ly + (j >> 1)
                col = lx + 16 * (j & 1)
            elif shape == "16x2 rows r,r+8":
                lx, ly = lanes & 15, lanes >> 4
                row = (warp * 2 + (j >> 1)) + 8 * ly
                col = lx + 16 * (j & 1)
            elif shape == "32x1":
                row = np.full(32, warp * 4 + j)
                col = lanes
            else:  # 8x4: 8 lanes per row, 4 rows; pixel slot j = column block
                lx, ly = lanes & 7, lanes >> 3
                row = warp * 4 + ly
                col = lx + 8 * j
            xs = np.floor(ui[y0 + row, x0 + col]).astype(int)
            ys = np.floor(uj[y0 + row, x0 + col]).astype(int)
            addr = ((ys - gy0) * pitch + (xs - gx0)) * ELT
            out.append(addr // 4)
    return np.array(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rots", default="bench")
    ap.add_argument("--n", type=int, default=3)
    ap.add_argument("--tiles", type=int, default=600)
    a = ap.parse_args()
    if a.rots == "bench":
        idx = np.linspace(3, 29, a.n).astype(int)
        rots = [{"roll": 0.01 * i, "pitch": 0.4 * math.sin(i), "yaw": 0.05 * i} for i in idx]
    else:
        rng = np.random.default_rng(0)
        rots = [{"roll": float(rng.uniform(-math.pi, math.pi)),
                 "pitch": float(math.asin(rng.uniform(-1, 1))),
                 "yaw": float(rng.uniform(-math.pi, math.pi))} for _ in range(a.n)]
    tables = {
        "shipped widths 48 48 48 64 56 80 72 88 96 40 32": boxes([48, 48, 48, 64, 56, 80, 72, 88, 96, 40, 32]),
        "by bank quality 48 48 48 88 80 96 56 72 64 40 32": boxes([48, 48, 48, 88, 80, 96, 56, 72, 64, 40, 32]),
        "quality + 104 112: 48 48 48 88 104 80 96 112 56 72 64 40 32": boxes([48, 48, 48, 88, 104, 80, 96, 112, 56, 72, 64, 40, 32]),
    }
    if ELT == 4:
        tables = {
            "shipped": boxes([48, 48, 48, 64, 56, 80, 72, 88, 96, 40, 32]),
            "fp32 quality 48 48 48 80 56 88 72 40 64 96 32": boxes([48, 48, 48, 80, 56, 88, 72, 40, 64, 96, 32]),
            "fp32 quality b 48 48 48 80 112 56 88 72 40 64 96": boxes([48, 48, 48, 80, 112, 56, 88, 72, 40, 64, 96]),
        }
    rng = np.random.default_rng(1)
    for tname, tab in tables.items():
        shapes = ("16x2",)
        res = {s: [] for s in shapes}
        per_box = {s: {} for s in shapes}
        used = np.zeros(len(tab), dtype=int)
        unboxed = 0
        total = 0
        for r in rots:
            uj, ui = O.equi2equi_coords([r], False, H, W, H, W)
            ui, uj = ui[0].numpy(), uj[0].numpy()
            for _ in range(a.tiles):
                ty, tx = int(rng.integers(0, H // 16)), int(rng.integers(0, W // 32))
                u = ui[ty * 16:ty * 16 + 16, tx * 32:tx * 32 + 32]
                v = uj[ty * 16:ty * 16 + 16, tx * 32:tx * 32 + 32]
                total += 1
                if u.max() - u.min() > W / 2:
                    unboxed += 1
                    continue
                gx0 = (int(np.floor(u.min())) - 2) & ~(16 // ELT - 1)
                gy0 = int(np.floor(v.min())) - 2
                nw = int(np.floor(u.max())) + 4 - gx0
                nh = int(np.floor(v.max())) + 4 - gy0
                m = next((i for i, (bw, bh) in enumerate(tab) if nw <= bw and nh <= bh), -1)
                if m < 0:
                    unboxed += 1
                    continue
                used[m] += 1
                for s in res:
                    v = wavefronts(tile_words(ui, uj, ty, tx, s, tab[m][0], gx0, gy0)).mean()
                    res[s].append(v)
                    per_box[s].setdefault(tab[m][0], []).append(v)
        print(f"{a.rots}: {tname}: unboxed {unboxed / total:.3f}; box use {np.round(used / max(used.sum(), 1), 2)}")
        for s, v in res.items():
            pb = ", ".join(f"{w}: {np.mean(x):.2f}" for w, x in sorted(per_box[s].items()))
            print(f"    lane shape {s}: {np.mean(v):.3f} wavefronts per tap LDS   (by pitch: {pb})")


if __name__ == "__main__":
    main()
