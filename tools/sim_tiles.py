#!/usr/bin/env python3
"""CPU model of the staged kernels' tile classification (no GPU): for a set of rotations,
which share of the output tiles takes the fast path / the exact path from a box / global
gathers, and how many source bytes the TMA boxes fetch per output pixel -- as a function of
tile shape, buffer capacity and box table.  Uses the oracle's coordinate chain (design aid,
not product code).

    python tools/sim_tiles.py [--rots bench|random] [--n 6]
"""
import argparse
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import remap_oracle as O  # noqa: E402

H, W = 2048, 4096


def rots_of(kind, n):
    if kind == "bench":
        idx = np.linspace(0, 31, n).astype(int)
        return [{"roll": 0.01 * i, "pitch": 0.4 * math.sin(i), "yaw": 0.05 * i} for i in idx]
    rng = np.random.default_rng(0)
    return [{"roll": float(rng.uniform(-math.pi, math.pi)),
             "pitch": float(math.asin(rng.uniform(-1, 1))),
             "yaw": float(rng.uniform(-math.pi, math.pi))} for _ in range(n)]


def classify(ui, uj, tw, th, boxes, elt=2, nc=3, margin=2):
    """ui, uj: (H, W).  boxes: list of (w, h).  Returns dict of shares and fetched px / out px."""
    per16 = 16 // elt
    ty, tx = H // th, W // tw
    u = ui.reshape(ty, th, tx, tw)
    v = uj.reshape(ty, th, tx, tw)
    umin, umax = u.min(axis=(1, 3)), u.max(axis=(1, 3))
    vmin, vmax = v.min(axis=(1, 3)), v.max(axis=(1, 3))
    straddle = (umax - umin) >= 0.5 * W
    border = (umin < 1) | (umax > W - 3) | (vmin < 1) | (vmax > H - 3)
    # smoothness along x: 4th / 2nd differences of nodes every 8 px (interior approximation)
    nodes_u = ui[:, ::8]
    nodes_v = uj[:, ::8]
    def d4(a):
        p = np.pad(a, ((0, 0), (2, 2)), mode="edge")
        return np.abs(p[:, :-4] - 4 * p[:, 1:-3] + 6 * p[:, 2:-2] - 4 * p[:, 3:-1] + p[:, 4:])
    def d2(a):
        p = np.pad(a, ((0, 0), (1, 1)), mode="edge")
        return np.abs(p[:, :-2] - 2 * p[:, 1:-1] + p[:, 2:])
    rough = (np.maximum(d4(nodes_u), d4(nodes_v)) >= 0.03) | (np.maximum(d2(nodes_u), d2(nodes_v)) >= 1.5)
    rough_t = rough.reshape(ty, th, tx, tw // 8).any(axis=(1, 3))
    bx0 = np.floor(umin).astype(int)
    bx1 = np.floor(umax).astype(int)
    by0 = np.floor(vmin).astype(int)
    by1 = np.floor(vmax).astype(int)
    gx0 = (bx0 - margin) & ~(per16 - 1)
    need_w = bx1 + 2 + margin - gx0
    need_h = by1 + 2 + margin - (by0 - margin)
    pick = np.full(need_w.shape, -1)
    fetched = np.zeros(need_w.shape)
    for m in reversed(range(len(boxes))):
        bw, bh = boxes[m]
        ok = (need_w <= bw) & (need_h <= bh)
        pick[ok] = m
        fetched[ok] = bw * bh
    boxed = (pick >= 0) & ~straddle
    fast = boxed & ~border & ~rough_t
    n = pick.size
    return {"fast": fast.sum() / n, "boxed_exact": (boxed & ~fast).sum() / n,
            "unboxed": (~boxed).sum() / n, "fetch_ratio": (fetched * boxed).sum() / (n * tw * th),
            "pick_hist": np.bincount(pick[boxed], minlength=len(boxes)) / n,
            "mask_unboxed": ~boxed, "need": (need_w, need_h)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rots", default="bench")
    ap.add_argument("--n", type=int, default=6)
    a = ap.parse_args()
    rots = rots_of(a.rots, a.n)
    cfgs = {
        "lean2 32x16 cap2048": (32, 16, [(48, 24), (48, 32), (48, 42), (64, 32), (56, 36), (80, 25), (72, 28), (88, 23), (96, 21), (40, 51), (32, 64)]),
    }
    def table(cap, pitches=(48, 56, 64, 72, 80, 96, 112, 128), short=38):
        t = [(p, min(short, cap // p)) for p in pitches[:3]]
        t += [(p, min(96, cap // p)) for p in pitches]
        return [b for b in t if b[1] >= 21]
    for cap in (2389, 3072, 4096):
        cfgs[f"lean3 32x32 cap{cap}"] = (32, 32, table(cap))
        cfgs[f"lean3 32x16 cap{cap}"] = (32, 16, table(cap))
    acc = {k: [] for k in cfgs}
    for r in rots:
        uj, ui = O.equi2equi_coords([r], False, H, W, H, W)
        ui, uj = ui[0].numpy(), uj[0].numpy()
        for k, (tw, th, boxes) in cfgs.items():
            acc[k].append(classify(ui, uj, tw, th, boxes))
    for k, rs in acc.items():
        f = lambda key: float(np.mean([r[key] for r in rs]))
        print(f"{k:24s} fast {f('fast'):.3f}  boxed-exact {f('boxed_exact'):.3f}  unboxed {f('unboxed'):.3f}"
              f"  fetched px / out px {f('fetch_ratio'):.2f}")
    # combined: 32x32 where it fits, else its two halves
    for cap in (2389, 3072, 4096):
        full, half = acc[f"lean3 32x32 cap{cap}"], acc[f"lean3 32x16 cap{cap}"]
        un, fetch = [], []
        for rf, rh in zip(full, half):
            mu = rf["mask_unboxed"]                      # (ty32, tx)
            hu = rh["mask_unboxed"].reshape(mu.shape[0], 2, mu.shape[1])
            un.append((mu[:, None, :] & hu).mean())
        print(f"lean3 combined cap{cap}: unboxed (neither tile nor half fits) {np.mean(un):.3f}; "
              f"tiles split into halves {np.mean([r['unboxed'] for r in full]):.3f}")


if __name__ == "__main__":
    main()
