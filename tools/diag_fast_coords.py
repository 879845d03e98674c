#!/usr/bin/env python3
"""Where is the fast-coordinate error largest?  (GPU box diagnostic.)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import equilib_b200

DEV = torch.device("cuda", 0)
ROTS = [{"roll": 0.0, "pitch": 0.0, "yaw": 0.0}, {"roll": 0.3, "pitch": -0.4, "yaw": 1.1},
        {"roll": 1.5, "pitch": 1.2, "yaw": -2.9}, {"roll": -0.7, "pitch": 1.5707, "yaw": 3.1}]
for h, w in ((512, 1024), (2048, 4096)):
    ys = torch.arange(h, dtype=torch.float32, device=DEV)[:, None].expand(h, w)
    xs = torch.arange(w, dtype=torch.float32, device=DEV)[None, :].expand(h, w)
    img = torch.stack([ys, xs])[None].expand(len(ROTS), -1, -1, -1).contiguous()
    for name, fn, args in (("equi2equi", equilib_b200.equi2equi, ()),
                           ("equi2pers", equilib_b200.equi2pers, (h // 2, h // 2 * 2, 90.0)),
                           ("equi2cube", equilib_b200.equi2cube, (h // 4, "horizon"))):
        ex = fn(img, ROTS, *args, clip_output=False, fast_coords=False)
        fa = fn(img, ROTS, *args, clip_output=False, fast_coords=True)
        for ch, nm in ((0, "y"), (1, "x")):
            d = (fa[:, ch] - ex[:, ch]).abs()
            if ch == 1:
                seam = (ex[:, 1] > w - 1.5) | (ex[:, 1] < 0.5) | (fa[:, 1] > w - 1.5) | (fa[:, 1] < 0.5)
                d = torch.where(seam, torch.zeros_like(d), d)
            k = int(d.argmax())
            b, r = divmod(k, d.shape[1] * d.shape[2])
            yy, xx = divmod(r, d.shape[2])
            print(f"{name} {h}x{w} d{nm}: max {float(d.max()):.3e} mean {float(d.mean()):.3e} "
                  f"p99.99 {float(d.flatten().float().kthvalue(int(d.numel()*0.9999)).values):.3e} at b={b} y={yy} x={xx} "
                  f"exact=({float(ex[b,0,yy,xx]):.3f},{float(ex[b,1,yy,xx]):.3f}) "
                  f"fast=({float(fa[b,0,yy,xx]):.3f},{float(fa[b,1,yy,xx]):.3f})")
