#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (equilib 0.6.0).

Run in the build container, where the reference checkout exists:

    python tests/golden/make_golden.py [--ref /root/reference] [--device cpu|cuda]

Each case stores the seeded inputs' recipe (not the pixels: inputs are
regenerated from ``tests/cases.py``), the reference's sampling grid
(``convert_grid`` output, captured at the ``torch_grid_sample`` call site, e.g.
equilib/equi2equi/torch.py:183) and the reference's output tensor.

``--device cuda`` is for a one-off run on a GPU box with the reference shipped
under ``baseline/_ref`` (git-ignored): it records the reference's own CUDA path
(CPU grid -> H2D -> CUDA ``F.grid_sample``), which is the parity target of the
product.  Files are then named ``*_cuda.npz``.
"""

import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import cases  # noqa: E402


def _patch_capture(mods, cap):
    for mod in mods:
        orig = mod.torch_grid_sample

        def wrapped(img, grid, out=None, mode="bilinear", backend="native", _orig=orig):
            cap["grid"] = grid.detach().float().cpu().clone()
            return _orig(img=img, grid=grid, out=out, mode=mode, backend=backend)

        mod.torch_grid_sample = wrapped


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--out", default=HERE)
    args = ap.parse_args()
    sys.path.insert(0, args.ref)

    import equilib
    from equilib.cube2equi import torch as c2e_t
    from equilib.equi2cube import torch as e2c_t
    from equilib.equi2equi import torch as e2e_t
    from equilib.equi2pers import torch as e2p_t
    from equilib.pers2equi import torch as p2e_t

    assert equilib.__version__ == "0.6.0", equilib.__version__
    cap = {}
    _patch_capture((e2e_t, e2p_t, e2c_t, c2e_t, p2e_t), cap)
    funcs = {
        "equi2equi": equilib.equi2equi,
        "equi2pers": equilib.equi2pers,
        "equi2cube": equilib.equi2cube,
        "cube2equi": equilib.cube2equi,
        "pers2equi": equilib.pers2equi,
    }
    dev = torch.device(args.device)
    suffix = "" if dev.type == "cpu" else "_cuda"

    store = {}
    for case in cases.golden_cases():
        if dev.type == "cpu" and case["dtype"] == "float16":
            continue  # the reference rejects fp16 on CPU (equi2equi/torch.py:109-114)
        img, kwargs = cases.build_inputs(case)
        img_t = cases.to_device(img, dev)
        out = funcs[case["transform"]](img_t, **kwargs)
        out = cases.to_numpy(out)
        name = case["name"]
        if isinstance(out, np.ndarray):
            store[name + "/out"] = out
        else:
            raise TypeError(type(out))
        g = cap["grid"].numpy()
        if case["transform"] == "cube2equi":
            g = g[:1]  # identical for every batch item
        store[name + "/grid"] = g.astype(np.float32)
    path = os.path.join(args.out, f"reference_outputs{suffix}.npz")
    np.savez_compressed(path, **store)
    print(f"wrote {path}: {len(store)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")

    # rotation matrices: the reference's create_rotation_matrices on a seeded sweep
    from equilib.torch_utils import create_rotation_matrices

    rots = cases.rotation_sweep(257)
    R0 = create_rotation_matrices(rots, z_down=False).numpy()
    R1 = create_rotation_matrices(rots, z_down=True).numpy()
    if dev.type == "cpu":
        rp = os.path.join(args.out, "reference_rotations.npz")
        np.savez_compressed(rp, z_up=R0, z_down=R1)
        print(f"wrote {rp}")


if __name__ == "__main__":
    main()
