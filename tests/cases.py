"""Seeded synthetic inputs shared by the golden generator, the oracle tests and
the GPU parity tests (SURVEY.md 8d "value distributions")."""

from __future__ import annotations

import math
from typing import Dict, List

import numpy as np
import torch

NP_DTYPES = {"float32": np.float32, "uint8": np.uint8, "float16": np.float16}
TORCH_DTYPES = {
    "float32": torch.float32,
    "uint8": torch.uint8,
    "float16": torch.float16,
    "bfloat16": torch.bfloat16,
}


def rotation_sweep(n: int, seed: int = 7) -> List[Dict[str, float]]:
    """n rotations: a few hand-picked corner cases, then a seeded random sweep."""
    fixed = [
        {"roll": 0.0, "pitch": 0.0, "yaw": 0.0},
        {"roll": 0.0, "pitch": math.pi / 4, "yaw": math.pi / 4},
        {"roll": 0.0, "pitch": math.pi / 3, "yaw": -math.pi / 3},
        {"roll": 0.1, "pitch": 0.3, "yaw": -0.5},
        {"roll": math.pi / 2, "pitch": -math.pi / 2, "yaw": math.pi},
    ]
    rng = np.random.default_rng(seed)
    out = list(fixed[:n])
    while len(out) < n:
        out.append(
            {
                "roll": float(rng.uniform(-math.pi, math.pi)),
                "pitch": float(rng.uniform(-math.pi / 2, math.pi / 2)),
                "yaw": float(rng.uniform(-math.pi, math.pi)),
            }
        )
    return out


def bench_rotations(n: int) -> List[Dict[str, float]]:
    """Distinct smooth rotations (the benchmarks/cache_benchmark.py:37-40 pattern)."""
    return [
        {"roll": 0.01 * i, "pitch": 0.4 * math.sin(i), "yaw": 0.05 * i} for i in range(n)
    ]


def noise_image(shape, dtype: str, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if dtype == "uint8":
        return rng.integers(0, 256, size=shape, dtype=np.uint8)
    return rng.random(shape, dtype=np.float32).astype(NP_DTYPES.get(dtype, np.float32))


def band_limited_image(shape, dtype: str, seed: int = 0, px: float = 512.0,
                       py: float = 384.0) -> np.ndarray:
    """0.5 + 0.5*sin(2 pi x/px + a_c)*cos(2 pi y/py + b_c): the parity image."""
    b, c, h, w = shape
    rng = np.random.default_rng(seed)
    pa = rng.uniform(0, 2 * math.pi, size=(b, c, 1, 1))
    pb = rng.uniform(0, 2 * math.pi, size=(b, c, 1, 1))
    x = np.arange(w, dtype=np.float64)[None, None, None, :]
    y = np.arange(h, dtype=np.float64)[None, None, :, None]
    v = 0.5 + 0.5 * np.sin(2 * math.pi * x / px + pa) * np.cos(2 * math.pi * y / py + pb)
    if dtype == "uint8":
        return np.clip(np.round(v * 255.0), 0, 255).astype(np.uint8)
    return v.astype(np.float32)


def to_torch(img: np.ndarray, dtype: str) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(img))
    return t.to(TORCH_DTYPES[dtype])


def to_device(obj, dev):
    if isinstance(obj, np.ndarray):
        return torch.from_numpy(obj).to(dev)
    if isinstance(obj, torch.Tensor):
        return obj.to(dev)
    if isinstance(obj, list):
        return [to_device(o, dev) for o in obj]
    if isinstance(obj, dict):
        return {k: to_device(v, dev) for k, v in obj.items()}
    raise TypeError(type(obj))


def to_numpy(obj):
    if isinstance(obj, torch.Tensor):
        t = obj.detach().cpu()
        if t.dtype in (torch.float16, torch.bfloat16):
            t = t.float()
        return t.numpy()
    return obj


def _case(name, transform, dtype="float32", mode="bilinear", **kw):
    d = dict(name=name, transform=transform, dtype=dtype, mode=mode)
    d.update(kw)
    return d


def golden_cases() -> List[dict]:
    """Small cases whose reference outputs are stored in tests/golden/."""
    cs = []
    modes = ("nearest", "bilinear", "bicubic")
    for m in modes:
        cs.append(_case(f"e2e_f32_{m}", "equi2equi", mode=m, shape=(2, 3, 32, 64)))
        cs.append(_case(f"e2e_u8_{m}", "equi2equi", "uint8", m, shape=(2, 3, 32, 64)))
        cs.append(_case(f"e2p_f32_{m}", "equi2pers", mode=m, shape=(2, 3, 32, 64),
                        height=24, width=32, fov_x=90.0))
        cs.append(_case(f"e2c_f32_{m}", "equi2cube", mode=m, shape=(2, 3, 32, 64),
                        w_face=16, cube_format="horizon"))
        cs.append(_case(f"c2e_f32_{m}", "cube2equi", mode=m, shape=(2, 3, 16, 96),
                        cube_format="horizon", height=32, width=64))
        cs.append(_case(f"p2e_f32_{m}", "pers2equi", mode=m, shape=(2, 3, 24, 32),
                        height=32, width=64, fov_x=90.0))
    cs.append(_case("e2e_f32_zdown", "equi2equi", shape=(2, 3, 32, 64), z_down=True))
    cs.append(_case("e2e_f32_resize", "equi2equi", shape=(2, 3, 32, 64), height=24, width=48))
    cs.append(_case("e2e_f32_noclip_bicubic", "equi2equi", mode="bicubic", shape=(2, 3, 32, 64),
                    clip_output=False))
    cs.append(_case("e2e_f32_single", "equi2equi", shape=(3, 32, 64), single=True))
    cs.append(_case("e2e_f32_gray", "equi2equi", shape=(3, 32, 64), gray_batch=True))
    cs.append(_case("e2p_f32_skew", "equi2pers", shape=(2, 3, 32, 64), height=20, width=36,
                    fov_x=60.0, skew=0.1, z_down=True))
    cs.append(_case("e2p_u8_bicubic", "equi2pers", "uint8", "bicubic", shape=(2, 3, 32, 64),
                    height=24, width=32, fov_x=100.0))
    cs.append(_case("e2c_f32_dice", "equi2cube", shape=(2, 3, 32, 64), w_face=16,
                    cube_format="dice"))
    cs.append(_case("e2c_f32_zdown", "equi2cube", shape=(2, 3, 32, 64), w_face=16,
                    cube_format="horizon", z_down=True))
    cs.append(_case("e2c_u8_dice", "equi2cube", "uint8", "bilinear", shape=(2, 3, 32, 64),
                    w_face=8, cube_format="dice"))
    cs.append(_case("c2e_f32_dice", "cube2equi", shape=(2, 3, 48, 64), cube_format="dice",
                    height=32, width=64))
    cs.append(_case("c2e_u8_bilinear", "cube2equi", "uint8", shape=(2, 3, 16, 96),
                    cube_format="horizon", height=32, width=64))
    cs.append(_case("c2e_f32_tall", "cube2equi", shape=(1, 3, 16, 96), cube_format="horizon",
                    height=48, width=64))
    cs.append(_case("p2e_f32_noclip", "pers2equi", shape=(2, 3, 24, 32), height=32, width=64,
                    fov_x=90.0, clip_output=False))
    cs.append(_case("p2e_f32_offset", "pers2equi", shape=(2, 3, 24, 32), height=32, width=64,
                    fov_x=70.0, skew=0.05, z_down=True, offset=0.5))
    cs.append(_case("p2e_u8_bilinear", "pers2equi", "uint8", shape=(2, 3, 24, 32), height=32,
                    width=64, fov_x=90.0))
    # fp16 is only recorded by the --device cuda run
    cs.append(_case("e2e_f16_bilinear", "equi2equi", "float16", shape=(2, 3, 32, 64)))
    return cs


def build_inputs(case: dict):
    """-> (image-or-cubemap as numpy, kwargs for the public func API)."""
    shape = tuple(case["shape"])
    seed = abs(hash(case["name"])) % 1000 if False else sum(ord(ch) for ch in case["name"])
    img = noise_image(shape, case["dtype"], seed)
    if case.get("offset"):
        img = (img + np.float32(case["offset"])).astype(img.dtype)
    t = case["transform"]
    kw = dict(mode=case["mode"])
    for k in ("z_down", "clip_output", "height", "width", "fov_x", "skew", "w_face",
              "cube_format"):
        if k in case:
            kw[k] = case[k]
    if t != "cube2equi":
        if case.get("single"):
            kw["rots"] = rotation_sweep(4, seed)[3]
        else:
            n = shape[0]
            kw["rots"] = rotation_sweep(n + 3, seed)[3:]
    return img, kw
