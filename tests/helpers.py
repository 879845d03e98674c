"""Helpers shared by the CPU and GPU test files."""

from __future__ import annotations

import os

import numpy as np
import torch

import oracle
from oracle import remap_oracle as O
from tests import cases

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name="reference_outputs.npz"):
    return np.load(os.path.join(GOLDEN_DIR, name))


def oracle_run(case: dict, img, kw: dict, flavour: str):
    """Call the oracle's func API for a case built by cases.build_inputs."""
    t = case["transform"]
    kw = dict(kw)
    fn = getattr(oracle, t)
    if t == "cube2equi":
        return fn(torch.from_numpy(img), flavour=flavour, **kw)
    rots = kw.pop("rots")
    x = torch.from_numpy(img)
    if case.get("gray_batch"):
        pass  # (B, H, W) + list of rots -> batch of grayscale, handled by the API
    return fn(x, rots, flavour=flavour, **kw)


def oracle_coords(case: dict, kw: dict):
    """(uj, ui) of the case through the oracle's coordinate chain."""
    t = case["transform"]
    shape = tuple(case["shape"])
    z = kw.get("z_down", False)
    rots = kw.get("rots")
    if isinstance(rots, dict):
        rots = [rots]
    if t == "equi2equi":
        h, w = shape[-2:]
        ho, wo = kw.get("height") or h, kw.get("width") or w
        return O.equi2equi_coords(rots, z, ho, wo, h, w)
    if t == "equi2pers":
        h, w = shape[-2:]
        return O.equi2pers_coords(rots, z, kw["height"], kw["width"], kw["fov_x"],
                                  kw.get("skew", 0.0), h, w)
    if t == "equi2cube":
        h, w = shape[-2:]
        return O.equi2cube_coords(rots, z, kw["w_face"], h, w)
    if t == "cube2equi":
        w_face = shape[-2] // 3 if kw["cube_format"] == "dice" else shape[-2]
        return O.cube2equi_coords(kw["height"], kw["width"], w_face)
    if t == "pers2equi":
        hp, wp = shape[-2:]
        return O.pers2equi_coords(rots, z, kw["height"], kw["width"], hp, wp, kw["fov_x"],
                                  kw.get("skew", 0.0))
    raise ValueError(t)


def frac_close(a, b, rtol, atol):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.mean(np.abs(a - b) <= atol + rtol * np.abs(b)))
