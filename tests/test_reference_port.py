"""The CPU-baseline port (oracle/reference_port.py) computes what the oracle does."""

import numpy as np
import pytest
import torch

import oracle
from oracle.reference_port import Equi2EquiCPU
from tests import cases, helpers


@pytest.mark.parametrize("mode", ["nearest", "bilinear", "bicubic"])
def test_port_matches_oracle(mode):
    img = torch.from_numpy(cases.noise_image((2, 3, 32, 64), "float32", 9))
    rots = cases.rotation_sweep(5)[3:]
    got = Equi2EquiCPU(mode=mode)(img, rots)
    ref = oracle.equi2equi(img, rots, mode=mode, flavour="cpu")
    if mode == "nearest":
        assert float((got == ref).float().mean()) >= 0.999
    else:
        assert helpers.frac_close(got.numpy(), ref.numpy(), 1e-5, 2e-5) >= 0.999


def test_port_matches_golden_reference_output():
    gold = helpers.load_golden()
    case = [c for c in cases.golden_cases() if c["name"] == "e2e_f32_bilinear"][0]
    img, kw = cases.build_inputs(case)
    got = Equi2EquiCPU(mode="bilinear")(torch.from_numpy(img), kw["rots"])
    assert helpers.frac_close(got.numpy(), gold["e2e_f32_bilinear/out"], 1e-5, 2e-5) >= 0.999
