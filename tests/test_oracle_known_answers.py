"""Known-answer vectors the reference's own test-suite holds for the path
(SURVEY.md 8c), restated against the oracle."""

import math
import os
import sys

import numpy as np
import pytest
import torch

from oracle import remap_oracle as O
from tests import cases, helpers

# tests/test_torch_utils/test_rotation.py:23-43 (Casio keisan ground truth, z_down=True)
GT = [
    ({"roll": 0.0, "pitch": 0.0, "yaw": 0.0}, [1.0, 0.0, 1.0], [1.0, 0.0, 1.0]),
    ({"roll": 0.0, "pitch": math.pi / 4, "yaw": math.pi / 4}, [1.0, 0.0, 0.0],
     [0.5, 0.5, -0.707106781186547524400844362105]),
    ({"roll": 0.0, "pitch": math.pi / 3, "yaw": -math.pi / 3}, [3.0, 2.0, 1.0],
     [2.91506350946109661690930792688, -1.04903810567665797014558475613,
      -2.09807621135331594029116951226]),
]


@pytest.mark.parametrize("rot,vin,vout", GT)
def test_rotation_known_answers(rot, vin, vout):
    R = O.rotation_matrix(**rot, z_down=True)
    got = R @ torch.tensor(vin, dtype=torch.float32)
    assert torch.allclose(got, torch.tensor(vout, dtype=torch.float32))


def test_rotation_z_down_flips_pitch_yaw():
    # tests/test_torch_utils/test_rotation.py:102-125
    r = {"roll": 0.3, "pitch": 0.2, "yaw": -0.7}
    a = O.rotation_matrix(**r, z_down=False)
    b = O.rotation_matrix(roll=0.3, pitch=-0.2, yaw=0.7, z_down=True)
    assert torch.equal(a, b)


def test_rotation_sweep_matches_reference_fixture():
    gold = helpers.load_golden("reference_rotations.npz")
    rots = cases.rotation_sweep(257)
    assert np.array_equal(O.rotation_matrices(rots, False).numpy(), gold["z_up"])
    assert np.array_equal(O.rotation_matrices(rots, True).numpy(), gold["z_down"])


@pytest.mark.parametrize("tail", [O._tail_plus_half, O._tail_minus_half])
def test_latitude_clamps_at_poles(tail):
    # tests/test_issue31_pole.py:24-49
    h = w = 64
    phis = torch.linspace(-math.pi / 2 + 1e-4, math.pi / 2 - 1e-4, 400)
    uj, ui = tail(phis, torch.zeros_like(phis), h, w)
    assert uj.min() >= 0 and uj.max() <= h - 1
    assert float(torch.diff(uj).abs().max()) < 2.0
    assert ui.min() >= 0 and ui.max() <= w


def test_identity_rotation_is_half_pixel_shift():
    # SURVEY appendix A / probe B6
    uj, ui = O.equi2equi_coords([{"roll": 0.0, "pitch": 0.0, "yaw": 0.0}], True, 8, 16, 8, 16)
    ys = torch.arange(8, dtype=torch.float32)[:, None].expand(8, 16)
    xs = torch.arange(16, dtype=torch.float32)[None, :].expand(8, 16)
    assert torch.allclose(uj[0, 1:-1], ys[1:-1] + 0.5, atol=1e-4)
    assert torch.allclose(ui[0, 1:], xs[1:] + 0.5, atol=1e-4)
    assert float(uj[0, -1].max()) == 7.0  # clamped to H-1


def test_native_has_no_seam_interpolation():
    # probe B4: x=7.5 on an 8-wide ramp reads column 7, a wrap would give 3.5
    img = torch.arange(8, dtype=torch.float32)[None, None, None, :].expand(1, 1, 4, 8)
    uj = torch.full((1, 1, 1), 1.0)
    ui = torch.full((1, 1, 1), 7.5)
    for flavour in ("cpu", "cuda"):
        out = O.sample(img.contiguous(), uj, ui, "bilinear", flavour)
        assert float(out) == 7.0


def test_nearest_rounds_half_to_even():
    img = torch.arange(8, dtype=torch.float32)[None, None, None, :].expand(1, 1, 4, 8)
    ui = torch.tensor([0.5, 1.5, 2.5, 3.5])[None, None, :]
    uj = torch.zeros_like(ui)
    out = O.sample(img.contiguous(), uj, ui, "nearest", "cpu")
    assert out.flatten().tolist() == [0.0, 2.0, 2.0, 4.0]


def test_uint8_cast_truncates_and_wraps():
    # probe B7
    v = torch.tensor([0.99, 1.5, 254.999, 255.6, 300.2, -0.4, -3.7])
    got = O._finish(v, v, torch.uint8, True)
    assert got.tolist() == [0, 1, 254, 255, 44, 0, 253]
    sat = O._finish(v, v, torch.uint8, True, saturate_uint8=True)
    assert sat.tolist() == [0, 1, 254, 255, 255, 0, 0]


def test_pers2equi_outside_fov_is_min_when_clipped():
    # probe B5
    pers = torch.rand(1, 3, 24, 32) + 0.5
    rots = [{"roll": 0.0, "pitch": 0.0, "yaw": 0.0}]
    import oracle

    a = oracle.pers2equi(pers, rots, 32, 64, 90.0, clip_output=True)
    b = oracle.pers2equi(pers, rots, 32, 64, 90.0, clip_output=False)
    assert float((b == 0).float().mean()) > 0.5
    assert torch.all(a[b == 0] == pers.min())


REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_live_reference_4k_frame_bilinear():
    """One full-size frame through the real reference (CPU) vs the oracle."""
    sys.path.insert(0, REF)
    try:
        import equilib
    finally:
        sys.path.remove(REF)
    img = cases.band_limited_image((1, 3, 512, 1024), "float32")
    rots = cases.bench_rotations(3)[2:]
    ref = equilib.equi2equi(torch.from_numpy(img), rots, mode="bilinear")
    import oracle

    out = oracle.equi2equi(torch.from_numpy(img), rots, mode="bilinear", flavour="cpu")
    assert helpers.frac_close(out.numpy(), ref.numpy(), 1e-6, 1e-6) > 0.9999
