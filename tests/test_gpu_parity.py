"""GPU parity tests: the CUDA path (through the public API -> torch custom op ->
C ABI) against the CPU oracle and against the committed reference outputs.

Tolerances (stated once, used below):
* coordinates: |d| <= 1e-6 * size px.  Every fp32 step of the chain is mirrored
  exactly except asinf/atan2f, which differ from the host libm by a few ulp of an
  angle of magnitude <= pi, i.e. <= ~1e-6 rad = 1.6e-7 * size px.
* fp32 bilinear/bicubic on the band-limited parity image:
  rtol 1e-5, atol 1e-5 (BASELINE.json north_star), >= 99.9 % of pixels.
* fp32 on uniform-noise images (worst case, |grad| ~ 1 / px): the same coordinate
  noise becomes <= size * 1e-6 in value; bound 5e-4 absolute.
* nearest: identical except where a coordinate sits within the noise of a .5 tie:
  >= 99.5 % identical pixels.
* uint8: identical code values except truncation flips: |d| <= 1, >= 99 % identical
  (bicubic overshoot wraps exactly like the reference, compared modulo 256).
* fp16 / bf16: oracle = fp32 math on the up-cast input rounded once; 1 ulp of the
  output type (2^-10 / 2^-7 relative).
"""

import math

import numpy as np
import pytest
import torch

import equilib_b200
import oracle
from equilib_b200 import _lib, _ops, _prep
from oracle import remap_oracle as O
from tests import cases, helpers

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda", 0)
GOLD = helpers.load_golden()
# outputs of the unmodified reference's own CUDA path, recorded once on a B200 by
# tests/golden/make_golden.py --device cuda (the parity target)
GOLD_CUDA = helpers.load_golden("reference_outputs_cuda.npz")
GOLDEN_CASES = [c for c in cases.golden_cases()]
MODES = ("nearest", "bilinear", "bicubic")


def rate(x, lo):
    """x >= lo, with x appended to gpurun_out/parity_rates.jsonl under the running test's
    name -- the bounds below are set ~0.1-0.3 % under the values measured on a B200."""
    import json
    import os

    x = float(x)
    rec = {"test": os.environ.get("PYTEST_CURRENT_TEST", "?").split("::")[-1].split(" ")[0],
           "rate": round(x, 6), "bound": lo}
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                               "gpurun_out", "parity_rates.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    return x >= lo


def product_run(case, img, kw):
    t = case["transform"]
    kw = dict(kw)
    fn = getattr(equilib_b200, t)
    x = torch.from_numpy(img).to(DEV)
    if t == "cube2equi":
        return fn(x, **kw)
    rots = kw.pop("rots")
    return fn(x, rots, **kw)


def close_except_seam(o, r, tol, frac=0.9999):
    """|o - r| <= tol for all but a handful of elements.  The reference does not
    interpolate across the longitude seam (column W-1 vs columns 0/1), so a pixel whose
    longitude sits within the last-ulp noise of asin/atan2 of the seam may land on either
    side; such pixels are counted, not bounded."""
    d = np.abs(np.asarray(o, dtype=np.float64) - np.asarray(r, dtype=np.float64))
    return float(np.mean(d <= tol)) >= frac


def compare(case, out, ref):
    out = cases.to_numpy(out)
    ref = cases.to_numpy(ref)
    assert out.shape == ref.shape, (out.shape, ref.shape)
    assert out.dtype == ref.dtype
    if case["mode"] == "nearest":
        assert np.mean(out == ref) >= 0.995
    elif case["dtype"] == "uint8":
        d = np.abs(out.astype(np.int32) - ref.astype(np.int32))
        d = np.minimum(d, 256 - d)
        assert np.mean(d <= 1) >= 0.9995  # (seam pixels, see close_except_seam)
        assert rate(np.mean(d == 0), 0.994)
    else:
        assert close_except_seam(out, ref, 5e-4, 0.9995)
        assert helpers.frac_close(out, ref, rtol=1e-5, atol=2e-5) >= 0.99


@pytest.mark.parametrize("case", GOLDEN_CASES, ids=[c["name"] for c in GOLDEN_CASES])
def test_golden_case_vs_oracle_and_reference(case):
    img, kw = cases.build_inputs(case)
    out = product_run(case, img, kw)
    torch.cuda.synchronize()
    if case["dtype"] == "float16":
        # policy: fp32 coordinates, one rounding at the end; the
        # reference's own fp16 path quantises the grid to fp16 and is NOT the target
        ref = helpers.oracle_run(case, img.astype(np.float32), kw, flavour="cuda")
        o = cases.to_numpy(out)
        assert out.dtype == torch.float16
        assert close_except_seam(o, ref.numpy(), 2 ** -10 + 5e-4, 0.9995)
        return
    ref = helpers.oracle_run(case, img, kw, flavour="cuda")
    compare(case, out, ref)
    # the committed outputs of the unmodified reference: its CUDA path (the parity
    # target), and -- for the interpolating modes -- its CPU path, which differs by
    # one rounding in the normalisation step (enough to flip nearest-mode ties)
    compare(case, out, GOLD_CUDA[case["name"] + "/out"])
    if case["mode"] != "nearest":
        compare(case, out, GOLD[case["name"] + "/out"])


# ---------------------------------------------------------------------------
# coordinate-level parity through eqb_remap_coords
# ---------------------------------------------------------------------------


def circ_diff(a, b, period):
    d = np.abs(a - b)
    return np.minimum(d, np.abs(period - d))


def gpu_coords(transform, rots, z_down, dst_hw, src_hw, extra=None, w_face=0, flags=0):
    b = len(rots) if rots else 1
    dh, dw = dst_hw
    mats = tab_x = tab_y = itab_x = None
    if transform == _lib.EQUI2EQUI:
        mats = _prep.mats_for("rot", rots, z_down, DEV)
        tab_x, tab_y = _prep.equirect_tables(dh, dw, DEV)
    elif transform == _lib.EQUI2PERS:
        mats = _prep.mats_for("equi2pers", rots, z_down, DEV, extra)
    elif transform == _lib.PERS2EQUI:
        mats = _prep.mats_for("pers2equi", rots, z_down, DEV, extra)
        tab_x, tab_y = _prep.equirect_tables(dh, dw, DEV)
    elif transform == _lib.EQUI2CUBE:
        mats = _prep.mats_for("rot", rots, not z_down, DEV)
        tab_x = _prep.cube_rng_table(w_face, DEV)
    elif transform == _lib.CUBE2EQUI:
        tab_x, tab_y, itab_x = _prep.cube2equi_tables(dh, dw, DEV)
    out = torch.empty((b, 2, dh, dw), dtype=torch.float32, device=DEV)
    _ops.coords(out, mats, tab_x, tab_y, itab_x, None, transform, b, src_hw[0], src_hw[1], dh, dw,
                w_face=w_face, n_cells=6 if transform == _lib.EQUI2CUBE else 0, flags=flags)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("z_down", [False, True])
def test_coords_equi2equi(z_down):
    rots = cases.rotation_sweep(6)
    h, w = 256, 512
    g = gpu_coords(_lib.EQUI2EQUI, rots, z_down, (h, w), (h, w))
    uj, ui = O.equi2equi_coords(rots, z_down, h, w, h, w)
    duj = np.abs(g[:, 0] - uj.numpy())
    dui = circ_diff(g[:, 1], ui.numpy(), w)
    print(f"equi2equi coords: max duj {duj.max():.2e} dui {dui.max():.2e} "
          f"mean {duj.mean():.2e}/{dui.mean():.2e} exact {np.mean(duj == 0):.3f}/{np.mean(dui == 0):.3f}")
    assert duj.max() <= 1e-6 * h * 4 and dui.max() <= 1e-6 * w * 4
    assert duj.mean() <= 1e-6 * h and dui.mean() <= 1e-6 * w


def test_coords_equi2pers():
    rots = cases.rotation_sweep(5)
    g = gpu_coords(_lib.EQUI2PERS, rots, False, (120, 160), (256, 512), extra=(120, 160, 90.0, 0.0))
    uj, ui = O.equi2pers_coords(rots, False, 120, 160, 90.0, 0.0, 256, 512)
    assert np.abs(g[:, 0] - uj.numpy()).max() <= 4e-6 * 256
    assert circ_diff(g[:, 1], ui.numpy(), 512).max() <= 4e-6 * 512


def test_coords_equi2cube():
    rots = cases.rotation_sweep(5)
    g = gpu_coords(_lib.EQUI2CUBE, rots, False, (64, 384), (256, 512), w_face=64)
    uj, ui = O.equi2cube_coords(rots, False, 64, 256, 512)
    assert np.abs(g[:, 0] - uj.numpy()).max() <= 4e-6 * 256
    assert circ_diff(g[:, 1], ui.numpy(), 512).max() <= 4e-6 * 512


def test_coords_cube2equi_bit_exact():
    # no transcendental on the device: tables come from the host, the rest is
    # mirrored arithmetic -> identical bits
    g = gpu_coords(_lib.CUBE2EQUI, None, False, (128, 256), (64, 384), w_face=64)
    uj, ui = O.cube2equi_coords(128, 256, 64)
    assert np.array_equal(g[0, 0], uj[0].numpy())
    assert np.array_equal(g[0, 1], ui[0].numpy())


def test_coords_pers2equi_bit_exact():
    rots = cases.rotation_sweep(5)
    g = gpu_coords(_lib.PERS2EQUI, rots, False, (128, 256), (96, 128), extra=(96, 128, 90.0, 0.0))
    uj, ui = O.pers2equi_coords(rots, False, 128, 256, 96, 128, 90.0, 0.0)
    assert np.array_equal(g[:, 0], uj.numpy())
    assert np.array_equal(g[:, 1], ui.numpy())


def test_fast_math_paths_match_libm():
    """The inlined sqrt/div/atan2 fast paths give the same bits as the CUDA math
    library calls they replace (include/equilib_b200.h EQB_FLAG_LIBM)."""
    rots = cases.rotation_sweep(24)
    for z in (False, True):
        a = gpu_coords(_lib.EQUI2EQUI, rots, z, (512, 1024), (2048, 4096))
        b = gpu_coords(_lib.EQUI2EQUI, rots, z, (512, 1024), (2048, 4096), flags=_lib.FLAG_LIBM)
        assert np.array_equal(a, b)
    a = gpu_coords(_lib.EQUI2PERS, rots, False, (480, 640), (2048, 4096), extra=(480, 640, 90.0, 0.0))
    b = gpu_coords(_lib.EQUI2PERS, rots, False, (480, 640), (2048, 4096), extra=(480, 640, 90.0, 0.0),
                   flags=_lib.FLAG_LIBM)
    assert np.array_equal(a, b)
    a = gpu_coords(_lib.EQUI2CUBE, rots, False, (256, 1536), (2048, 4096), w_face=256)
    b = gpu_coords(_lib.EQUI2CUBE, rots, False, (256, 1536), (2048, 4096), w_face=256,
                   flags=_lib.FLAG_LIBM)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("shape", [1, 2, 3])
def test_warp_tile_shapes_are_equivalent(shape, monkeypatch):
    img = torch.from_numpy(cases.noise_image((2, 3, 96, 200), "float32", 17)).to(DEV)
    rots = cases.rotation_sweep(5)[3:]
    base = {m: equilib_b200.equi2equi(img, rots, mode=m, height=77, width=150) for m in MODES}
    monkeypatch.setenv("EQUILIB_B200_WARP_SHAPE", str(shape))
    for m in MODES:
        assert torch.equal(equilib_b200.equi2equi(img, rots, mode=m, height=77, width=150), base[m])
    u8 = torch.from_numpy(cases.noise_image((2, 3, 96, 200), "uint8", 18)).to(DEV)
    monkeypatch.delenv("EQUILIB_B200_WARP_SHAPE")
    ref = equilib_b200.equi2cube(u8, rots, 24, "dice")
    monkeypatch.setenv("EQUILIB_B200_WARP_SHAPE", str(shape))
    assert torch.equal(equilib_b200.equi2cube(u8, rots, 24, "dice"), ref)


@pytest.mark.parametrize("dtype", ["float32", "bfloat16", "uint8"])
def test_staged_and_direct_kernels_agree(dtype, monkeypatch):
    """Large bilinear launches take the shared-memory staged kernel (tiles whose source
    box does not fit fall back to direct gathers inside it); EQUILIB_B200_NO_STAGE=1
    forces the direct-gather kernel.  Same bits either way, poles and seam included."""
    img = torch.from_numpy(cases.noise_image((3, 3, 512, 1024), "uint8" if dtype == "uint8"
                                             else "float32", 5)).to(DEV)
    if dtype == "bfloat16":
        img = img.to(torch.bfloat16)
    rots = [{"roll": 0.0, "pitch": 0.0, "yaw": 0.0}, {"roll": 0.3, "pitch": -0.4, "yaw": 1.1},
            {"roll": 1.5, "pitch": 1.2, "yaw": -2.9}]
    def run_all():
        kw = dict(fast_coords=False)  # same (exact) coordinate chain on both paths
        return [
            equilib_b200.equi2equi(img, rots, **kw),
            equilib_b200.equi2pers(img, rots, 480, 640, 100.0, **kw),
            equilib_b200.equi2cube(img, rots, 256, "dice", **kw),
            equilib_b200.equi2cube(img, rots, 256, "horizon", **kw),
            equilib_b200.pers2equi(img, rots, 512, 1024, 90.0, **kw),
            equilib_b200.equi2equi(img[:, :, :, :1000], rots, height=500, width=1002, **kw),
            # bicubic: staged tiles read the 4x4 taps from the box, border tiles reflect
            equilib_b200.equi2equi(img, rots, mode="bicubic", **kw),
            equilib_b200.equi2pers(img, rots, 480, 640, 100.0, mode="bicubic", **kw),
            equilib_b200.equi2cube(img, rots, 256, "dice", mode="bicubic", clip_output=False, **kw),
            equilib_b200.pers2equi(img, rots, 512, 1024, 90.0, mode="bicubic", **kw),
            # cube2equi from a true horizon strip is a plain image for the box loads
            equilib_b200.cube2equi(img[:, :, :128, :768].contiguous(), "horizon", 512, 1024),
            equilib_b200.cube2equi(img[:, :, :128, :768].contiguous(), "horizon", 512, 1024,
                                   mode="bicubic", clip_output=False),
        ]
    staged = run_all()                      # TMA box loads
    monkeypatch.setenv("EQUILIB_B200_NO_TMA", "1")
    staged_cp_async = run_all()             # 16-byte cp.async row chunks
    monkeypatch.setenv("EQUILIB_B200_NO_STAGE", "1")
    direct = run_all()                      # global gathers
    for a, b, c in zip(staged, staged_cp_async, direct):
        assert torch.equal(a, c)
        assert torch.equal(b, c)


HARD_ROTS = [{"roll": 0.0, "pitch": 0.0, "yaw": 0.0}, {"roll": 0.3, "pitch": -0.4, "yaw": 1.1},
             {"roll": 1.5, "pitch": 1.2, "yaw": -2.9}, {"roll": -0.7, "pitch": 1.5707, "yaw": 3.1}]


def test_fast_coordinates_error_in_pixels():
    """Fast coordinates = exact chain on every 4th pixel + cubic interpolation along x.
    Sampling a ramp image (pixel value == its own x or y) bilinearly returns the
    sampling coordinate itself, so fast-vs-exact output differences ARE the coordinate
    error in pixels.  Bound: 2e-3 px max, 2e-4 px mean at 2K x 4K (an fp32 coordinate
    there resolves 2.4e-4 px; the reference's own fp32 chain is off by up to 0.23 px at
    the poles vs float64, SURVEY probe B3).  The error is dominated by the skipped
    normalise / un-normalise round trip (its own rounding noise), not the cubic."""
    h, w = 2048, 4096
    ys = torch.arange(h, dtype=torch.float32, device=DEV)[:, None].expand(h, w)
    xs = torch.arange(w, dtype=torch.float32, device=DEV)[None, :].expand(h, w)
    img = torch.stack([ys, xs])[None].expand(len(HARD_ROTS), -1, -1, -1).contiguous()
    for fn, args in ((equilib_b200.equi2equi, ()), (equilib_b200.equi2pers, (1024, 1024, 90.0)),
                     (equilib_b200.equi2cube, (512, "horizon"))):
        exact = fn(img, HARD_ROTS, *args, clip_output=False, fast_coords=False)
        fast = fn(img, HARD_ROTS, *args, clip_output=False, fast_coords=True)
        dy = (fast[:, 0] - exact[:, 0]).abs()
        dx = (fast[:, 1] - exact[:, 1]).abs()
        # at the seam the exact path blends columns W-1 and 0 of the ramp: skip it
        seam = (exact[:, 1] > w - 1.5) | (exact[:, 1] < 0.5) | (fast[:, 1] > w - 1.5)
        dx = torch.where(seam, torch.zeros_like(dx), dx)
        print(f"{fn.__name__}: coordinate error max dy {float(dy.max()):.2e} dx {float(dx.max()):.2e} "
              f"mean {float(dy.mean()):.2e}/{float(dx.mean()):.2e} px")
        assert float(dy.max()) <= 3e-3 and float(dx.max()) <= 3e-3
        assert float(dy.mean()) <= 2e-4 and float(dx.mean()) <= 2e-4


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.uint8])
def test_low_precision_images_at_staged_sizes(dtype):
    """16/8-bit images on the default path of large launches (staged kernel + fast
    coordinates) against the oracle policy: fp32 math on the up-cast input, one
    rounding at the end; 1 ulp of the output type (truncation flip for uint8)."""
    if dtype == torch.uint8:
        base = torch.from_numpy(cases.band_limited_image((2, 3, 512, 1024), "uint8", px=96, py=64))
    else:
        base = torch.from_numpy(cases.band_limited_image((2, 3, 512, 1024), "float32", px=96,
                                                         py=64)).to(dtype)
    rots = HARD_ROTS[1:3]
    out = equilib_b200.equi2equi(base.to(DEV), rots)
    ref = oracle.equi2equi(base, rots, flavour="cuda")
    assert out.dtype == dtype
    o, r = out.float().cpu().numpy(), ref.float().numpy()
    d = np.abs(o - r)
    if dtype == torch.uint8:
        assert np.mean(d <= 1) >= 0.9999 and rate(np.mean(d == 0), 0.994)
    else:
        ulp = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
        assert close_except_seam(o, r, ulp * 1.01 + 2e-4, 0.99999)
        assert rate(np.mean(d == 0), 0.994)
    exact = equilib_b200.equi2equi(base.to(DEV), rots, fast_coords=False)
    same = float((exact == out).float().mean())
    print(f"{dtype}: fast == exact for {same:.5f} of outputs")
    assert rate(same, 0.994)


# ---------------------------------------------------------------------------
# sampler-only parity: same coordinates in, so only the sampler differs
# ---------------------------------------------------------------------------


@pytest.mark.parametrize("mode", MODES)
def test_grid_sample_matches_oracle_sampler(mode):
    rng = np.random.default_rng(5)
    img = rng.random((2, 3, 40, 56), dtype=np.float32)
    # includes the seam, the poles, out-of-range and exact-integer coordinates
    uj = rng.uniform(-1.5, 41.0, size=(2, 33, 47)).astype(np.float32)
    ui = rng.uniform(-1.5, 57.5, size=(2, 33, 47)).astype(np.float32)
    uj[:, 0, :8] = np.arange(8, dtype=np.float32)
    ui[:, 0, :8] = 55.0
    grid = np.stack([uj, ui], axis=1)
    out = equilib_b200.grid_sample(torch.from_numpy(img).to(DEV), torch.from_numpy(grid).to(DEV),
                                   mode=mode)
    ref = O.sample(torch.from_numpy(img), torch.from_numpy(uj), torch.from_numpy(ui), mode, "cuda")
    o, r = out.cpu().numpy(), ref.numpy()
    if mode == "bicubic":
        assert np.abs(o - r).max() <= 2e-6
    else:
        assert np.array_equal(o, r)


# ---------------------------------------------------------------------------
# the north-star tolerance on the band-limited parity image, realistic size
# ---------------------------------------------------------------------------


@pytest.mark.parametrize("mode", ["bilinear", "bicubic"])
def test_equi2equi_fp32_band_limited_1e5(mode):
    img = cases.band_limited_image((2, 3, 512, 1024), "float32")
    rots = cases.bench_rotations(5)[3:]
    out = equilib_b200.equi2equi(torch.from_numpy(img).to(DEV), rots, mode=mode)
    ref = oracle.equi2equi(torch.from_numpy(img), rots, mode=mode, flavour="cuda")
    frac = helpers.frac_close(out.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-5)
    print(f"equi2equi {mode} band-limited: frac within 1e-5 = {frac:.6f}, "
          f"max abs {np.abs(out.cpu().numpy() - ref.numpy()).max():.2e}")
    assert frac >= 0.999


def test_equi2pers_fp32_band_limited_1e5():
    img = cases.band_limited_image((1, 3, 512, 1024), "float32")
    rot = {"roll": 0.1, "pitch": 0.3, "yaw": -0.5}
    out = equilib_b200.equi2pers(torch.from_numpy(img[0]).to(DEV), rot, 240, 320, 90.0)
    ref = oracle.equi2pers(torch.from_numpy(img[0]), rot, 240, 320, 90.0)
    assert out.shape == (3, 240, 320)
    assert helpers.frac_close(out.cpu().numpy(), ref.numpy(), 1e-5, 1e-5) >= 0.999


# ---------------------------------------------------------------------------
# dtypes
# ---------------------------------------------------------------------------


@pytest.mark.parametrize("dtype,ulp", [(torch.float16, 2.0 ** -10), (torch.bfloat16, 2.0 ** -7)])
@pytest.mark.parametrize("mode", MODES)
def test_16bit_inputs(dtype, ulp, mode):
    img = torch.from_numpy(cases.band_limited_image((2, 3, 128, 256), "float32", px=96, py=64))
    img = img.to(dtype)
    rots = cases.rotation_sweep(5)[3:]
    out = equilib_b200.equi2equi(img.to(DEV), rots, mode=mode)
    assert out.dtype == dtype
    ref = oracle.equi2equi(img, rots, mode=mode, flavour="cuda")
    assert ref.dtype == dtype
    o, r = out.float().cpu().numpy(), ref.float().numpy()
    if mode == "nearest":
        assert rate(np.mean(o == r), 0.995)
    else:
        assert close_except_seam(o, r, ulp * 1.01 + 2e-4)
        assert rate(np.mean(o == r), 0.997)


@pytest.mark.parametrize("odd", [False, True])
def test_unaligned_output_rows_take_the_scalar_store_path(odd):
    # odd widths / offset views break the vector-store alignment rules
    w_out = 37 if odd else 40
    img = cases.noise_image((2, 3, 32, 64), "uint8", 3)
    rots = cases.rotation_sweep(5)[3:]
    out = equilib_b200.equi2pers(torch.from_numpy(img).to(DEV), rots, 21, w_out, 80.0)
    ref = oracle.equi2pers(torch.from_numpy(img), rots, 21, w_out, 80.0)
    d = np.abs(out.cpu().numpy().astype(int) - ref.numpy().astype(int))
    assert np.mean(d <= 1) >= 0.9999 and rate(np.mean(d == 0), 0.99)


def test_strided_and_expanded_sources():
    base = torch.from_numpy(cases.noise_image((1, 3, 32, 64), "float32", 11)).to(DEV)
    rots = cases.rotation_sweep(7)[3:]
    expanded = base.expand(4, -1, -1, -1)  # stride-0 batch: one equirect, many views
    a = equilib_b200.equi2pers(expanded, rots, 24, 32, 90.0)
    b = equilib_b200.equi2pers(base.repeat(4, 1, 1, 1), rots, 24, 32, 90.0)
    assert torch.equal(a, b)
    wide = torch.zeros((4, 3, 40, 80), device=DEV)
    wide[:, :, 4:36, 8:72] = base
    c = equilib_b200.equi2pers(wide[:, :, 4:36, 8:72], rots, 24, 32, 90.0)
    assert torch.equal(a, c)


def test_class_api_is_bitwise_equal_to_func_api():
    # tests/test_cache_consistency.py:88-118 of the reference
    img = torch.from_numpy(cases.noise_image((2, 3, 32, 64), "float32", 42)).to(DEV)
    for rots in (cases.rotation_sweep(5)[3:], cases.rotation_sweep(9)[7:]):
        e2p = equilib_b200.Equi2Pers(24, 32, 90.0)
        for _ in range(2):
            assert torch.equal(e2p(img, rots), equilib_b200.equi2pers(img, rots, 24, 32, 90.0))
        e2e = equilib_b200.Equi2Equi()
        assert torch.equal(e2e(img, rots), equilib_b200.equi2equi(img, rots))
        e2c = equilib_b200.Equi2Cube(16, "dice")
        assert torch.equal(e2c(img, rots), equilib_b200.equi2cube(img, rots, 16, "dice"))
        p2e = equilib_b200.Pers2Equi(32, 64)
        assert torch.equal(p2e(img, rots, 90.0), equilib_b200.pers2equi(img, rots, 32, 64, 90.0))


def test_equi2equi_class_ignores_ctor_size_like_the_reference():
    img = torch.zeros((3, 64, 128), device=DEV)
    out = equilib_b200.Equi2Equi(height=32, width=64)(img, {"roll": 0.0, "pitch": 0.0, "yaw": 0.0})
    assert out.shape == (3, 64, 128)


# ---------------------------------------------------------------------------
# cube formats
# ---------------------------------------------------------------------------


@pytest.mark.parametrize("mode", MODES)
def test_equi2cube_formats_agree(mode):
    img_np = cases.noise_image((2, 3, 32, 64), "float32", 21)
    img = torch.from_numpy(img_np).to(DEV)
    rots = cases.rotation_sweep(5)[3:]
    hz = equilib_b200.equi2cube(img, rots, 16, "horizon", mode=mode)
    dice = equilib_b200.equi2cube(img, rots, 16, "dice", mode=mode)
    lst = equilib_b200.equi2cube(img, rots, 16, "list", mode=mode)
    dct = equilib_b200.equi2cube(img, rots, 16, "dict", mode=mode)
    ref_dice = oracle.equi2cube(torch.from_numpy(img_np), rots, 16, "dice", mode=mode)
    assert dice.shape == (2, 3, 48, 64)
    assert torch.equal(O.format_to_horizon(dice.cpu(), "dice"), hz.cpu())
    assert torch.equal((dice.cpu() == 0) | (ref_dice != 0), torch.ones_like(ref_dice, dtype=bool))
    for b in range(2):
        for f, k in enumerate("FRBLUD"):
            face = hz[b, :, :, f * 16:(f + 1) * 16]
            assert torch.equal(lst[b][f], face)
            assert torch.equal(dct[b][k], face)
    # (clip_output=False: the clamp range is the min/max of the whole input batch)
    single = equilib_b200.equi2cube(img[0], rots[0], 16, "dict", mode=mode, clip_output=False)
    batch = equilib_b200.equi2cube(img, rots, 16, "dict", mode=mode, clip_output=False)
    assert isinstance(single, dict) and torch.equal(single["F"], batch[0]["F"])
    with pytest.raises(NotImplementedError):
        equilib_b200.equi2cube(img, rots, 16, "cross")


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("dtype", ["float32", "uint8"])
def test_cube2equi_formats_agree(mode, dtype):
    hz_np = cases.noise_image((2, 3, 16, 96), dtype, 31)
    hz = torch.from_numpy(hz_np).to(DEV)
    ref = oracle.cube2equi(torch.from_numpy(hz_np), "horizon", 32, 64, mode=mode)
    out = equilib_b200.cube2equi(hz, "horizon", 32, 64, mode=mode)
    case = dict(dtype=dtype, mode=mode)
    compare(case, out, ref)
    dice = O.horizon_to_format(torch.from_numpy(hz_np), "dice").to(DEV)
    assert torch.equal(equilib_b200.cube2equi(dice, "dice", 32, 64, mode=mode), out)
    lst = [[hz[b, :, :, f * 16:(f + 1) * 16].contiguous() for f in range(6)] for b in range(2)]
    assert torch.equal(equilib_b200.cube2equi(lst, "list", 32, 64, mode=mode), out)
    dct = [{k: lst[b][f] for f, k in enumerate("FRBLUD")} for b in range(2)]
    assert torch.equal(equilib_b200.cube2equi(dct, "dict", 32, 64, mode=mode), out)
    # non-contiguous faces (views of the strip) take the stacking path
    views = [[hz[b, :, :, f * 16:(f + 1) * 16] for f in range(6)] for b in range(2)]
    assert torch.equal(equilib_b200.cube2equi(views, "list", 32, 64, mode=mode), out)
    one = equilib_b200.cube2equi(dct[0], "dict", 32, 64, mode=mode, clip_output=False)
    full = equilib_b200.cube2equi(dct, "dict", 32, 64, mode=mode, clip_output=False)
    assert one.shape == (3, 32, 64) and torch.equal(one, full[0])


# ---------------------------------------------------------------------------
# size-independent properties at full (BASELINE.json) sizes
# ---------------------------------------------------------------------------


def test_full_size_nearest_only_moves_pixels():
    """4K nearest: every output value is one of the input's values at the oracle's
    rounded coordinate; checked through a checksum of per-row checksums on an image
    whose pixel value encodes its own position."""
    h, w = 2048, 4096
    ys = torch.arange(h, dtype=torch.float32, device=DEV)[:, None]
    xs = torch.arange(w, dtype=torch.float32, device=DEV)[None, :]
    img = torch.stack([ys.expand(h, w), xs.expand(h, w)])[None]  # value == own (y, x)
    rots = [{"roll": 0.3, "pitch": -0.4, "yaw": 1.1}]
    out = equilib_b200.equi2equi(img, rots, mode="nearest")
    coords = gpu_coords(_lib.EQUI2EQUI, rots, False, (h, w), (h, w))
    uj = torch.from_numpy(coords[:, 0]).to(DEV)
    ui = torch.from_numpy(coords[:, 1]).to(DEV)
    # the sampler may only pick one of the <= 4 integer neighbours of (uj, ui)
    assert float((out[0, 0] - uj[0]).abs().max()) <= 1.0 + 1e-3
    dx = (out[0, 1] - ui[0]).abs()
    dx = torch.minimum(dx, (w - dx).abs())
    assert float(dx.max()) <= 1.0 + 1e-3
    assert float(out[0, 0].min()) >= 0 and float(out[0, 0].max()) <= h - 1


def test_full_size_linearity_and_constant_image():
    """bilinear is linear in the image: f(2x) == 2 f(x) bit-for-bit (power of two),
    and a constant image stays constant up to 2 ulp (weights sum to 1 +- rounding)."""
    img = torch.from_numpy(cases.band_limited_image((1, 3, 2048, 4096), "float32")).to(DEV)
    rots = [{"roll": 0.2, "pitch": 0.5, "yaw": -2.0}]
    a = equilib_b200.equi2equi(img, rots, clip_output=False)
    b = equilib_b200.equi2equi(img * 2, rots, clip_output=False)
    assert torch.equal(a * 2, b)
    const = torch.full((1, 1, 2048, 4096), 0.75, device=DEV)
    c = equilib_b200.equi2equi(const, rots, clip_output=False)
    assert float((c - 0.75).abs().max()) <= 2.5 * 2 ** -24
    d = equilib_b200.equi2equi(const, rots, clip_output=True, exact_clip=True)
    assert float(d.max()) == 0.75  # the reference's clamp removes the overshoot


def test_full_size_cube_round_trip_psnr():
    """equi -> cube (w=1024) -> equi at 4K: smooth image survives (no reference test
    pins this)."""
    img = torch.from_numpy(cases.band_limited_image((1, 3, 2048, 4096), "float32")).to(DEV)
    rot = [{"roll": 0.0, "pitch": 0.0, "yaw": 0.0}]
    for fmt in ("dice", "horizon", "dict"):
        cube = equilib_b200.equi2cube(img, rot, 1024, fmt, mode="bilinear")
        back = equilib_b200.cube2equi(cube, fmt, 2048, 4096, mode="bilinear")
        back = back if back.dim() == 4 else back[None]
        mse = float(((back[:, :, 64:-64] - img[:, :, 64:-64]) ** 2).mean())
        psnr = 10 * math.log10(1.0 / max(mse, 1e-20))
        print(f"round trip {fmt}: PSNR {psnr:.1f} dB")
        assert psnr > 35.0


def test_pers2equi_fill_follows_clip_order():
    pers = (torch.rand(1, 3, 96, 128, device=DEV) + 0.5)
    rots = [{"roll": 0.0, "pitch": 0.0, "yaw": 0.0}]
    a = equilib_b200.pers2equi(pers, rots, 128, 256, 90.0, clip_output=True)
    b = equilib_b200.pers2equi(pers, rots, 128, 256, 90.0, clip_output=False)
    assert float((b == 0).float().mean()) > 0.5
    assert torch.all(a[b == 0] == pers.min())


def test_uint8_bicubic_wraps_like_the_reference_and_can_saturate():
    img = np.zeros((1, 1, 32, 64), np.uint8)
    img[..., 32:] = 255  # step edge -> bicubic overshoot
    rot = [{"roll": 0.0, "pitch": 0.0, "yaw": 0.3}]
    x = torch.from_numpy(img).to(DEV)
    wrap = equilib_b200.equi2equi(x, rot, mode="bicubic")
    ref = oracle.equi2equi(torch.from_numpy(img), rot, mode="bicubic")
    d = np.abs(wrap.cpu().numpy().astype(int) - ref.numpy().astype(int))
    d = np.minimum(d, 256 - d)
    assert d.max() <= 1
    sat = equilib_b200.equi2equi(x, rot, mode="bicubic", saturate_uint8=True)
    ref_sat = oracle.equi2equi(torch.from_numpy(img), rot, mode="bicubic", saturate_uint8=True)
    assert np.abs(sat.cpu().numpy().astype(int) - ref_sat.numpy().astype(int)).max() <= 1
    assert int(sat.max()) == 255 and int(sat.min()) == 0


def test_errors_raise_before_launch():
    img = torch.zeros((2, 3, 16, 32), device=DEV)
    rots = cases.rotation_sweep(5)[3:]
    n0 = _lib.launch_count()
    with pytest.raises(AssertionError):
        equilib_b200.equi2equi(img, rots[:1])
    with pytest.raises(ValueError):
        equilib_b200.equi2equi(img, rots, mode="lanczos")
    with pytest.raises(ValueError):
        equilib_b200.equi2equi(img, rots, backend="pure")
    with pytest.raises(TypeError):
        equilib_b200.equi2equi(img, rots, bogus=1)
    assert _lib.launch_count() == n0


def test_launches_are_counted_and_on_current_stream():
    img = torch.rand((1, 3, 64, 128), device=DEV)
    rots = [{"roll": 0.0, "pitch": 0.1, "yaw": 0.2}]
    s = torch.cuda.Stream()
    n0 = _lib.launch_count()
    with torch.cuda.stream(s):
        a = equilib_b200.equi2equi(img, rots)
    s.synchronize()
    assert _lib.launch_count() == n0 + 1
    b = equilib_b200.equi2equi(img, rots)
    assert torch.equal(a, b)


def test_float64_images_run_in_float32_and_come_back_as_float64():
    """The reference accepts float64 (equilib/equi2equi/torch.py:87-95); here such images are
    computed by the float32 kernels and returned as float64 (K.accepts_float64)."""
    img = torch.from_numpy(cases.band_limited_image((2, 3, 64, 128), "float32", px=40, py=30))
    rots = cases.rotation_sweep(5)[3:]
    out = equilib_b200.equi2equi(img.double().to(DEV), rots)
    assert out.dtype == torch.float64
    assert torch.equal(out, equilib_b200.equi2equi(img.to(DEV), rots).double())
    ref = oracle.equi2equi(img, rots, flavour="cuda")
    assert float((out.cpu().float() - ref).abs().max()) <= 2e-5
    cube = equilib_b200.equi2cube(img.double().to(DEV), rots, 16, "dict")
    assert cube[0]["F"].dtype == torch.float64
    back = equilib_b200.cube2equi(cube, "dict", 64, 128)
    assert back.dtype == torch.float64


def test_cuda_graph_capture_and_replay():
    """A transform call is capturable (no allocation outside torch's pool, no sync, launch on
    the capturing stream, tensor maps and matrices by value / from caches): replaying the
    graph on new frame contents gives what a fresh call gives."""
    rots = cases.rotation_sweep(5)[3:]
    op = equilib_b200.Equi2Pers(480, 640, 90.0)   # >= 2^18 pixels: the TMA-staged kernel
    frames = [torch.from_numpy(cases.noise_image((2, 3, 512, 1024), "float32", s)).to(DEV)
              for s in (31, 32)]
    static_in = frames[0].clone()
    op(static_in, rots)  # warm-up: device tables and matrices are uploaded (and cached) here
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        op(static_in, rots)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    n0 = _lib.launch_count()
    with torch.cuda.graph(g):
        static_out = op(static_in, rots)
    assert _lib.launch_count() == n0 + 1
    for f in frames:
        static_in.copy_(f)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(static_out, op(f, rots))
    assert _lib.launch_count() == n0 + 1 + len(frames)  # replays launch nothing from the host


def test_host_pipeline_matches_direct_call():
    """The returned host tensor is complete when the call returns (no synchronize here: the
    pipeline hands completion to the HOST by default), for batches shorter and longer than
    its ring of device buffers, and with sync=False through the `done` event."""
    from equilib_b200.pipeline import HostPipeline

    frames = torch.from_numpy(cases.noise_image((23, 3, 64, 128), "uint8", 77)).pin_memory()
    rots = cases.rotation_sweep(26)[3:]
    op = equilib_b200.Equi2Equi(mode="bilinear")
    want = op(frames.to(DEV), rots).cpu()
    for chunk, depth in ((4, 3), (2, 2), (32, 3), (1, 4)):
        pipe = HostPipeline(op, chunk=chunk, depth=depth)
        got = pipe(frames, rots)
        assert got.is_pinned() and torch.equal(got, want), (chunk, depth)
        out = torch.zeros_like(frames).pin_memory()
        assert pipe(frames, rots, out=out) is out
        assert torch.equal(out, want)
        out.zero_()
        pipe(frames, rots, out=out, sync=False)
        pipe.done.synchronize()
        assert torch.equal(out, want)


def test_host_views_pipeline_matches_direct_call():
    from equilib_b200.pipeline import HostViewsPipeline

    equi = torch.from_numpy(cases.band_limited_image((1, 3, 256, 512), "float32", px=96,
                                                     py=64))[0].to(torch.float16).pin_memory()
    rots = [{"roll": 0.0, "pitch": 0.3 * math.sin(i), "yaw": 0.4 * i} for i in range(11)]
    op = equilib_b200.Equi2Pers(96, 128, 90.0)
    want = op(equi.to(DEV)[None].expand(11, -1, -1, -1), rots).cpu()
    got = HostViewsPipeline(op, chunk=4)(equi, rots)
    assert got.is_pinned() and torch.equal(got, want)


def test_get_bounding_fov_matches_oracle_perimeter():
    """equilib/equi2pers/torch.py:247-351: rounded (row, col) of the view's perimeter."""
    h, w, H, W = 48, 64, 256, 512
    rots = cases.rotation_sweep(6)[3:]
    equi = torch.zeros((3, 3, H, W), device=DEV)
    got = equilib_b200.Equi2Pers(h, w, 90.0).get_bounding_fov(equi, rots)
    assert got.shape == (3, 2 * w + 2 * (h - 2), 2) and got.dtype == np.int64
    uj, ui = O.equi2pers_coords(rots, False, h, w, 90.0, 0.0, H, W)
    g = torch.stack([uj, ui], dim=1)  # (B, 2, h, w)
    per = []
    per += [g[:, :, 0, x] for x in range(w)]
    per += [g[:, :, y, w - 1] for y in range(1, h)]
    per += [g[:, :, h - 1, x] for x in range(w - 2, 0, -1)]
    per += [g[:, :, y, 0] for y in range(h - 1, 0, -1)]
    want = np.rint(torch.stack(per, dim=1).numpy()).astype(np.int64)
    d = np.abs(got - want)
    d[..., 1] = np.minimum(d[..., 1], W - d[..., 1])  # seam
    assert np.mean(d <= 1) >= 0.9999 and rate(np.mean(d == 0), 0.995)
    one = equilib_b200.get_bounding_fov(equi[0], rots[0], h, w, 90.0)
    assert one.shape == (2 * w + 2 * (h - 2), 2) and np.array_equal(one, got[0])


# ---------------------------------------------------------------------------
# staged-kernel edge cases at sizes that actually take it (>= 2^18 output pixels)
# ---------------------------------------------------------------------------


def _oracle_vs_product(fn_name, img, rots, *args, dtype_tol=None, **kw):
    out = getattr(equilib_b200, fn_name)(img.to(DEV), rots, *args, **kw)
    ref = getattr(oracle, fn_name)(img, rots, *args, flavour="cuda", **kw)
    o, r = out.float().cpu().numpy(), ref.float().numpy()
    return o, r


@pytest.mark.parametrize("channels", [1, 4])
def test_staged_channel_counts(channels):
    img = torch.from_numpy(cases.band_limited_image((2, channels, 512, 1024), "float32", px=96,
                                                    py=64)).to(torch.bfloat16)
    o, r = _oracle_vs_product("equi2equi", img, HARD_ROTS[1:3])
    assert close_except_seam(o, r, 2.0 ** -7 * 1.01 + 2e-4, 0.99999) and rate(np.mean(o == r), 0.997)


def test_staged_expanded_batch_single_equirect_many_views():
    """BASELINE config 5 shape: one equirect, stride-0 batch, many rotations (the TMA map
    is encoded without a batch dimension)."""
    one = torch.from_numpy(cases.band_limited_image((1, 3, 512, 1024), "float32", px=96, py=64))
    one = one.to(torch.float16)
    rots = cases.rotation_sweep(9)[3:]
    dev = one.to(DEV)
    a = equilib_b200.equi2pers(dev.expand(6, -1, -1, -1), rots, 256, 256, 90.0)
    b = equilib_b200.equi2pers(dev.repeat(6, 1, 1, 1), rots, 256, 256, 90.0)
    assert torch.equal(a, b)
    ref = oracle.equi2pers(one.repeat(6, 1, 1, 1), rots, 256, 256, 90.0, flavour="cuda")
    d = (a.float().cpu() - ref.float()).abs()
    # The reference does not interpolate across the seam: ui just below W reads column W-1,
    # ui just above 0 blends columns 0 and 1.  A pixel whose longitude sits within the
    # coordinate noise of the seam can therefore land on either side (one pixel here, on a
    # view looking at the pole); everything else is within one fp16 ulp.
    tol = 2.0 ** -10 * 1.01 + 2e-4
    assert float((d <= tol).float().mean()) >= 0.99999
    assert rate(float((d == 0).float().mean()), 0.99)


def test_staged_sizes_not_multiples_of_the_tile():
    img = torch.from_numpy(cases.noise_image((2, 3, 500, 1000), "uint8", 9))
    rots = HARD_ROTS[1:3]
    # 1000-wide rows are 8-byte aligned but not 16: cp.async/TMA need 16 -> direct kernel;
    # 1008-wide rows are staged with a ragged last tile
    for w in (1000, 1008):
        x = img if w == 1000 else torch.from_numpy(cases.noise_image((2, 3, 500, 1008), "uint8", 9))
        out = equilib_b200.equi2equi(x.to(DEV), rots, height=503, width=1012)
        ref = oracle.equi2equi(x, rots, height=503, width=1012, flavour="cuda")
        d = np.abs(out.cpu().numpy().astype(int) - ref.numpy().astype(int))
        assert np.mean(d <= 1) >= 0.9999 and rate(np.mean(d == 0), 0.995)
        exact = equilib_b200.equi2equi(x.to(DEV), rots, height=503, width=1012, fast_coords=False)
        d2 = np.abs(exact.cpu().numpy().astype(int) - ref.numpy().astype(int))
        assert np.mean(d2 <= 1) >= 0.9999 and rate(np.mean(d2 == 0), 0.997)


def test_staged_8k_source_bicubic_uint8():
    """BASELINE config 4 shape (one frame): 8K uint8 source, bicubic, 1080p view."""
    img = torch.from_numpy(cases.band_limited_image((1, 3, 4096, 8192), "uint8", px=300, py=200))
    rot = [HARD_ROTS[1]]
    out = equilib_b200.equi2pers(img.to(DEV), rot, 1080, 1920, 90.0, mode="bicubic")
    ref = oracle.equi2pers(img, rot, 1080, 1920, 90.0, mode="bicubic", flavour="cuda")
    d = np.abs(out.cpu().numpy().astype(int) - ref.numpy().astype(int))
    d = np.minimum(d, 256 - d)
    assert np.mean(d <= 1) >= 0.9999 and rate(np.mean(d == 0), 0.99)


def test_equi2cube_fast_coordinates_faces_of_64():
    """Fast coordinates need face widths that are multiples of 64; others use the exact
    chain.  Both must agree with the oracle within one code value."""
    img = torch.from_numpy(cases.band_limited_image((2, 3, 512, 1024), "uint8", px=96, py=64))
    rots = HARD_ROTS[1:3]
    for w_face in (256, 200):
        out = equilib_b200.equi2cube(img.to(DEV), rots, w_face, "dice")
        ref = oracle.equi2cube(img, rots, w_face, "dice", flavour="cuda")
        d = np.abs(out.cpu().numpy().astype(int) - ref.numpy().astype(int))
        assert np.mean(d <= 1) >= 0.9999 and rate(np.mean(d == 0), 0.997)


# --- channels-last uint8 -------------------------------------------------


@pytest.mark.parametrize("channels", [3, 4])
@pytest.mark.parametrize("transform", ["equi2equi", "equi2pers"])
def test_channels_last_uint8_in_place(transform, channels):
    """A uint8 RGB / RGBX batch in torch.channels_last (what decoders hand over) is read
    and written in that layout: same values as the planar call, within one code value of
    the oracle, output in channels_last too."""
    shape = (2, channels, 512, 1024)
    img = torch.from_numpy(0.5 * cases.band_limited_image(shape, "float32", px=96, py=64) * 255
                           + 0.5 * cases.noise_image(shape, "uint8", 5)).to(torch.uint8)
    rots = HARD_ROTS[1:3]
    args = () if transform == "equi2equi" else (512, 768, 100.0)
    fn = getattr(equilib_b200, transform)
    planar = fn(img.to(DEV), rots, *args)
    x_cl = img.to(DEV).contiguous(memory_format=torch.channels_last)
    n0 = _lib.lib().eqb_launch_count()
    out = fn(x_cl, rots, *args)
    assert _lib.lib().eqb_launch_count() == n0 + 1
    assert out.shape == planar.shape and out.is_contiguous(memory_format=torch.channels_last)
    assert not out.is_contiguous()
    d = (out.int() - planar.int()).abs()
    assert int(d.max()) <= 1 and rate(float((d == 0).float().mean()), 0.999)
    ref = getattr(oracle, transform)(img, rots, *args, flavour="cuda")
    d = np.abs(out.cpu().numpy().astype(int) - ref.numpy().astype(int))
    assert np.mean(d <= 1) >= 0.9999 and rate(np.mean(d == 0), 0.996)


def test_channels_last_ragged_and_pole_tiles():
    """Widths that are not multiples of the 64-pixel tile, a view looking at a pole (tiles
    that fit no box gather from global memory) and the right image edge."""
    img = torch.from_numpy(cases.noise_image((2, 3, 512, 1040), "uint8", 11))
    rots = [HARD_ROTS[0], {"roll": 0.3, "pitch": math.pi / 2 - 0.05, "yaw": 1.0}]
    x_cl = img.to(DEV).contiguous(memory_format=torch.channels_last)
    out = equilib_b200.equi2equi(x_cl, rots, height=516, width=1036)
    assert out.is_contiguous(memory_format=torch.channels_last)
    ref = oracle.equi2equi(img, rots, height=516, width=1036, flavour="cuda")
    d = np.abs(out.cpu().numpy().astype(int) - ref.numpy().astype(int))
    assert np.mean(d <= 1) >= 0.9995 and rate(np.mean(d == 0), 0.995)


def test_channels_last_falls_back_to_planar_where_not_served():
    img = torch.from_numpy(cases.noise_image((1, 3, 128, 256), "uint8", 3))
    x_cl = img.to(DEV).contiguous(memory_format=torch.channels_last)
    rots = [HARD_ROTS[1]]
    # small launch, bicubic, another transform: converted like every input of the reference
    for out, want in (
        (equilib_b200.equi2equi(x_cl, rots), equilib_b200.equi2equi(img.to(DEV), rots)),
        (equilib_b200.equi2equi(x_cl, rots, mode="bicubic"),
         equilib_b200.equi2equi(img.to(DEV), rots, mode="bicubic")),
        (equilib_b200.equi2cube(x_cl, rots, 64, "horizon"),
         equilib_b200.equi2cube(img.to(DEV), rots, 64, "horizon")),
    ):
        assert out.is_contiguous() and torch.equal(out, want)
    # the C ABI refuses the flag where the kernel does not exist
    big = torch.zeros((1, 3, 512, 1024), dtype=torch.float16, device=DEV)
    big = big.contiguous(memory_format=torch.channels_last)
    with pytest.raises(_lib.EquilibB200Error, match="channels-last"):
        from equilib_b200 import _common as K
        tab_x, tab_y = _prep.equirect_tables(512, 1024, DEV)
        K.launch(big, torch.empty_like(big), transform=_lib.EQUI2EQUI, mode="bilinear",
                 flags=_lib.FLAG_CHANNELS_LAST, batch=1, channels=3, src_hw=(512, 1024),
                 dst_hw=(512, 1024), src_strides=big.stride()[:3],
                 dst_strides=big.stride()[:3],
                 mats=_prep.mats_for("rot", rots, False, DEV), tab_x=tab_x, tab_y=tab_y)


def test_lean_kernel_random_rotations_match_the_exact_path():
    """The headline kernel (fast coordinates, speculative TMA boxes, per-lane fallback)
    against the exact-coordinate path on random rotations, including views that put a
    source pole or the seam anywhere in the frame."""
    rng = np.random.default_rng(2024)
    img = torch.from_numpy(cases.band_limited_image((8, 3, 512, 1024), "float32", px=96, py=64))
    img = img.to(torch.bfloat16).to(DEV)
    for trial in range(3):
        rots = [{"roll": float(rng.uniform(-math.pi, math.pi)),
                 "pitch": float(rng.choice([rng.uniform(-1.6, 1.6), math.pi / 2, -math.pi / 2,
                                            1.5, -1.45])),
                 "yaw": float(rng.uniform(-math.pi, math.pi))} for _ in range(8)]
        fast = equilib_b200.equi2equi(img, rots)
        exact = equilib_b200.equi2equi(img, rots, fast_coords=False)
        d = (fast.float() - exact.float()).abs()
        # one bf16 ulp of values in [0, 1] is 2^-8 at most; seam pixels may flip (see
        # close_except_seam)
        assert float((d <= 2.0 ** -8 * 1.01).float().mean()) >= 0.9999
        assert rate(float((d == 0).float().mean()), 0.9975)
        u8 = (img.float() * 255).to(torch.uint8)
        a = equilib_b200.equi2equi(u8, rots).int()
        b = equilib_b200.equi2equi(u8, rots, fast_coords=False).int()
        assert float(((a - b).abs() <= 1).float().mean()) >= 0.9999


# --- static-camera map ----------------------------------------------------


@pytest.mark.parametrize("dtype,channels", [(torch.bfloat16, 3), (torch.uint8, 4),
                                            (torch.float16, 1)])
def test_static_remap_equals_the_exact_path(dtype, channels):
    """A map built once for fixed rotations gives what the exact-coordinate call gives,
    up to its 16-bit bilinear fractions (1.5e-5 of a pixel)."""
    shape = (3, channels, 512, 1024)
    img = torch.from_numpy(cases.band_limited_image(shape, "float32", px=96, py=64))
    img = (img * 255).to(torch.uint8) if dtype == torch.uint8 else img.to(dtype)
    img = img.to(DEV)
    rots = HARD_ROTS[1:4]  # includes a pole-up view (tiles that fit no box)
    sr = equilib_b200.StaticRemap.equi2equi(rots, img.shape, dtype, DEV, height=520, width=1012)
    n0 = _lib.lib().eqb_launch_count()
    out = sr(img)
    assert _lib.lib().eqb_launch_count() == n0 + 1  # one kernel per frame batch
    ref = equilib_b200.equi2equi(img, rots, height=520, width=1012, fast_coords=False)
    d = (out.float() - ref.float()).abs()
    tol = 1.0 if dtype == torch.uint8 else (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10)
    assert float((d <= tol * 1.01).float().mean()) >= 0.9999
    assert rate(float((d == 0).float().mean()), 0.995)
    # a second batch through the same map
    img2 = img.flip(3).contiguous()
    assert torch.equal(sr(img2), sr(img2.clone()))
    # perspective views of the same frames
    sp = equilib_b200.StaticRemap.equi2pers(rots, img.shape, dtype, DEV, 512, 768, 100.0)
    ref_p = equilib_b200.equi2pers(img, rots, 512, 768, 100.0, fast_coords=False)
    dp = (sp(img).float() - ref_p.float()).abs()
    assert float((dp <= tol * 1.01).float().mean()) >= 0.9999
    with pytest.raises(ValueError):
        sr(img[:2])


@pytest.mark.parametrize("channels", [3, 4])
def test_static_remap_channels_last_uint8(channels):
    """Decoder-layout frames through a static map: in place, same values as planar."""
    shape = (2, channels, 512, 1024)
    img = torch.from_numpy(cases.noise_image(shape, "uint8", 21)).to(DEV)
    rots = HARD_ROTS[1:3]
    sr = equilib_b200.StaticRemap.equi2equi(rots, shape, torch.uint8, DEV)
    planar = sr(img)
    out = sr(img.contiguous(memory_format=torch.channels_last))
    assert out.is_contiguous(memory_format=torch.channels_last) and not out.is_contiguous()
    assert torch.equal(out, planar)
    ref = equilib_b200.equi2equi(img, rots, fast_coords=False)
    d = (out.int() - ref.int()).abs()
    assert int(d.max()) <= 1 and rate(float((d == 0).float().mean()), 0.99)


# --- strip-node kernel, bicubic (round 2) ---------------------------------------------------


@pytest.mark.parametrize("dtype", ["float32", "float16", "bfloat16", "uint8"])
@pytest.mark.parametrize("channels", [1, 3, 4])
def test_lean2_bicubic_equals_the_direct_kernel(dtype, channels, monkeypatch):
    """Large bicubic launches of equi2equi / equi2pers / equi2cube take the strip-node kernel
    (4 x 4 taps from the TMA box, tiles at the source border / poles / seam on its general
    path).  With the exact chain (``fast_coords=False``; the fp32 default) it must produce the
    bits of the direct-gather kernel: every channel count, ragged sizes, with and without the
    clamp."""
    img = torch.from_numpy(cases.noise_image((2, channels, 512, 1024),
                                             "uint8" if dtype == "uint8" else "float32", 7))
    img = img.to(DEV)
    if dtype not in ("uint8", "float32"):
        img = img.to(getattr(torch, dtype))
    rots = [{"roll": 0.3, "pitch": -0.4, "yaw": 1.1}, {"roll": 1.5, "pitch": 1.2, "yaw": -2.9}]

    def run_all():
        kw = dict(mode="bicubic", fast_coords=False)
        return [
            equilib_b200.equi2equi(img, rots, **kw),
            equilib_b200.equi2equi(img, rots, clip_output=False, **kw),
            equilib_b200.equi2equi(img, rots, height=500, width=1002, **kw),
            equilib_b200.equi2pers(img, rots, 480, 640, 100.0, **kw),
            equilib_b200.equi2cube(img, rots, 256, "dice", **kw),
            equilib_b200.equi2cube(img, rots, 192, "horizon", clip_output=False, **kw),
        ]

    n0 = _lib.launch_count()
    lean2 = run_all()
    assert _lib.launch_count() > n0
    monkeypatch.setenv("EQUILIB_B200_NO_STAGE", "1")
    direct = run_all()
    for i, (a, b) in enumerate(zip(lean2, direct)):
        assert torch.equal(a, b), f"output {i}: {int((a != b).sum())} elements differ"


@pytest.mark.parametrize("dtype", ["bfloat16", "uint8", "float32"])
def test_lean2_bicubic_interpolated_coordinates_vs_oracle(dtype):
    """Interpolated coordinates of the strip's nodes under bicubic sampling (the default for
    8/16-bit images, opt-in for fp32).  Against the oracle on the band-limited
    image: within one code value / one ulp of the output type (fp32: 2e-3 px of coordinate
    noise times the image gradient)."""
    img = torch.from_numpy(cases.band_limited_image((2, 3, 512, 1024),
                                                    "uint8" if dtype == "uint8" else "float32",
                                                    px=96, py=64))
    x = img if dtype in ("uint8", "float32") else img.to(getattr(torch, dtype))
    rots = HARD_ROTS[1:3]
    out = equilib_b200.equi2equi(x.to(DEV), rots, mode="bicubic", fast_coords=True)
    ref = oracle.equi2equi(x, rots, mode="bicubic", flavour="cuda")
    if dtype == "uint8":
        d = np.abs(out.cpu().numpy().astype(int) - ref.numpy().astype(int))
        d = np.minimum(d, 256 - d)
        assert np.mean(d <= 1) >= 0.9995 and rate(np.mean(d == 0), 0.998)  # measured 0.99978
    elif dtype == "bfloat16":
        o, r = out.float().cpu().numpy(), ref.float().numpy()
        assert close_except_seam(o, r, 2.0 ** -7, 0.9995)
        assert rate(np.mean(o == r), 0.998)  # measured 0.99951
    else:
        assert close_except_seam(out.cpu().numpy(), ref.numpy(), 2e-3, 0.9995)


# --- fused equi -> cube -> equi ----------------------------------------------


@pytest.mark.parametrize("dtype", ["float32", "bfloat16", "uint8"])
@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
def test_fused_cube_chain_equals_two_launches(dtype, mode):
    """equi2cube2equi == cube2equi(equi2cube(x)) of this package bit for bit (exact chain on
    both sides), whatever the cube format of the two-call version, including taps that bleed
    across faces, faces that are not multiples of 64 and single-image shape rules."""
    img = torch.from_numpy(cases.noise_image((3, 3, 256, 512),
                                             "uint8" if dtype == "uint8" else "float32", 11))
    img = img.to(DEV)
    if dtype == "bfloat16":
        img = img.to(torch.bfloat16)
    rots = [{"roll": 0.0, "pitch": 0.0, "yaw": 0.0}, {"roll": 0.3, "pitch": -0.4, "yaw": 1.1},
            {"roll": 1.5, "pitch": 1.2, "yaw": -2.9}]
    for w_face, (h, w), fmt in ((128, (256, 512), "horizon"), (100, (264, 520), "dice"),
                                (64, (128, 256), "dict")):
        cube = equilib_b200.equi2cube(img, rots, w_face, fmt, mode=mode, fast_coords=False)
        two = equilib_b200.cube2equi(cube, fmt, h, w, mode=mode)
        n0 = _lib.launch_count()
        one = equilib_b200.equi2cube2equi(img, rots, w_face, h, w, mode=mode, fused=True)
        assert _lib.launch_count() == n0 + 1
        assert one.shape == two.shape and one.dtype == two.dtype
        assert torch.equal(one, two), f"{fmt}: {int((one != two).sum())} elements differ"
    # single image + dict of rotations -> (C, H, W), like the two calls
    one = equilib_b200.equi2cube2equi(img[1], rots[1], 64, 128, 256, mode=mode, fused=True)
    cube = equilib_b200.equi2cube(img[1], rots[1], 64, "horizon", mode=mode, fast_coords=False)
    assert torch.equal(one, equilib_b200.cube2equi(cube, "horizon", 128, 256, mode=mode))
    # class API, z_down
    obj = equilib_b200.Equi2Cube2Equi(64, 128, 256, z_down=True, mode=mode)
    cube = equilib_b200.equi2cube(img, rots, 64, "horizon", z_down=True, mode=mode,
                                  fast_coords=False)
    two = equilib_b200.cube2equi(cube, "horizon", 128, 256, mode=mode)
    assert torch.equal(obj(img, rots, fused=True), two)
    # fused=None: nearest fused (16/32-bit) or two launches, bilinear two launches with the
    # default coordinates -- nearest is exact either way
    auto = obj(img, rots)
    if mode == "nearest" or dtype == "float32":
        assert torch.equal(auto, two)
    else:
        d = (auto.float() - two.float()).abs()
        assert float((d <= (1.0 if dtype == "uint8" else 2.0 ** -7)).float().mean()) >= 0.9995


def test_fused_cube_chain_vs_oracle_composition():
    """Against the oracle's own two-step composition (reference semantics end to end)."""
    img = cases.band_limited_image((2, 3, 256, 512), "float32", px=96, py=64)
    rots = HARD_ROTS[1:3]
    cube = oracle.equi2cube(torch.from_numpy(img), rots, 128, "horizon", flavour="cuda")
    ref = oracle.cube2equi(cube, "horizon", 256, 512, flavour="cuda")
    for fused in (True, None):
        out = equilib_b200.equi2cube2equi(torch.from_numpy(img).to(DEV), rots, 128, 256, 512,
                                          fused=fused)
        assert helpers.frac_close(out.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=2e-5) >= 0.999
    # bicubic and exact_clip take the two launches behind the same entry point
    out = equilib_b200.equi2cube2equi(torch.from_numpy(img).to(DEV), rots, 128, 256, 512,
                                      mode="bicubic")
    cube = oracle.equi2cube(torch.from_numpy(img), rots, 128, "horizon", mode="bicubic",
                            flavour="cuda")
    ref = oracle.cube2equi(cube, "horizon", 256, 512, mode="bicubic", flavour="cuda")
    assert helpers.frac_close(out.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=5e-5) >= 0.995


@pytest.mark.parametrize("dtype", ["float32", "uint8"])
def test_lean2_pers2equi_and_small_sources_equal_the_direct_kernel(dtype, monkeypatch):
    """pers2equi bicubic in the strip-node kernel -- fill strips decided from tile corners, fill
    tiles from node masks, in-view tiles from TMA boxes -- against the direct-gather kernel, bit
    for bit: wide and narrow fields of view, skew, z_down, a view straddling the seam, a tiny
    perspective image (boxes larger than the source).  Likewise bicubic equi2equi from a source
    much smaller than one box (every tile touches the border where ATen reflects taps)."""
    shape = (2, 3, 270, 480)
    img = torch.from_numpy(cases.noise_image(shape, "uint8" if dtype == "uint8" else "float32", 3))
    img = img.to(DEV)
    rots = [{"roll": 0.2, "pitch": -0.3, "yaw": 3.1}, {"roll": -1.2, "pitch": 1.3, "yaw": 0.4}]

    def run_all():
        return [
            equilib_b200.pers2equi(img, rots, 512, 1024, 150.0, mode="bicubic"),
            equilib_b200.pers2equi(img, rots, 512, 1024, 35.0, mode="bicubic", clip_output=False),
            equilib_b200.pers2equi(img, rots, 512, 1024, 100.0, skew=0.3, z_down=True,
                                   mode="bicubic"),
            equilib_b200.pers2equi(img[:, :, :48, :64].contiguous(), rots, 520, 1000, 90.0,
                                   mode="bicubic"),
            equilib_b200.equi2equi(img[:, :, :32, :64].contiguous(), rots, height=512, width=1024,
                                   mode="bicubic", fast_coords=False),
        ]

    lean2 = run_all()
    monkeypatch.setenv("EQUILIB_B200_NO_STAGE", "1")
    direct = run_all()
    for i, (a, b) in enumerate(zip(lean2, direct)):
        assert torch.equal(a, b), f"output {i}: {int((a != b).sum())} elements differ"


@pytest.mark.parametrize("dtype", ["float32", "float16", "bfloat16", "uint8"])
@pytest.mark.parametrize("channels", [1, 3, 4])
def test_lean2_nearest_equals_the_direct_kernel(dtype, channels, monkeypatch):
    """Large nearest launches of equi2equi / equi2pers / equi2cube take the strip-node kernel
    (one tap per channel from the TMA box, the exact chain on every pixel: the tap is a rounding
    decision).  Output elements are source elements, so the result must equal the direct-gather
    kernel's bit for bit -- noise image, poles, seam, ragged sizes, NaN-free."""
    img = torch.from_numpy(cases.noise_image((2, channels, 512, 1024),
                                             "uint8" if dtype == "uint8" else "float32", 13))
    img = img.to(DEV)
    if dtype not in ("uint8", "float32"):
        img = img.to(getattr(torch, dtype))
    rots = [{"roll": 0.3, "pitch": -0.4, "yaw": 1.1}, {"roll": 1.5, "pitch": 1.2, "yaw": -2.9}]

    def run_all():
        kw = dict(mode="nearest")
        return [
            equilib_b200.equi2equi(img, rots, **kw),
            equilib_b200.equi2equi(img, rots, height=500, width=1002, **kw),
            equilib_b200.equi2pers(img, rots, 480, 640, 100.0, **kw),
            equilib_b200.equi2cube(img, rots, 256, "dice", **kw),
            equilib_b200.equi2cube(img, rots, 192, "horizon", **kw),
        ]

    lean2 = run_all()
    monkeypatch.setenv("EQUILIB_B200_NO_STAGE", "1")
    direct = run_all()
    for i, (a, b) in enumerate(zip(lean2, direct)):
        assert torch.equal(a, b), f"output {i}: {int((a != b).sum())} elements differ"


def test_lean2_nearest_on_rounding_ties():
    """Identity and pure-yaw views put every row coordinate on a rounding tie (y + 0.5): the
    tie guard must hand ALL those pixels to the exact chain -- per tile inside a mixed batch
    (header flag), up front when the whole batch is pure yaw (host) -- and the result must
    still equal the direct kernel's."""
    img = torch.from_numpy(cases.noise_image((3, 3, 512, 1024), "uint8", 17)).to(DEV)
    mixed = [{"roll": 0.0, "pitch": 0.0, "yaw": 0.0}, {"roll": 0.0, "pitch": 0.0, "yaw": 0.7},
             {"roll": 0.2, "pitch": -0.3, "yaw": 1.0}]
    yaw = [{"roll": 0.0, "pitch": 0.0, "yaw": 0.3 * i} for i in range(3)]
    for x in (img, img.to(torch.bfloat16)):
        for rots in (mixed, yaw):
            out = equilib_b200.equi2equi(x, rots, mode="nearest")
            exact = equilib_b200.equi2equi(x, rots, mode="nearest", fast_coords=False)
            assert torch.equal(out, exact), f"{int((out != exact).sum())} elements differ"
