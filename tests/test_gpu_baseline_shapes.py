"""GPU parity at the sizes BASELINE.json's configs name (VERDICT r01 "parity holes").

The small-size tests of test_gpu_parity.py pin every code path against the oracle; the
fast-coordinate error, the TMA box selection and the pole / seam fall-backs are
size-dependent, so the same comparison is repeated here at 2048 x 4096 (configs 2 / 5),
w_face = 1024 (config 3) and 1080p <-> 8K (config 4), on a band-limited image (the parity
image of SURVEY 8d) AND on uniform noise (worst case: |grad| ~ 1 per pixel, every last-ulp
difference of asinf / atan2f shows).

Every test appends its measured match rates to gpurun_out/parity_rates.jsonl (copied to
profiles/ by hand after a run; DESIGN.md section 2 tabulates them).  The bounds asserted
below sit ~0.1 % under the rates measured on a B200, so drift shows.

Tolerances: fp32 -- 1e-5 relative + 1e-5 absolute on the band-limited image (north_star),
5e-4 absolute on noise; bf16 / fp16 -- 1 ulp of the output type against the fp32 oracle
rounded once; uint8 -- identical code values up to truncation flips (|d| <= 1);
nearest -- identical away from .5 ties.
"""

import json
import math
import os

import numpy as np
import pytest
import torch

import equilib_b200
import oracle
from tests import cases

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda", 0)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H4K, W4K = 2048, 4096

# bench.py's rotations i = 0, 7, 19, 31 (the headline pattern) and a pole-up view
HEADLINE_ROTS = [cases.bench_rotations(32)[i] for i in (0, 7, 19, 31)] + \
    [{"roll": 0.2, "pitch": math.pi / 2 - 0.01, "yaw": 1.0}]


def record(name, **rates):
    rec = {"test": name}
    rec.update({k: (round(float(v), 6) if isinstance(v, (float, np.floating)) else v)
                for k, v in rates.items()})
    print(json.dumps(rec))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_rates.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass


def image_4k(kind, dtype, n, channels=3, seed=0):
    shape = (n, channels, H4K, W4K)
    if dtype == torch.uint8:
        a = cases.noise_image(shape, "uint8", seed) if kind == "noise" else \
            cases.band_limited_image(shape, "uint8", seed)
        return torch.from_numpy(a)
    a = cases.noise_image(shape, "float32", seed) if kind == "noise" else \
        cases.band_limited_image(shape, "float32", seed)
    return torch.from_numpy(a).to(dtype)


def u8_diff(a, b):
    d = (a.to(torch.int16) - b.to(torch.int16)).abs()
    return torch.minimum(d, 256 - d)


# ---------------------------------------------------------------------------------------
# (a) headline shape: 3 x 2048 x 4096 bf16 / uint8 bilinear equi2equi
# ---------------------------------------------------------------------------------------

# Share of outputs identical to the oracle, per frame (= per rotation), measured on a B200
# (profiles/r02_parity_rates.jsonl; DESIGN.md 2 tabulates them).  Frame 0 is the IDENTITY
# rotation: every sampling coordinate is x + 0.5 exactly, every output the mean of two
# neighbours -- a rounding tie of the 8/16-bit result wherever the neighbours' sum is odd in
# the last place, so interpolated coordinates (1e-4 px off) flip ~10 % of them by one ulp
# while the exact chain reproduces the ties.  Bounds: (default path, exact path).
HEADLINE_BOUNDS = {
    # (dtype, image): [(frame 0 = identity), (other frames, worst)]
    ("bf16", "band"): [(0.890, 0.992), (0.9985, 0.9989)],   # measured .897/.9935, .99966/.99993
    ("bf16", "noise"): [(0.955, 0.996), (0.974, 0.9943)],   # measured .959/.9974, .9753/.9954
    ("u8", "band"): [(0.905, 0.9915), (0.9964, 0.9984)],    # measured .910/.9927, .9974/.9995
    ("u8", "noise"): [(0.898, 0.993), (0.986, 0.9964)],     # measured .903/.9945, .9871/.9974
}


@pytest.mark.parametrize("kind", ["band", "noise"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.uint8], ids=["bf16", "u8"])
def test_headline_4k_vs_oracle(dtype, kind):
    """Config 2's own shape against the oracle (fp32 reference chain on the up-cast input,
    one rounding at the end): the default path (interpolated coordinates) and the
    exact-coordinate path, each frame a different headline rotation."""
    name = "bf16" if dtype == torch.bfloat16 else "u8"
    n = len(HEADLINE_ROTS)
    img = image_4k(kind, dtype, n)
    ref = oracle.equi2equi(img, HEADLINE_ROTS, mode="bilinear", flavour="cuda")
    x = img.to(DEV)
    out = equilib_b200.equi2equi(x, HEADLINE_ROTS).cpu()
    out_exact = equilib_b200.equi2equi(x, HEADLINE_ROTS, fast_coords=False).cpu()
    assert out.dtype == dtype and out.shape == ref.shape
    if dtype == torch.uint8:
        d, de = u8_diff(out, ref).float(), u8_diff(out_exact, ref).float()
        tol = 1.0
    else:
        d = (out.float() - ref.float()).abs()
        de = (out_exact.float() - ref.float()).abs()
        tol = 2.0 ** -7 * 1.01  # one bf16 ulp of values in [0.5, 1)
    same = [float((d[i] == 0).float().mean()) for i in range(n)]
    same_e = [float((de[i] == 0).float().mean()) for i in range(n)]
    within = float((d <= tol).float().mean())
    within_e = float((de <= tol).float().mean())
    record("headline_4k", dtype=name, image=kind,
           identical_default_per_frame=[round(v, 5) for v in same],
           identical_exact_per_frame=[round(v, 5) for v in same_e],
           within_tol_default=within, within_tol_exact=within_e,
           max_default=float(d.max()), max_exact=float(de.max()))
    # seam pixels (no interpolation across the seam in the reference) may differ by more
    assert within >= 0.9999 and within_e >= 0.9999
    (id_f, id_e), (lo_f, lo_e) = HEADLINE_BOUNDS[(name, kind)]
    assert same[0] >= id_f and same_e[0] >= id_e, (same[0], same_e[0])
    assert min(same[1:]) >= lo_f and min(same_e[1:]) >= lo_e, (same, same_e)


def test_headline_4k_fp32_vs_oracle():
    """fp32 at 4K (the dtype the reference itself accepts): north_star tolerance on the
    band-limited image, the noise image as the stated worst case."""
    rots = HEADLINE_ROTS[1:3] + HEADLINE_ROTS[4:]
    for kind in ("band", "noise"):
        img = image_4k(kind, torch.float32, len(rots), seed=3)
        ref = oracle.equi2equi(img, rots, mode="bilinear", flavour="cuda").numpy()
        out = equilib_b200.equi2equi(img.to(DEV), rots).cpu().numpy()
        d = np.abs(out - ref)
        close = float(np.mean(d <= 1e-5 + 1e-5 * np.abs(ref)))
        record("headline_4k_fp32", image=kind, within_1e5=close, within_5e4=float(np.mean(d <= 5e-4)),
               identical=float(np.mean(d == 0)), max=float(d.max()))
        if kind == "band":
            assert close >= 0.9995
        assert float(np.mean(d <= 5e-4)) >= 0.9995


# ---------------------------------------------------------------------------------------
# (b) config 3: w_face = 1024 cubes, both directions, dice / horizon / dict
# ---------------------------------------------------------------------------------------


def _cube_to_horizon(cube, fmt):
    return oracle.remap_oracle.format_to_horizon(cube, fmt)


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
@pytest.mark.parametrize("fmt", ["dice", "horizon", "dict"])
def test_config3_cube_w1024_vs_oracle(fmt, mode):
    w_face = 1024
    rot = [{"roll": 0.1, "pitch": -0.3, "yaw": 0.7}]
    img = image_4k("band", torch.float32, 1, seed=5)
    ref = oracle.equi2cube(img, rot, w_face, fmt, mode=mode, flavour="cuda")
    out = equilib_b200.equi2cube(img.to(DEV), rot, w_face, fmt, mode=mode)
    # single-item lists come back as [dict] / [list of faces] like the reference
    if fmt == "dict":
        o = torch.stack([out[0][k].cpu() for k in ("F", "R", "B", "L", "U", "D")])
        r = torch.stack([ref[0][k] for k in ("F", "R", "B", "L", "U", "D")])
    else:
        o, r = out.cpu(), ref
    d = (o - r).abs().numpy()
    if mode == "nearest":
        same = float(np.mean(d == 0))
        record("config3_equi2cube", fmt=fmt, mode=mode, identical=same)
        assert same >= 0.9995
    else:
        close = float(np.mean(d <= 1e-5 + 1e-5 * np.abs(r.numpy())))
        record("config3_equi2cube", fmt=fmt, mode=mode, within_1e5=close, max=float(d.max()))
        assert close >= 0.9995

    # and back: the oracle's cube (so both directions are checked independently)
    cube_dev = cases.to_device(ref, DEV)
    back = equilib_b200.cube2equi(cube_dev, fmt, H4K, W4K, mode=mode).cpu().numpy()
    back_ref = oracle.cube2equi(ref, fmt, H4K, W4K, mode=mode, flavour="cuda").numpy()
    db = np.abs(back - back_ref)
    if mode == "nearest":
        same = float(np.mean(db == 0))
        record("config3_cube2equi", fmt=fmt, mode=mode, identical=same)
        assert same >= 0.99999  # (no transcendental on the device: bit-exact coordinates)
    else:
        same = float(np.mean(db == 0))
        record("config3_cube2equi", fmt=fmt, mode=mode, identical=same, max=float(db.max()))
        assert float(db.max()) <= 1e-6 and same >= 0.999


def test_config3_cube_w1024_uint8_dice():
    rot = [{"roll": -0.2, "pitch": 0.4, "yaw": -1.3}]
    img = image_4k("noise", torch.uint8, 1, seed=6)
    ref = oracle.equi2cube(img, rot, 1024, "dice", mode="bilinear", flavour="cuda")
    out = equilib_b200.equi2cube(img.to(DEV), rot, 1024, "dice", mode="bilinear").cpu()
    d = u8_diff(out, ref)
    back = equilib_b200.cube2equi(ref.to(DEV), "dice", H4K, W4K, mode="bilinear").cpu()
    back_ref = oracle.cube2equi(ref, "dice", H4K, W4K, mode="bilinear", flavour="cuda")
    db = u8_diff(back, back_ref)
    record("config3_uint8_dice", e2c_identical=float((d == 0).float().mean()),
           e2c_max=int(d.max()), c2e_identical=float((db == 0).float().mean()), c2e_max=int(db.max()))
    assert float((d <= 1).float().mean()) >= 0.9999
    assert float((d == 0).float().mean()) >= 0.9933  # measured 0.9944
    assert int(db.max()) == 0


# ---------------------------------------------------------------------------------------
# (c) config 5: one 4K fp16 equirect expanded to 8 views of 1024 x 1024
# ---------------------------------------------------------------------------------------


def test_config5_multiview_fp16_vs_oracle():
    views = [{"roll": 0.0, "pitch": 0.35 * math.sin(0.7 * i), "yaw": 2 * math.pi * i / 8}
             for i in range(8)]
    equi = image_4k("band", torch.float16, 1, seed=8)
    ref = oracle.equi2pers(equi.float().expand(8, -1, -1, -1), views, 1024, 1024, 90.0,
                           mode="bilinear", flavour="cuda")
    ref16 = ref.to(torch.float16)
    x = equi.to(DEV).expand(8, -1, -1, -1)  # stride-0 batch: one resident equirect
    out = equilib_b200.equi2pers(x, views, 1024, 1024, 90.0, mode="bilinear").cpu()
    out_exact = equilib_b200.equi2pers(x, views, 1024, 1024, 90.0, mode="bilinear",
                                       fast_coords=False).cpu()
    assert out.dtype == torch.float16
    ulp = 2.0 ** -11  # fp16 ulp of values in [0.5, 1)
    d = (out.float() - ref16.float()).abs()
    de = (out_exact.float() - ref16.float()).abs()
    record("config5_multiview_fp16", identical_default=float((d == 0).float().mean()),
           identical_exact=float((de == 0).float().mean()),
           within1ulp_default=float((d <= ulp * 1.01).float().mean()), max_default=float(d.max()))
    assert float((d <= ulp * 1.01 + 1e-5).float().mean()) >= 0.9999
    assert float((d == 0).float().mean()) >= 0.996   # measured 0.9973
    assert float((de == 0).float().mean()) >= 0.9985  # measured 0.9995


# ---------------------------------------------------------------------------------------
# (d) config 4: 1080p <-> 8K uint8 bicubic, one frame each way
# ---------------------------------------------------------------------------------------


def test_config4_pers2equi_1080p_to_8k_uint8_bicubic():
    rot = [{"roll": 0.05, "pitch": 0.2, "yaw": -0.4}]
    pers = torch.from_numpy(cases.band_limited_image((1, 3, 1080, 1920), "uint8", 9, px=240.0,
                                                     py=180.0))
    ref = oracle.pers2equi(pers, rot, 4096, 8192, 90.0, mode="bicubic", flavour="cuda")
    out = equilib_b200.pers2equi(pers.to(DEV), rot, 4096, 8192, 90.0, mode="bicubic").cpu()
    d = u8_diff(out, ref)
    same = float((d == 0).float().mean())
    # inside the field of view only (the fill is trivially equal)
    inside = ref != ref.min()
    same_in = float((d[inside] == 0).float().mean()) if bool(inside.any()) else 1.0
    record("config4_pers2equi_8k_u8_bicubic", identical=same, identical_inside_fov=same_in,
           max=int(d.max()), inside_share=float(inside.float().mean()))
    assert float((d <= 1).float().mean()) >= 0.99999 and same_in >= 0.9988  # measured 0.9998


def test_config4_equi2pers_8k_to_1080p_uint8_bicubic():
    rot = [{"roll": 0.05, "pitch": 0.2, "yaw": -0.4}]
    equi = torch.from_numpy(cases.band_limited_image((1, 3, 4096, 8192), "uint8", 10))
    ref = oracle.equi2pers(equi, rot, 1080, 1920, 90.0, mode="bicubic", flavour="cuda")
    out = equilib_b200.equi2pers(equi.to(DEV), rot, 1080, 1920, 90.0, mode="bicubic").cpu()
    d = u8_diff(out, ref)
    record("config4_equi2pers_8k_u8_bicubic", identical=float((d == 0).float().mean()),
           max=int(d.max()))
    assert float((d <= 1).float().mean()) >= 0.9999
    assert float((d == 0).float().mean()) >= 0.9965  # measured 0.9975


# ---------------------------------------------------------------------------------------
# (e) StaticRemap against the ORACLE (its old test compared it with the
#     product's own exact path only)
# ---------------------------------------------------------------------------------------


@pytest.mark.parametrize("compact", [False, True], ids=["map8", "map4"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.uint8], ids=["bf16", "u8"])
def test_static_remap_4k_vs_oracle(dtype, compact):
    """8-byte map (16-bit fractions) and the compact 4-byte map (tap relative to the tile's
    box, 8-bit fractions: coordinates to 1/510 pixel) against the oracle."""
    rots = HEADLINE_ROTS[1:3] + HEADLINE_ROTS[4:]
    img = image_4k("band", dtype, len(rots), seed=11)
    ref = oracle.equi2equi(img, rots, mode="bilinear", flavour="cuda")
    x = img.to(DEV)
    sr = equilib_b200.StaticRemap.equi2equi(rots, x.shape, dtype, DEV, compact=compact)
    assert sr.map.numel() * 4 == x.shape[0] * H4K * W4K * (4 if compact else 8)
    out = sr(x).cpu()
    if dtype == torch.uint8:
        d = u8_diff(out, ref)
        tol = 1
    else:
        d = (out.float() - ref.float()).abs()
        tol = 2.0 ** -7 * 1.01
    same = float((d == 0).float().mean())
    record("static_remap_4k", dtype=str(dtype), compact=compact, identical=same,
           within_tol=float((d <= tol).float().mean()), max=float(d.max()))
    assert float((d <= tol).float().mean()) >= 0.9999
    if compact:  # measured: see profiles/r02_parity_rates.jsonl
        assert same >= (0.9915 if dtype == torch.uint8 else 0.996)  # measured 0.9928 / 0.9973
    else:
        assert same >= (0.9925 if dtype == torch.uint8 else 0.9989)  # measured 0.9935 / 0.9999
    # perspective views through a map, against the oracle as well
    sp = equilib_b200.StaticRemap.equi2pers(rots, x.shape, dtype, DEV, 1024, 1024, 90.0,
                                            compact=compact)
    ref_p = oracle.equi2pers(img, rots, 1024, 1024, 90.0, mode="bilinear", flavour="cuda")
    op = sp(x).cpu()
    dp = u8_diff(op, ref_p) if dtype == torch.uint8 else (op.float() - ref_p.float()).abs()
    record("static_remap_4k_pers", dtype=str(dtype), compact=compact,
           identical=float((dp == 0).float().mean()),
           within_tol=float((dp <= tol).float().mean()))
    assert float((dp <= tol).float().mean()) >= 0.9999


# ---------------------------------------------------------------------------------------
# (f) nearest at 4K from interpolated coordinates + tie guard: identical to the exact chain
# ---------------------------------------------------------------------------------------


@pytest.mark.parametrize("dtype", ["bfloat16", "uint8"])
def test_nearest_4k_tie_guard_equals_the_exact_chain(dtype):
    """8/16-bit nearest on large launches rounds INTERPOLATED coordinates and re-evaluates the
    exact chain for every pixel within 4e-3 px of a rounding tie (remap_lean2.cuh).  On a noise
    image at the headline size -- where the interpolation error is largest (1.7e-3 px) -- the
    result must equal the exact chain's on every pixel, and with it the oracle's wherever the
    coordinate is no tie."""
    rots = cases.bench_rotations(32)
    rots = [rots[7], rots[19], {"roll": 0.1, "pitch": math.pi / 2 - 0.02, "yaw": 1.0},
            {"roll": 2.1, "pitch": -1.0, "yaw": -2.0}]
    img = torch.from_numpy(cases.noise_image((len(rots), 3, 2048, 4096),
                                             "uint8" if dtype == "uint8" else "float32", 21))
    x = img.to(DEV) if dtype == "uint8" else img.to(DEV).to(torch.bfloat16)
    out = equilib_b200.equi2equi(x, rots, mode="nearest")
    exact = equilib_b200.equi2equi(x, rots, mode="nearest", fast_coords=False)
    same = float((out == exact).float().mean())
    record(f"nearest_4k_tie_guard_{dtype}", identical_to_exact_chain=same)
    assert torch.equal(out, exact), f"{int((out != exact).sum())} elements differ"
    pers = equilib_b200.equi2pers(x, rots, 1024, 1024, 90.0, mode="nearest")
    pers_exact = equilib_b200.equi2pers(x, rots, 1024, 1024, 90.0, mode="nearest",
                                        fast_coords=False)
    assert torch.equal(pers, pers_exact)
    ref = oracle.equi2equi(x[:1].cpu(), rots[:1], mode="nearest", flavour="cuda")
    assert float((out[:1].cpu() == ref).float().mean()) >= 0.9995
    if dtype == "uint8":
        # an 8K source: an fp32 coordinate resolves 1e-3 px there, the guard scales with it
        big = torch.from_numpy(cases.noise_image((1, 3, 4096, 8192), "uint8", 22)).to(DEV)
        for r in (rots[1], rots[3]):
            a = equilib_b200.equi2equi(big, [r], mode="nearest")
            e = equilib_b200.equi2equi(big, [r], mode="nearest", fast_coords=False)
            assert torch.equal(a, e), f"8K: {int((a != e).sum())} elements differ"
            a = equilib_b200.equi2pers(big, [r], 1080, 1920, 90.0, mode="nearest")
            e = equilib_b200.equi2pers(big, [r], 1080, 1920, 90.0, mode="nearest",
                                       fast_coords=False)
            assert torch.equal(a, e)
