"""CPU-only checks of the host side: the C ABI loads and exports what the header
declares, host prep (rotation matrices, tables, A matrices) equals the oracle, and
the public API rejects what it must before any launch."""

import ctypes
import os
import re

import numpy as np
import pytest
import torch

import equilib_b200
from equilib_b200 import _lib, _prep
from oracle import remap_oracle as O
from tests import cases, helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPU = torch.device("cpu")


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "equilib_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(eqb_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), (declared, set(_lib.EXPORTS))
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert _lib.lib().eqb_abi_version() == _lib.ABI_VERSION
    assert _lib.lib().eqb_last_error() == b""


def test_descriptor_validation_without_gpu():
    d = _lib.RemapDesc()
    L = _lib.lib()
    assert L.eqb_remap(ctypes.byref(d), None) == -1
    assert b"batch" in L.eqb_last_error()
    assert L.eqb_remap(None, None) == -1
    d.transform = 99
    assert L.eqb_remap(ctypes.byref(d), None) == -1
    assert b"transform" in L.eqb_last_error()


def test_new_entry_points_validate_without_gpu():
    """Channels-last probe and the static-map entry points reject bad descriptors on the
    host, before any launch."""
    L = _lib.lib()
    d = _lib.RemapDesc()
    assert L.eqb_remap_supports_channels_last(ctypes.byref(d)) == 0
    assert L.eqb_remap_supports_channels_last(None) == 0
    assert L.eqb_last_error() == b""  # a probe, not an error
    assert L.eqb_map_build(ctypes.byref(d), None, None, None, None) == -1
    assert L.eqb_remap_mapped(None, None, None, None) == -1
    # a well-formed equi2equi descriptor whose image type the map path does not serve
    d.transform, d.mode, d.dtype = _lib.EQUI2EQUI, 1, _lib.F32
    d.batch, d.channels = 1, 3
    d.src_h, d.src_w, d.dst_h, d.dst_w = 64, 128, 64, 128
    d.mats = d.tab_x = d.tab_y = 16  # non-null (never dereferenced on the host)
    rc = L.eqb_map_build(ctypes.byref(d), 16, 16, 16, None)
    assert rc == -2 and b"static maps" in L.eqb_last_error()
    assert (_lib.MAP_TILE_W, _lib.MAP_TILE_H) == (64, 8)


def test_struct_layout_matches_header():
    # sizeof(eqb_remap_desc): 10 int32 + 2x(ptr + 3 int64) + 5 ptr + int64 + ptr
    #                         + 2 int32 + 12 int64 + ptr
    assert ctypes.sizeof(_lib.RemapDesc) == 40 + 2 * 32 + 4 * 8 + 8 + 8 + 8 + 8 + 96 + 8
    assert ctypes.sizeof(_lib.Image) == 8 + 5 * 4 + 4 + 3 * 8


def test_rotation_matrices_match_reference_fixture_and_oracle():
    gold = helpers.load_golden("reference_rotations.npz")
    rots = cases.rotation_sweep(257)
    assert np.array_equal(_prep.rotation_matrices_host(rots, False).numpy(), gold["z_up"])
    assert np.array_equal(_prep.rotation_matrices_host(rots, True).numpy(), gold["z_down"])
    more = cases.rotation_sweep(2000, seed=123)
    for z in (False, True):
        assert torch.equal(_prep.rotation_matrices_host(more, z), O.rotation_matrices(more, z))


def test_matrix_prep_matches_oracle():
    rots = cases.rotation_sweep(9)
    a = _prep.mats_for("equi2pers", rots, False, CPU, (48, 64, 90.0, 0.0)).view(-1, 3, 3)
    assert torch.equal(a, O.equi2pers_matrix(rots, False, 48, 64, 90.0, 0.0))
    a = _prep.mats_for("pers2equi", rots, True, CPU, (48, 64, 70.0, 0.1)).view(-1, 3, 3)
    assert torch.equal(a, O.pers2equi_matrix(rots, True, 48, 64, 70.0, 0.1))
    a = _prep.mats_for("rot", rots, True, CPU).view(-1, 3, 3)
    assert torch.equal(a, O.rotation_matrices(rots, True))


def test_tables_match_oracle():
    tx, ty = _prep.equirect_tables(48, 96, CPU)
    cphi, sphi, cth, sth = O._equirect_ray_tables(48, 96)
    assert torch.equal(tx.view(-1, 2)[:, 0], cth) and torch.equal(tx.view(-1, 2)[:, 1], sth)
    assert torch.equal(ty, torch.cat([cphi, sphi]))
    tx, ty, itx = _prep.cube2equi_tables(64, 128, CPU)
    t = O.cube2equi_tables(64, 128)
    assert torch.equal(tx, torch.cat([t["tanx"], t["cosx"], t["sinth"], t["costh"]]))
    assert torch.equal(ty, torch.cat([t["ntan"], t["cu"], t["cd"]]))
    assert torch.equal(itx, torch.cat([t["face"], t["thr"]]).to(torch.int32))


def test_caches_are_bounded_and_reused():
    _prep.clear_caches()
    a = _prep.equirect_tables(8, 16, CPU)
    assert _prep.equirect_tables(8, 16, CPU) is a
    for h in range(8, 8 + 2 * _prep.TABLE_CACHE_MAXSIZE, 1):
        _prep.equirect_tables(h, 16, CPU)
    assert len(_prep._tables) <= _prep.TABLE_CACHE_MAXSIZE
    r1 = cases.rotation_sweep(4)
    m = _prep.mats_for("rot", r1, False, CPU)
    assert _prep.mats_for("rot", [dict(r) for r in r1], False, CPU) is m
    assert _prep.mats_for("rot", r1, True, CPU) is not m


ROT = {"roll": 0.0, "pitch": 0.0, "yaw": 0.0}


def test_api_surface_matches_reference_names():
    for n in ("Equi2Pers", "Equi2Equi", "Equi2Cube", "Cube2Equi", "Pers2Equi", "equi2pers",
              "equi2equi", "equi2cube", "cube2equi", "pers2equi"):
        assert hasattr(equilib_b200, n)
    import inspect

    sig = inspect.signature(equilib_b200.equi2pers)
    assert list(sig.parameters)[:9] == ["equi", "rots", "height", "width", "fov_x", "skew",
                                        "mode", "z_down", "clip_output"]
    sig = inspect.signature(equilib_b200.equi2equi)
    assert list(sig.parameters)[:7] == ["src", "rots", "clip_output", "mode", "z_down", "height",
                                        "width"]
    sig = inspect.signature(equilib_b200.equi2cube)
    assert list(sig.parameters)[:7] == ["equi", "rots", "w_face", "cube_format", "z_down",
                                        "clip_output", "mode"]
    sig = inspect.signature(equilib_b200.cube2equi)
    assert list(sig.parameters)[:6] == ["cubemap", "cube_format", "height", "width",
                                        "clip_output", "mode"]
    sig = inspect.signature(equilib_b200.pers2equi)
    assert list(sig.parameters)[:9] == ["pers", "rots", "height", "width", "fov_x", "skew",
                                        "mode", "z_down", "clip_output"]


def test_no_cpu_fallback():
    x = torch.zeros(3, 8, 16)
    with pytest.raises(ValueError, match="no CPU path"):
        equilib_b200.equi2equi(x, ROT)
    with pytest.raises(ValueError, match="numpy"):
        equilib_b200.equi2pers(np.zeros((3, 8, 16), np.float32), ROT, 4, 4, 90.0)
    with pytest.raises(ValueError):
        equilib_b200.equi2cube("nope", ROT, 4, "dice")
    with pytest.raises(ValueError, match="no CPU path"):
        equilib_b200.cube2equi({k: torch.zeros(3, 4, 4) for k in "FRBLUD"}, "dict", 8, 16)
    with pytest.raises(AssertionError):
        equilib_b200.cube2equi(torch.zeros(3, 4, 24), "horizon", 10, 16)  # not multiples of 8


def test_product_never_imports_oracle_or_grid_sample():
    pkg = os.path.join(ROOT, "equilib_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            code = re.sub(r'""".*?"""', "", src, flags=re.S)
            code = re.sub(r"#.*", "", code)
            assert "oracle" not in code, fn
            assert "F.grid_sample" not in code and "functional" not in code, fn


def test_host_pipeline_chunking():
    from equilib_b200.pipeline import HostPipeline

    assert [list(r) for r in HostPipeline.chunks(10, 4)] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]
    assert [list(r) for r in HostPipeline.chunks(3, 4)] == [[0, 1, 2]]
    assert HostPipeline.chunks(0, 4) == []


def test_library_carries_the_source_hash_of_this_tree():
    """build() decides by content, not by modification time: the hash compiled into the
    library (eqb_source_hash) must be the hash of the sources it sits next to."""
    from equilib_b200 import build

    assert len(build.source_hash()) == 16
    assert build.built_hash() == _lib.source_hash()
    assert _lib.source_hash() == build.source_hash(), "stale libequilib_b200.so: run build()"


def test_backend_native_is_an_alias_and_pure_is_rejected():
    from equilib_b200 import _common

    for b in (None, "b200", "native"):
        kw = {"backend": b, "cache": object()}
        _common.check_backend(kw)
        assert kw == {}
    with pytest.raises(ValueError):
        _common.check_backend({"backend": "pure"})


def test_fused_cube_chain_entry_point_validates_without_gpu():
    """eqb_equi2cube2equi takes the two descriptors of the launches it replaces and rejects
    anything else on the host, before any launch; the Python entry has no CPU path either."""
    L = _lib.lib()
    a, b = _lib.RemapDesc(), _lib.RemapDesc()
    assert L.eqb_equi2cube2equi(None, None, None) == -1
    assert L.eqb_equi2cube2equi(ctypes.byref(a), ctypes.byref(b), None) == -1
    assert b"EQUI2CUBE followed by CUBE2EQUI" in L.eqb_last_error()
    for d in (a, b):
        d.mode, d.dtype, d.batch, d.channels = 1, _lib.F32, 2, 3
        d.w_face, d.n_cells = 16, 6
        d.mats = d.tab_x = d.tab_y = d.itab_x = 16  # non-null, never dereferenced on the host
    a.transform, a.src_h, a.src_w, a.dst_h, a.dst_w = _lib.EQUI2CUBE, 64, 128, 16, 96
    b.transform, b.src_h, b.src_w, b.dst_h, b.dst_w = _lib.CUBE2EQUI, 16, 96, 64, 128
    a.src, a.src_stride_b, a.src_stride_c, a.src_stride_h = 16, 3 * 64 * 128, 64 * 128, 128
    b.mode = 2  # bicubic on one side only
    assert L.eqb_equi2cube2equi(ctypes.byref(a), ctypes.byref(b), None) == -1
    assert b"differ" in L.eqb_last_error()
    a.mode = 2
    assert L.eqb_equi2cube2equi(ctypes.byref(a), ctypes.byref(b), None) == -2
    assert b"nearest and bilinear" in L.eqb_last_error()
    a.mode = b.mode = 1
    assert L.eqb_equi2cube2equi(ctypes.byref(a), ctypes.byref(b), None) == -1  # dst is null
    assert b"null" in L.eqb_last_error()
    assert hasattr(equilib_b200, "Equi2Cube2Equi") and hasattr(equilib_b200, "equi2cube2equi")
    with pytest.raises(ValueError, match="no CPU path"):
        equilib_b200.equi2cube2equi(torch.zeros(3, 8, 16), ROT, 4, 8, 16)


def test_environment_switches_follow_os_environ(monkeypatch):
    """flags_of reads the EQUILIB_B200_* switches from the bytes map behind os.environ (a
    launch's host path must not pay ten os.environ.get calls): it has to see what
    monkeypatch / os.environ assignments do, immediately."""
    from equilib_b200 import _common as K

    opt = K.Options()
    base = K.flags_of(torch.uint8, opt)
    assert base == 0
    monkeypatch.setenv("EQUILIB_B200_NO_TMA", "1")
    assert K.flags_of(torch.uint8, opt) == _lib.FLAG_NO_TMA
    monkeypatch.setenv("EQUILIB_B200_NO_LEAN2", "1")
    monkeypatch.setenv("EQUILIB_B200_FAST_COORDS", "0")
    assert K.flags_of(torch.uint8, opt) == (_lib.FLAG_NO_TMA | _lib.FLAG_NO_LEAN2
                                            | _lib.FLAG_EXACT_COORDS)
    monkeypatch.delenv("EQUILIB_B200_NO_TMA")
    monkeypatch.delenv("EQUILIB_B200_NO_LEAN2")
    monkeypatch.setenv("EQUILIB_B200_FAST_COORDS", "1")
    assert K.flags_of(torch.float32, opt) == _lib.FLAG_FAST_COORDS
    assert K.flags_of(torch.float32, K.Options(fast_coords=False)) == _lib.FLAG_EXACT_COORDS
    monkeypatch.delenv("EQUILIB_B200_FAST_COORDS")
    monkeypatch.setenv("EQUILIB_B200_WARP_SHAPE", "2")
    assert K.flags_of(torch.uint8, opt) == 2 << _lib.FLAG_WARP_SHAPE_SHIFT
    monkeypatch.delenv("EQUILIB_B200_WARP_SHAPE")
    assert K.flags_of(torch.uint8, K.Options(saturate_uint8=True)) == _lib.FLAG_SATURATE_U8
