"""The C-ABI library loads on a CPU-only box and exports exactly what include/oetr_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT
from oetr_b200 import cabi, weights


def _declared():
    text = open(os.path.join(ROOT, "include", "oetr_b200.h")).read()
    return sorted(set(re.findall(r"OETR_API\s+[\w\s\*]+?\b(oetr_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared() == sorted(cabi.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = cabi.load_library()
    for name in _declared():
        assert getattr(lib, name) is not None, name
    assert lib.oetr_abi_version() == 1
    assert lib.oetr_packed_weight_count() == weights.PACKED_COUNT == 6443525


def test_argument_errors_do_not_need_a_gpu():
    lib = cabi.load_library()
    h = ctypes.c_void_p()
    assert lib.oetr_create(None, 0, 0, 0, 0, 100, 100, ctypes.byref(h)) == cabi.OETR_E_ARG
    assert b"null" in lib.oetr_last_error()
    buf = (ctypes.c_float * 4)()
    rc = lib.oetr_create(ctypes.cast(buf, ctypes.c_void_p), 4, 0, 7, 0, 100, 100, ctypes.byref(h))
    assert rc == cabi.OETR_E_ARG and b"attention_mode" in lib.oetr_last_error()
    rc = lib.oetr_create(ctypes.cast(buf, ctypes.c_void_p), 4, 0, 0, 0, 100, 100, ctypes.byref(h))
    assert rc == cabi.OETR_E_ARG and b"packed floats" in lib.oetr_last_error()
    assert lib.oetr_destroy(None) == 0


def test_encoder_tile_geometry_is_consistent():
    """oetr_selftest_geometry (host only): flat and per-image encoder tilings cover every token exactly once and
    the per-image gather of partial summaries matches the tiles, over a range of batch sizes and map geometries."""
    lib = cabi.load_library()
    n = ctypes.c_int()
    expect_flat = {(32, 20, 20, 20, 20): 200, (8, 20, 20, 20, 20): 50, (16, 26, 26, 26, 26): 172, (1, 20, 20, 20, 20): 8}
    for b in (1, 2, 3, 5, 8, 16, 21, 32):
        for fm1, fm2 in (((20, 20), (20, 20)), ((26, 26), (26, 26)), ((20, 20), (15, 20)), ((14, 17), (20, 20)),
                         ((8, 16), (16, 8)), ((12, 12), (11, 12)), ((100, 100), (2, 2)), ((1, 1), (1, 1)), ((7, 13), (11, 5)),
                         ((19, 19), (19, 19)), ((100, 100), (100, 100))):
            rc = lib.oetr_selftest_geometry(b, fm1[0], fm1[1], fm2[0], fm2[1], ctypes.byref(n))
            assert rc == 0, (b, fm1, fm2, lib.oetr_last_error())
            flat = fm1[0] * fm1[1] >= 128 and fm2[0] * fm2[1] >= 128
            assert (n.value > 0) == flat, (b, fm1, fm2, n.value)
            key = (b, fm1[0], fm1[1], fm2[0], fm2[1])
            if key in expect_flat:
                assert n.value == expect_flat[key], (key, n.value)
    assert lib.oetr_selftest_geometry(0, 1, 1, 1, 1, None) == cabi.OETR_E_ARG
