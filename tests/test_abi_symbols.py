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
