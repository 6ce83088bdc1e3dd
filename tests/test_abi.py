"""The C-ABI library: loads, exports every symbol include/flou_b200.h declares, agrees on the
descriptor layout, validates descriptors, and refuses to run without a CUDA device (no CPU
fallback).  No compute calls: CPU only."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import flou_b200 as F
from flou_b200 import _lib as L
from common import Case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "flou_b200.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(flou_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(L.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/flou_b200.h but not exported"


def test_python_binding_covers_the_header():
    assert sorted(L.SYMBOLS) == _declared_functions()


def test_descriptor_layout_matches_the_library():
    # the library rejects a descriptor whose struct_size differs from its own sizeof
    case = Case(2, (3, 3), 4)
    disc, _ = case.product(create=False)
    d = disc._desc
    n = C.c_int64(0)
    assert L.lib().flou_b200_partition_plan(C.byref(d), C.byref(n), None, None, None, None, None,
                                            None, None) == L.OK
    d.struct_size += 8
    assert L.lib().flou_b200_partition_plan(C.byref(d), C.byref(n), None, None, None, None, None,
                                            None, None) == L.EINVAL
    assert b"ABI" in L.lib().flou_b200_last_error()
    d.struct_size -= 8


@pytest.mark.parametrize("field,value", [("nd", 4), ("np", 9), ("np", 1), ("nv", 3), ("equation", 7),
                                         ("divop", 5), ("numflux", 9), ("tpflux", 3),
                                         ("numflux_avg", 4), ("geometry", 2), ("ne", 0)])
def test_descriptor_validation(field, value):
    disc, _ = Case(2, (3, 3), 4).product(create=False)
    d = disc._desc
    setattr(d, field, value)
    rc = L.lib().flou_b200_partition_plan(C.byref(d), None, None, None, None, None, None, None, None)
    assert rc == L.EINVAL
    with pytest.raises(ValueError):
        L.check(rc)


def test_connectivity_inconsistency_is_rejected():
    disc, _ = Case(2, (3, 3), 4).product(create=False)
    disc._keep["facepos"][0, 0] = 3
    rc = L.lib().flou_b200_partition_plan(C.byref(disc._desc), None, None, None, None, None, None,
                                          None, None)
    assert rc == L.EINVAL


def test_supported_matrix():
    lib = L.lib()
    for nd in (1, 2, 3):
        for npn in range(2, 9):
            for geom in (L.GEOM_CARTESIAN, L.GEOM_GENERAL):
                assert lib.flou_b200_supported(nd, npn, L.EQ_EULER, L.OP_SPLIT, L.FLUX_CHANDRASEKHAR, geom)
                assert lib.flou_b200_supported(nd, npn, L.EQ_EULER, L.OP_STRONG, 0, geom)
                assert lib.flou_b200_supported(nd, npn, L.EQ_LINEAR_ADVECTION, L.OP_STRONG, 0, geom)
    assert not lib.flou_b200_supported(3, 9, L.EQ_EULER, L.OP_STRONG, 0, 0)
    assert not lib.flou_b200_supported(4, 4, L.EQ_EULER, L.OP_STRONG, 0, 0)
    assert not lib.flou_b200_supported(2, 4, L.EQ_LINEAR_ADVECTION, L.OP_SPLIT, L.FLUX_CHANDRASEKHAR, 0)
    # HybridDivOperator (row f2): Euler, either two-point flux, Cartesian or general sub-grids
    for nd in (1, 2, 3):
        for tp in (L.FLUX_STDAVERAGE, L.FLUX_CHANDRASEKHAR):
            assert lib.flou_b200_supported(nd, 4, L.EQ_EULER, L.OP_HYBRID, tp, L.GEOM_CARTESIAN)
    assert lib.flou_b200_supported(2, 4, L.EQ_EULER, L.OP_HYBRID, L.FLUX_CHANDRASEKHAR, L.GEOM_GENERAL)
    assert not lib.flou_b200_supported(2, 4, L.EQ_LINEAR_ADVECTION, L.OP_HYBRID, L.FLUX_STDAVERAGE, 0)
    assert not lib.flou_b200_supported(2, 4, L.EQ_EULER, L.OP_HYBRID, L.FLUX_LXF, 0)


def test_no_cpu_fallback_without_a_device():
    if F.device_count() > 0:
        pytest.skip("a CUDA device is visible on this box")
    with pytest.raises(L.FlouB200Error) as err:
        Case(2, (3, 3), 4).product()
    assert "no usable CUDA device" in str(err.value)
    h = C.c_void_p()
    assert L.lib().flou_b200_rhs(h, None, None, 0.0) == L.EINVAL      # null handle, never computes


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libflou_b200.so")
    with pytest.raises(L.FlouB200Error) as err:
        L.lib()
    assert "no CPU fallback" in str(err.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "flou.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "oracle/" not in text and "liboracle" not in text, f
