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


def test_descriptor_mirrors_match_the_header_offsets():
    """offsetof of every flou_b200_desc field (tests/abi_offsets.c, compiled here with gcc) against
    the ctypes mirror and the generated Julia table the un-run shim asserts at load time."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "gen_desc_offsets", os.path.join(ROOT, "flou.jl_b200", "julia", "gen_desc_offsets.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    rows, size = gen.offsets()
    assert size == C.sizeof(L.Desc)
    assert [r[0] for r in rows] == [f[0] for f in L.Desc._fields_]
    for name, off, sz in rows:
        assert getattr(L.Desc, name).offset == off, name
        assert getattr(L.Desc, name).size == sz, name
    committed = open(os.path.join(ROOT, "flou.jl_b200", "julia", "desc_offsets.jl")).read()
    assert committed == gen.render(rows, size), "stale desc_offsets.jl: run flou.jl_b200/julia/gen_desc_offsets.py"
    # the Julia struct lists the same fields in the same order
    jl = open(os.path.join(ROOT, "flou.jl_b200", "julia", "FlouB200.jl")).read()
    body = jl[jl.index("struct Desc"):jl.index("\nend", jl.index("struct Desc"))]
    fields = re.findall(r"(\w+)::", body)
    assert fields == [r[0] for r in rows]


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


def test_descriptor_value_ranges_are_checked():
    """Byte / enum tables: an unknown bc_kind, an orientation code outside the dimension's range, a
    boundary-face id outside [1, nf] or decreasing offsets are refused (EINVAL), not read blindly."""
    case = Case(2, (3, 3), 4, periodic=[("3", "4")], bcs={"1": ("slip", None), "2": ("outflow", None)})

    def rc_after(mutate):
        disc, _ = case.product(create=False)
        mutate(disc._keep)
        return L.lib().flou_b200_partition_plan(C.byref(disc._desc), None, None, None, None, None, None,
                                                None, None)
    assert rc_after(lambda k: None) == L.OK
    assert rc_after(lambda k: k["bc_kind"].__setitem__(0, 7)) == L.EINVAL
    assert rc_after(lambda k: k["bc_kind"].__setitem__(1, -1)) == L.EINVAL
    assert rc_after(lambda k: k["orientation"].__setitem__(2, 2)) == L.EINVAL      # 2-D: 0..1
    assert rc_after(lambda k: k["bc_faces"].__setitem__(0, 0)) == L.EINVAL
    assert rc_after(lambda k: k["bc_faces"].__setitem__(0, 10 ** 6)) == L.EINVAL
    assert rc_after(lambda k: k["bc_offsets"].__setitem__(1, k["bc_offsets"][2] + 1)) == L.EINVAL
    disc3, _ = Case(3, (2, 2, 2), 3).product(create=False)
    disc3._keep["orientation"][0] = 8
    assert L.lib().flou_b200_partition_plan(C.byref(disc3._desc), None, None, None, None, None, None,
                                            None, None) == L.EINVAL
