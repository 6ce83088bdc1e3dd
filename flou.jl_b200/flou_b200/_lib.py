"""ctypes binding of libflou_b200.so (C ABI in include/flou_b200.h).

This is the only door into the compute path: if the CUDA library is missing, fails to load
or finds no device, everything here raises -- there is no CPU fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FLOU_B200_LIB: alternative build of the same library (kernel-tuning experiments only)
LIB_PATH = os.environ.get("FLOU_B200_LIB") or os.path.join(_HERE, "libflou_b200.so")

OK, EINVAL, ECUDA, ENCCL, EDOMAIN, EUNSUPPORTED = range(6)

EQ_LINEAR_ADVECTION, EQ_EULER = 0, 1
OP_STRONG, OP_SPLIT, OP_HYBRID = 0, 1, 2
FLUX_STDAVERAGE, FLUX_LXF, FLUX_CHANDRASEKHAR, FLUX_SCALARDISSIPATION, FLUX_MATRIXDISSIPATION = range(5)
BC_INFLOW, BC_OUTFLOW, BC_SLIP, BC_TABLE = range(4)
GEOM_CARTESIAN, GEOM_GENERAL = 0, 1
MONITOR_KINETIC_ENERGY, MONITOR_ENTROPY = 0, 1
FLAG_NO_GRAPH = 1
FLAG_FUSED = 2
FLAG_NODE_KERNEL = 4
FLAG_LINE_KERNEL = 8


class DomainError(ArithmeticError):
    """Julia's DomainError: the state left the admissible set (FlouTime.jl:37-52)."""


class FlouB200Error(RuntimeError):
    pass


class Desc(C.Structure):
    """flou_b200_desc (include/flou_b200.h)."""
    _fields_ = [
        ("struct_size", C.c_int32), ("nd", C.c_int32), ("nv", C.c_int32), ("np", C.c_int32),
        ("equation", C.c_int32), ("divop", C.c_int32), ("tpflux", C.c_int32),
        ("numflux", C.c_int32), ("numflux_avg", C.c_int32), ("geometry", C.c_int32),
        ("intensity", C.c_double), ("gamma", C.c_double),
        ("a", C.c_double * 3), ("dx", C.c_double * 3),
        ("ne", C.c_int64), ("nf", C.c_int64),
        ("faceinds", C.c_void_p), ("facepos", C.c_void_p), ("eleminds", C.c_void_p),
        ("elempos", C.c_void_p), ("orientation", C.c_void_p),
        ("D", C.c_void_p), ("Ds", C.c_void_p), ("Dsharp", C.c_void_p),
        ("lminus", C.c_void_p), ("lplus", C.c_void_p),
        ("dgminus", C.c_void_p), ("dgplus", C.c_void_p), ("weights", C.c_void_p),
        ("jac", C.c_void_p), ("metric", C.c_void_p), ("fjac", C.c_void_p), ("frames", C.c_void_p),
        ("nbound", C.c_int32), ("bc_kind", C.c_void_p), ("bc_offsets", C.c_void_p),
        ("bc_faces", C.c_void_p), ("bc_state", C.c_void_p), ("bc_table", C.c_void_p),
        ("elem_begin", C.c_int64), ("elem_end", C.c_int64),
        ("rank", C.c_int32), ("nranks", C.c_int32), ("part_offsets", C.c_void_p),
        ("device", C.c_int32), ("flags", C.c_int32), ("blend", C.c_double),
        ("sub_frames", C.c_void_p), ("sub_jac", C.c_void_p),
    ]


# every symbol include/flou_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "flou_b200_create": (C.c_int32, [C.POINTER(Desc), C.POINTER(C.c_void_p)]),
    "flou_b200_destroy": (C.c_int32, [C.c_void_p]),
    "flou_b200_upload_state": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "flou_b200_download_state": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "flou_b200_rhs": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]),
    "flou_b200_lsrk2n_advance": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_double, C.c_double, C.c_int64]),
    "flou_b200_timeintegrate": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                            C.c_int64]),
    "flou_b200_max_dt": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_double, C.POINTER(C.c_double)]),
    "flou_b200_monitor": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_double)]),
    "flou_b200_zhang_shu": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_double]),
    "flou_b200_set_stage_limiter": (C.c_int32, [C.c_void_p, C.c_int32, C.c_double]),
    "flou_b200_set_source": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "flou_b200_set_bc_table": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "flou_b200_boundary_traces": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "flou_b200_lsrk2n_stage": (C.c_int32, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int32]),
    "flou_b200_project_equispaced": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "flou_b200_synchronize": (C.c_int32, [C.c_void_p]),
    "flou_b200_status": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32)]),
    "flou_b200_last_error": (C.c_char_p, []),
    "flou_b200_ndofs_local": (C.c_int64, [C.c_void_p]),
    "flou_b200_stream": (C.c_void_p, [C.c_void_p]),
    "flou_b200_device_state": (C.c_void_p, [C.c_void_p]),
    "flou_b200_kernel_launches": (C.c_int64, [C.c_void_p]),
    "flou_b200_kernel_info": (C.c_int32, [C.c_void_p] + [C.POINTER(C.c_int32)] * 4),
    "flou_b200_profile": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                      C.POINTER(C.c_int64)]),
    "flou_b200_timer_start": (C.c_int32, [C.c_void_p]),
    "flou_b200_timer_stop": (C.c_int32, [C.c_void_p, C.POINTER(C.c_float)]),
    "flou_b200_pin_host": (C.c_int32, [C.c_void_p, C.c_uint64]),
    "flou_b200_unpin_host": (C.c_int32, [C.c_void_p]),
    "flou_b200_device_count": (C.c_int32, []),
    "flou_b200_supported": (C.c_int32, [C.c_int32] * 6),
    "flou_b200_partition_plan": (C.c_int32, [C.POINTER(Desc), C.POINTER(C.c_int64),
                                             C.POINTER(C.c_int32), C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.POINTER(C.c_int64),
                                             C.POINTER(C.c_int64)]),
    "flou_b200_nccl_unique_id": (C.c_int32, [C.c_char_p]),
    "flou_b200_comm_init": (C.c_int32, [C.c_void_p, C.c_char_p]),
}

_lib = None


def lib():
    """Load the CUDA library; loud failure if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FlouB200Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (make -C flou.jl_b200/csrc). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc):
    if rc == OK:
        return
    msg = lib().flou_b200_last_error().decode(errors="replace")
    if rc in (EINVAL, EUNSUPPORTED):
        raise ValueError(msg)           # Julia: ArgumentError
    if rc == EDOMAIN:
        raise DomainError(msg)
    raise FlouB200Error(f"flou_b200 error {rc}: {msg}")


def device_count():
    return int(lib().flou_b200_device_count())
