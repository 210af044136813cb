"""ctypes binding of libfieldconv_b200.so (the C ABI in include/fieldconv_b200.h).

The product path has no CPU fallback: if the library is missing or a call fails this raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfieldconv_b200.so")

GEMM_SIMT_FP32 = 0
GEMM_TC_3XTF32 = 1
GEMM_TC_TF32 = 2
GEMM_TC_2XF16 = 3
GEMM_MASK = 0xff
FLAG_HAVE_CONTRIB = 0x100
# Python-level only (ops.py): run the layer through fcb_fwd_pk_f32 / fcb_bwd_pk_f32 (packed fp16 operand planes written by
# the aggregation kernels, bulk-copied by the contraction kernels); never passed to the library
FLAG_PACKED = 0x200
# Python-level only: backward through fcb_bwd_pk_f32 with a packed G even when the forward kept the fp32 contrib layout
FLAG_PACKED_G = 0x800
# Python-level only: run the forward through fcb_fwd_fused_f32 (band_limit <= 1: contrib never leaves the SM)
FLAG_FUSED = 0x400

_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_I = ctypes.c_int
_F = ctypes.c_float
_SZ = ctypes.c_size_t
_PSZ = ctypes.POINTER(ctypes.c_size_t)

# name -> argtypes (all return int unless noted); mirrors include/fieldconv_b200.h
SIGNATURES = {
    "fcb_version": [],
    "fcb_plan_workspace_bytes": [_I64, _I64, _I, _PSZ],
    "fcb_plan_build": [_P, _P, _P, _P, _P, _P, _F, _I64, _I64, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P],
    "fcb_plan_dense_workspace_bytes": [_I64, _I64, _PSZ],
    "fcb_plan_build_dense": [_P, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _SZ, _P],
    "fcb_precomp_workspace_bytes": [_I64, _PSZ],
    "fcb_precomp_expand_f32": [_P, _P, _P, _F, _I64, _I64, _I, _I, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _SZ, _P],
    "fcb_fwd_workspace_bytes": [_I64, _I, _I, _I, _I, _I, _PSZ],
    "fcb_fwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _I, _I, _I, _P, _SZ, _P],
    "fcb_bound_f32": [_P, _I64, _P, _P],
    "fcb_fwd_act_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _I, _I, _I, _P, _SZ, _P],
    "fcb_fwd_act_pk_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _I, _I, _I, _P, _SZ, _P],
    "fcb_bwd_workspace_bytes": [_I64, _I, _I, _I, _I, _I, _PSZ],
    "fcb_bwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _I, _I, _I, _P, _SZ, _P],
    "fcb_fwd_dense_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _I, _I, _I, _P, _SZ, _P],
    "fcb_bwd_dense_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _I, _I, _I, _P, _SZ, _P],
    "fcb_plan_norm": [_P, _P, _I64, _P, _P],
    "fcb_pk_contrib_bytes": [_I64, _I, _I, _I, _PSZ],
    "fcb_fwd_pk_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _I, _I, _I, _P, _SZ, _P],
    "fcb_bwd_pk_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _I, _I, _I, _P, _SZ, _P],
    "fcb_fwd_fused_workspace_bytes": [_I, _I, _I, _I, _PSZ],
    "fcb_fwd_fused_f32": [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I, _I, _I, _I, _P, _SZ, _P],
    "fcb_lift_aggregate_f32": [_P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _P],
    "fcb_lift_aggregate_bwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _P],
    "fcb_echo_fwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _I, _P],
    "fcb_echo_bwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _I, _I, _P],
    "fcb_radius_workspace_bytes": [_I64, _PSZ],
    "fcb_radius_count": [_P, _I64, _F, _I, _F, _F, _F, _P, _P, _SZ, _P],
    "fcb_radius_fill": [_P, _I64, _F, _I, _F, _F, _F, _P, _P, _P, _SZ, _P],
    "fcb_aggregate_f32": [_P, _P, _P, _P, _P, _I64, _I, _I, _I, _I, _P],
    "fcb_gemm_workspace_bytes": [_I64, _I, _I64, _I, _I, _I, _I, _PSZ],
    "fcb_gemm_tc_feasible": [_I, _I64, _I, _I, _I],
    "fcb_gemm_f32": [_P, _P, _P, _I64, _I, _I64, _I64, _I64, _I64, _I, _I, _I64, _I64, _I64, _I, _P, _P, _P, _SZ, _I, _P],
    "fcb_sort_workspace_bytes": [_I64, _PSZ],
    "fcb_sort_pairs_u32": [_P, _P, _P, _P, _I64, _I, _P, _SZ, _P],
    "fcb_modrelu_fwd_f32": [_P, _P, _P, _I64, _I, _P],
    "fcb_modrelu_bwd_workspace_bytes": [_I64, _I, _PSZ],
    "fcb_modrelu_bwd_f32": [_P, _P, _P, _P, _P, _P, _I64, _I, _P, _SZ, _P],
    "fcb_profile_enable": [_I],
    "fcb_profile_disable": [],
    "fcb_profile_collect": [ctypes.c_char_p, _SZ, ctypes.POINTER(ctypes.c_float), _I, ctypes.POINTER(_I)],
}

_lib = None


def launch_count():
    """Kernels launched by the library so far in this process (fcb_launch_count)."""
    return int(load().fcb_launch_count())


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "fieldconv_b200: %s is missing — build it with `python -m fieldconv_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = ctypes.c_int
        lib.fcb_pk_supported.argtypes = [_I64, _I, _I, _I, _I]
        lib.fcb_pk_supported.restype = ctypes.c_int
        lib.fcb_fused_supported.argtypes = [_I, _I, _I, _I]
        lib.fcb_fused_supported.restype = ctypes.c_int
        lib.fcb_last_error.argtypes = []
        lib.fcb_last_error.restype = ctypes.c_char_p
        lib.fcb_launch_count.argtypes = []
        lib.fcb_launch_count.restype = ctypes.c_ulonglong
        _lib = lib
    return _lib


def call(name, *args):
    """Invoke an int-returning entry point; raise with the library's message on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, lib.fcb_last_error().decode()))
    return rc


def query_bytes(name, *args):
    out = ctypes.c_size_t(0)
    call(name, *args, ctypes.byref(out))
    return int(out.value)


def tc_feasible(n, k, trans_a=0, split_k=1, flags=GEMM_TC_3XTF32):
    return bool(load().fcb_gemm_tc_feasible(int(n), int(k), int(trans_a), int(split_k), int(flags)))


def pk_supported(n, ci, co, band_limit, n_rings):
    """Whether the packed-operand path (fcb_fwd_pk_f32 / fcb_bwd_pk_f32) takes this layer shape."""
    return bool(load().fcb_pk_supported(int(n), int(ci), int(co), int(band_limit), int(n_rings)))


def fused_supported(ci, co, band_limit, n_rings):
    """Whether the fused forward (fcb_fwd_fused_f32) takes this layer shape."""
    return bool(load().fcb_fused_supported(int(ci), int(co), int(band_limit), int(n_rings)))


def ptr(t):
    return 0 if t is None else t.data_ptr()


class Bounds(ctypes.Structure):
    """struct fcb_bounds of the header: optional operand bounds (device pointers to one float, 0 = not available)."""
    _fields_ = [("x", ctypes.c_void_p), ("gy", ctypes.c_void_p), ("act", ctypes.c_void_p), ("w", ctypes.c_void_p)]


def bounds(x=None, gy=None, act=None, w=None):
    """-> argument for a `const fcb_bounds*` parameter (None when nothing is known: the library then computes what it needs)."""
    if x is None and gy is None and act is None and w is None:
        return None
    return ctypes.byref(Bounds(ptr(x) or None, ptr(gy) or None, ptr(act) or None, ptr(w) or None))


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def profile_enable(max_records=65536):
    call("fcb_profile_enable", max_records)


def profile_collect(max_records=65536):
    """-> list of (kernel name, ms) in launch order; call after torch.cuda.synchronize()."""
    names = ctypes.create_string_buffer(max_records * 24)
    ms = (ctypes.c_float * max_records)()
    cnt = ctypes.c_int(0)
    call("fcb_profile_collect", names, len(names), ms, max_records, ctypes.byref(cnt))
    call("fcb_profile_disable")
    lst = names.value.decode().split("\n")[:cnt.value]
    return [(lst[i], float(ms[i])) for i in range(cnt.value)]
