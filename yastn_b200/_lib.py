"""ctypes binding of the C ABI declared in include/yastn_b200.h.

There is deliberately no fallback: if libyastn_b200.so is missing or a symbol cannot be resolved the
import raises, and every call converts a non-zero status into RuntimeError(yb_last_error()).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libyastn_b200.so")

# symbol -> (restype, argtypes); must list every function of include/yastn_b200.h (tests check this)
_c = ctypes
_vp, _i64p = _c.c_void_p, _c.POINTER(_c.c_int64)
SIGNATURES = {
    "yb_abi_version": (_c.c_int, []),
    "yb_last_error": (_c.c_char_p, []),
    "yb_copy_plan_create": (_c.c_int, [_vp, _c.c_int64, _c.c_int, _c.c_int, _c.c_int, _c.POINTER(_vp)]),
    "yb_copy_plan_info": (_c.c_int, [_vp, _i64p]),
    "yb_copy_run": (_c.c_int, [_vp, _vp, _vp, _c.c_int64, _c.c_int, _vp]),
    "yb_copy_plan_destroy": (None, [_vp]),
    "yb_gemm_plan_create": (_c.c_int, [_vp, _c.c_int64, _vp, _c.c_int64, _c.c_int, _c.c_int, _c.POINTER(_vp)]),
    "yb_gemm_plan_create_scatter": (_c.c_int, [_vp, _c.c_int64, _vp, _c.c_int64, _vp, _c.c_int64, _vp, _vp, _vp, _vp, _vp, _vp,
                                               _c.c_int, _c.c_int, _c.POINTER(_vp)]),
    "yb_gemm_plan_info": (_c.c_int, [_vp, _i64p]),
    "yb_gemm_run": (_c.c_int, [_vp, _vp, _vp, _vp, _c.c_int, _vp]),
    "yb_gemm_plan_destroy": (None, [_vp]),
    "yb_ew_plan_create": (_c.c_int, [_vp, _c.c_int64, _vp, _c.c_int64, _c.c_int, _c.c_int, _c.POINTER(_vp)]),
    "yb_ew_plan_info": (_c.c_int, [_vp, _i64p]),
    "yb_ew_run": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "yb_ew_plan_destroy": (None, [_vp]),
    "yb_svd_plan_create": (_c.c_int, [_vp, _c.c_int64, _c.c_int, _c.c_int, _c.POINTER(_vp)]),
    "yb_svd_run": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _c.c_int, _c.c_int, _vp]),
    "yb_svd_plan_destroy": (None, [_vp]),
    "yb_tables_result_size": (_c.c_int64, []),
    "yb_tables_result_fetch": (_c.c_int, [_vp, _c.c_int64]),
    "yb_tables_merge": (_c.c_int, [_vp, _c.c_int64, _c.c_int64, _vp, _c.c_int64, _c.c_int64, _vp, _c.c_int, _c.c_int, _c.c_int, _vp, _c.c_int]),
    "yb_tables_scatter": (_c.c_int, [_vp, _c.c_int64, _c.c_int, _vp, _c.c_int64, _vp]),
    "yb_tables_add": (_c.c_int, [_vp, _c.c_int64, _c.c_int64, _vp]),
    "yb_chain_create": (_c.c_int, [_vp, _c.c_int64, _c.c_int64, _c.POINTER(_vp)]),
    "yb_chain_run": (_c.c_int, [_vp, _vp, _c.c_int64, _vp]),
    "yb_chain_steps": (_c.c_int64, [_vp]),
    "yb_chain_destroy": (None, [_vp]),
    "yb_peer_alloc": (_c.c_int, [_c.c_int64, _c.c_int, _c.POINTER(_vp), _vp]),
    "yb_peer_open": (_c.c_int, [_vp, _c.c_int, _c.POINTER(_vp)]),
    "yb_peer_close": (_c.c_int, [_vp]),
    "yb_peer_free": (_c.c_int, [_vp]),
}

ABI_VERSION = 4
YB_F64, YB_C128 = 0, 1
YB_COPY_ZERO_DST, YB_COPY_CONJ = 1, 2
YB_GEMM_CONJ_A, YB_GEMM_CONJ_B = 1, 2
YB_CHAIN_COPY, YB_CHAIN_GEMM = 0, 1

_lib = None


def load():
    """Load the shared library (once). Raises if it has not been built (python -m yastn_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `python -m yastn_b200.build` "
                              "(yastn_b200 has no CPU or torch fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if lib.yb_abi_version() != ABI_VERSION:
            raise ImportError("libyastn_b200.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(f"yastn_b200: {_lib.yb_last_error().decode()} (status {rc})")
