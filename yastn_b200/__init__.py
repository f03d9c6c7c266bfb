"""yastn_b200 — B200-native (sm_100a) backend for YASTN's block-sparse symmetric contraction hot path.

The package holds exactly what that path needs:
  csrc/            hand-written CUDA kernels + the C ABI (include/yastn_b200.h)
  _lib.py          ctypes binding of the C ABI (fails loudly when the library is missing)
  plans.py         translation of YASTN's host metadata tuples into device plan tables (cached by identity)
  backend_b200.py  the drop-in backend functions (same names/signatures as yastn.backend.backend_torch)
  tensordot.py     fused merge -> grouped GEMM -> scatter pipeline and multi-GPU sector sharding
"""
__version__ = "0.1.0"
