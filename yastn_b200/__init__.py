"""yastn_b200 — B200-native (sm_100a) backend for YASTN's block-sparse symmetric contraction hot path.

The package holds exactly what that path needs:
  csrc/             hand-written CUDA kernels + the C ABI (include/yastn_b200.h)
  _lib.py           ctypes binding of the C ABI (fails loudly when the library is missing)
  plans.py          translation of YASTN's host metadata tuples into device plan tables, and the plan cache
  backend_b200.py   the drop-in backend functions (same names/signatures as yastn.backend.backend_torch)
  yastn_backend.py  the backend module object / activate() that plugs those functions into an unmodified YASTN
  decomp.py, cusolver_svdp.py   sector decompositions: batched Jacobi kernel, gesvdp, sector-parallel schedule
  chain.py          several tensordots in a row (Heff2, environment updates) recorded once, replayed by one library call
  sharding.py, peer.py   multi-GPU sector sharding and the NVLink peer-arena block exchange
  spmd.py           an unmodified YASTN program on several GPUs: sharded contractions and decompositions, replicated tensors
"""
__version__ = "0.3.0"
