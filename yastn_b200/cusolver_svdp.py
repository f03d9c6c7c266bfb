"""cuSOLVER's polar-decomposition SVD (cusolverDnXgesvdp) through ctypes, for the large sectors of the block SVD.

The reference factorises every charge sector with ``torch.linalg.svd(driver='gesvd')`` (yastn/backend/linalg/
torch_svd_gesdd.py:17): QR iteration on the bidiagonal form, mostly BLAS-2 work that leaves a B200 idle — 85 ms for one
640 x 640 complex128 sector, the critical path of a D = 4096 DMRG sweep.  ``gesvdp`` (QDWH polar decomposition followed by
a Hermitian eigensolver) does the same factorisation with matrix-matrix products; torch exposes only gesvd / gesvdj /
gesvda, so the routine is bound here directly from the cuSOLVER library torch itself loads.  This is a library path like
the rest of ``decomp.py``, opt-in per sector size (``YASTN_B200_SVDP_MIN``).
"""
import ctypes
import glob
import os
import threading

import torch

_lock = threading.Lock()
_state = {"lib": None, "tls": threading.local()}

CUDA_R_64F, CUDA_C_64F = 1, 5
CUSOLVER_EIG_MODE_VECTOR = 1


def _load():
    with _lock:
        if _state["lib"] is None:
            cands = []
            for root in (os.path.dirname(torch.__file__), os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "cusolver", "lib")):
                cands += sorted(glob.glob(os.path.join(root, "**", "libcusolver.so*"), recursive=True))
            cands += ["libcusolver.so.11", "libcusolver.so"]
            err = None
            for c in cands:
                try:
                    lib = ctypes.CDLL(c)
                    if hasattr(lib, "cusolverDnXgesvdp"):
                        _state["lib"] = lib
                        break
                except OSError as e:
                    err = e
            if _state["lib"] is None:
                raise ImportError(f"cusolverDnXgesvdp not found ({err})")
            lib = _state["lib"]
            vp, i64, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t
            lib.cusolverDnCreate.argtypes = [ctypes.POINTER(vp)]
            lib.cusolverDnSetStream.argtypes = [vp, vp]
            lib.cusolverDnCreateParams.argtypes = [ctypes.POINTER(vp)]
            lib.cusolverDnXgesvdp_bufferSize.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, i64, i64, ctypes.c_int, vp, i64, ctypes.c_int, vp,
                                                         ctypes.c_int, vp, i64, ctypes.c_int, vp, i64, ctypes.c_int,
                                                         ctypes.POINTER(sz), ctypes.POINTER(sz)]
            lib.cusolverDnXgesvdp.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, i64, i64, ctypes.c_int, vp, i64, ctypes.c_int, vp,
                                              ctypes.c_int, vp, i64, ctypes.c_int, vp, i64, ctypes.c_int, vp, sz, vp, sz, vp,
                                              ctypes.POINTER(ctypes.c_double)]
    return _state["lib"]


def _handle(device):
    """One cuSOLVER handle (+ params object) per host thread and device."""
    lib = _load()
    tls = _state["tls"]
    key = f"h{device.index}"
    h = getattr(tls, key, None)
    if h is None:
        handle, params = ctypes.c_void_p(), ctypes.c_void_p()
        with torch.cuda.device(device):
            if lib.cusolverDnCreate(ctypes.byref(handle)) != 0 or lib.cusolverDnCreateParams(ctypes.byref(params)) != 0:
                raise RuntimeError("cusolverDnCreate failed")
        h = (handle, params)
        setattr(tls, key, h)
    return h


def available():
    try:
        _load()
        return True
    except ImportError:
        return False


def svd(A):
    """Thin SVD of a 2-D CUDA tensor (float64 / complex128): returns (U [m,k], S [k], Vh [k,n], err_sigma) with views that the
    caller copies into its own storage.  Runs on torch's current stream; no host synchronisation apart from gesvdp's own."""
    m, n = A.shape
    if m < n:      # gesvdp wants m >= n: factorise A^H = V S U^H
        U, S, Vh, err = svd(A.conj().t())
        return Vh.conj().t(), S, U.conj().t(), err
    lib = _load()
    dev = A.device
    handle, params = _handle(dev)
    cplx = A.is_complex()
    dt = CUDA_C_64F if cplx else CUDA_R_64F
    k = n
    with torch.cuda.device(dev):
        lib.cusolverDnSetStream(handle, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        # gesvdp DESTROYS its input (it is overwritten by the polar factor): always a private column-major copy (lda = m).
        # (A.t().contiguous() aliases the caller's storage when A is the transposed view of a contiguous matrix.)
        a = torch.empty((n, m), dtype=A.dtype, device=dev)
        a.copy_(A.t())                                 # copy_ also resolves a lazy conj bit
        U = torch.empty((k, m), dtype=A.dtype, device=dev)      # column-major m x k
        V = torch.empty((k, n), dtype=A.dtype, device=dev)      # column-major n x k
        S = torch.empty(k, dtype=torch.float64, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        wd, wh = ctypes.c_size_t(), ctypes.c_size_t()
        rc = lib.cusolverDnXgesvdp_bufferSize(handle, params, CUSOLVER_EIG_MODE_VECTOR, 1, m, n, dt, a.data_ptr(), m, CUDA_R_64F, S.data_ptr(),
                                              dt, U.data_ptr(), m, dt, V.data_ptr(), n, dt, ctypes.byref(wd), ctypes.byref(wh))
        if rc != 0:
            raise RuntimeError(f"cusolverDnXgesvdp_bufferSize failed with status {rc}")
        work = torch.empty(max(wd.value, 1), dtype=torch.uint8, device=dev)
        hwork = (ctypes.c_char * max(wh.value, 1))()
        err = ctypes.c_double()
        rc = lib.cusolverDnXgesvdp(handle, params, CUSOLVER_EIG_MODE_VECTOR, 1, m, n, dt, a.data_ptr(), m, CUDA_R_64F, S.data_ptr(),
                                   dt, U.data_ptr(), m, dt, V.data_ptr(), n, dt, work.data_ptr(), wd.value, ctypes.cast(hwork, ctypes.c_void_p),
                                   wh.value, info.data_ptr(), ctypes.byref(err))
        if rc != 0:
            raise RuntimeError(f"cusolverDnXgesvdp failed with status {rc}")
    return U.t(), S, V.conj(), err.value
