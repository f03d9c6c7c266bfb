"""CPU stand-in for the device side of the C ABI (TEST INFRASTRUCTURE ONLY — never imported by yastn_b200).

``install()`` swaps yastn_b200.plans.CopyPlan / GemmPlan for numpy interpreters of the same int64 tables
(tests/table_exec.py, which executes them exactly as include/yastn_b200.h specifies) and lifts the "CUDA tensors
only" checks, so that the *host logic* — plan construction from YASTN's metas, plan caching, dtype promotion, lazy
conj handling, autograd wiring, the yastn backend module — can be driven by the real YASTN on a box without a GPU.
The kernels themselves are only ever tested on the GPU (tests marked ``gpu``).
"""
import ctypes

import numpy as np
import torch

from table_exec import exec_copy, exec_gemm
from yastn_b200 import backend_b200 as bk
from yastn_b200 import plans, yastn_backend, _lib


def _view(ptr, n, itemsize):
    if n == 0:
        return np.zeros(0, dtype=np.float64 if itemsize == 8 else np.complex128)
    buf = (ctypes.c_double * (n * itemsize // 8)).from_address(ptr)
    arr = np.ctypeslib.as_array(buf)
    return arr if itemsize == 8 else arr.view(np.complex128)


class CpuCopyPlan:
    def __init__(self, recs, rank, itemsize, device, covered=None):
        self.recs = np.ascontiguousarray(recs, dtype=np.int64)
        self.rank, self.itemsize, self.device, self.covered = rank, itemsize, device, covered
        r = rank
        ext = self.recs[:, 2:2 + r]
        live = (ext > 0).all(axis=1)
        src_live = live & (self.recs[:, 0] != np.iinfo(np.int64).min)
        self.src_n = int((self.recs[src_live, 0] + ((ext[src_live] - 1) * self.recs[src_live, 2 + r:2 + 2 * r]).sum(axis=1)).max() + 1) if src_live.any() else 0
        self.dst_n = int((self.recs[live, 1] + ((ext[live] - 1) * self.recs[live, 2 + 2 * r:]).sum(axis=1)).max() + 1) if live.any() else 0

    def run(self, src_ptr, dst_ptr, dst_elems, flags, stream):
        dst = _view(dst_ptr, max(dst_elems, self.dst_n), self.itemsize)
        if flags & _lib.YB_COPY_ZERO_DST:
            dst[:dst_elems] = 0
        src = _view(src_ptr, self.src_n, self.itemsize)
        if flags & _lib.YB_COPY_CONJ:
            src = src.conj()
        exec_copy(self.recs, self.rank, src, dst)


class CpuGemmPlan:
    def __init__(self, problems, segments, dtype_code, device, scatter=None):
        self.scatter = None if scatter is None else [np.ascontiguousarray(t, dtype=np.int64) for t in scatter]
        self.problems = np.ascontiguousarray(problems, dtype=np.int64)
        self.segments = np.ascontiguousarray(segments, dtype=np.int64)
        self.itemsize = 16 if dtype_code == _lib.YB_C128 else 8
        na = nb = nc = 0
        for (M, N, offC, ldc, s0, s1) in self.problems:
            if M == 0 or N == 0:
                continue
            nc = max(nc, offC + (M - 1) * ldc + N)
            for (K, offA, sAm, sAk, offB, sBk, sBn) in self.segments[s0:s1]:
                if K == 0:
                    continue
                na = max(na, offA + (M - 1) * sAm + (K - 1) * sAk + 1)
                nb = max(nb, offB + (K - 1) * sBk + (N - 1) * sBn + 1)
        if self.scatter is not None:
            _, row_ptr, row_cuts, col_ptr, col_cuts, dst_ptr, dst = self.scatter
            for s in range(len(row_ptr) - 1):
                rc, cc = row_cuts[row_ptr[s]:row_ptr[s + 1]], col_cuts[col_ptr[s]:col_ptr[s + 1]]
                d = dst[dst_ptr[s]:dst_ptr[s + 1]].reshape(len(rc) - 1, len(cc) - 1)
                nc = max(nc, int((d + np.diff(rc)[:, None] * np.diff(cc)[None, :]).max()))
        self.na, self.nb, self.nc = int(na), int(nb), int(nc)

    def run(self, a_ptr, b_ptr, c_ptr, flags, stream):
        A, B, C = _view(a_ptr, self.na, self.itemsize), _view(b_ptr, self.nb, self.itemsize), _view(c_ptr, self.nc, self.itemsize)
        exec_gemm(self.problems, self.segments, A, B, C, bool(flags & _lib.YB_GEMM_CONJ_A), bool(flags & _lib.YB_GEMM_CONJ_B), self.scatter)


_saved = {}


def install():
    if _saved:
        return
    _saved.update(CopyPlan=plans.CopyPlan, GemmPlan=plans.GemmPlan, on_device=bk._on_device, check=bk._check, native=yastn_backend._native)
    plans.CopyPlan, plans.GemmPlan = CpuCopyPlan, CpuGemmPlan
    bk._on_device = lambda dev, launch: launch(None)

    def check(t, name):
        if t.dtype not in bk._DTYPE_CODE:
            raise TypeError(f"yastn_b200.{name}: dtype {t.dtype} not supported (float64 / complex128 only)")
    bk._check = check
    yastn_backend._native = lambda *ts: all(t.dtype in (torch.float64, torch.complex128) for t in ts)
    bk.clear_plan_cache()


def uninstall():
    if not _saved:
        return
    plans.CopyPlan, plans.GemmPlan = _saved["CopyPlan"], _saved["GemmPlan"]
    bk._on_device, bk._check, yastn_backend._native = _saved["on_device"], _saved["check"], _saved["native"]
    bk.clear_plan_cache()
    _saved.clear()
