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

from table_exec import exec_copy, exec_gemm, exec_ew
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


class CpuEwPlan:
    """Elementwise plans on CPU tensors: the run call gets raw pointers, so the extents of the views are derived from the
    records (sources and destination are addressed exactly as the kernel would)."""

    def __init__(self, recs, itemsize, device, traces=None):
        self.recs = np.ascontiguousarray(recs, dtype=np.int64).reshape(-1, 16)
        self.traces = np.zeros((0, 16), dtype=np.int64) if traces is None else np.ascontiguousarray(traces, dtype=np.int64).reshape(-1, 16)
        self.itemsize = itemsize

    def run(self, dst_ptr, src_ptrs, aux_ptr, stream):
        ABSENT = np.iinfo(np.int64).min
        big = 1 << 40      # views are created lazily large: only touched indices are accessed

        def ext_dst():
            m = 0
            for r in self.recs:
                mode, d, n = int(r[0]), int(r[1]), int(r[2])
                if mode == 3:
                    post, naxis, nfull = max(int(r[9]), 1), max(int(r[10]), 1), int(r[11])
                    pre = n // (post * naxis) if post * naxis else 0
                    m = max(m, d + pre * nfull * post)
                else:
                    m = max(m, d + n)
            return m
        dst = _view(dst_ptr, ext_dst(), self.itemsize)
        srcs = []
        for k in range(4):
            p = src_ptrs[k] if k < len(src_ptrs) else None
            if p is None:
                srcs.append(None)
                continue
            m = 0
            for r in self.recs:
                mode, n, off = int(r[0]), int(r[2]), int(r[3 + k])
                if off == ABSENT and mode == 0:
                    continue
                if mode in (0, 1, 3):
                    m = max(m, off + n)
                elif mode == 2:
                    post, naxis, nfull = max(int(r[9]), 1), max(int(r[10]), 1), int(r[11])
                    m = max(m, off + (n // (post * naxis)) * nfull * post)
                else:
                    for row in self.traces[int(r[8]):int(r[8]) + int(r[11])]:
                        nd = int(row[3])
                        m = max(m, int(row[0]) + (int(row[1]) - 1) * int(row[2]) + int(((row[4:4 + nd] - 1) * row[10:10 + nd]).sum()) + 1)
            srcs.append(_view(p, m, self.itemsize))
        aux = None
        if aux_ptr is not None:
            modes = set(int(r[0]) for r in self.recs)
            na = max((int(r[8]) + max(int(r[10]), 1) for r in self.recs), default=0)
            if modes & {2, 3}:
                buf = (ctypes.c_int64 * na).from_address(aux_ptr)
                aux = np.ctypeslib.as_array(buf)
            else:
                aux = _view(aux_ptr, na, self.itemsize)
        exec_ew(self.recs, self.traces, dst, srcs, aux)


class CpuChain:
    """Interpreter of a chain's step table (include/yastn_b200.h, yb_chain_run) over the CPU plans above."""

    def __init__(self, table, plan_list, nslots):
        self.table, self.plans, self.nslots = np.array(table, dtype=np.int64), list(plan_list), nslots

    def run(self, ptrs, stream):
        assert len(ptrs) == self.nslots
        for kind, flags, pi, sa, oa, sb, ob, sc, oc, n in self.table.tolist():
            plan = self.plans[pi]
            if kind == _lib.YB_CHAIN_COPY:
                plan.run(ptrs[sa] + oa, ptrs[sc] + oc, n, flags, None)
            else:
                plan.run(ptrs[sa] + oa, ptrs[sb] + ob, ptrs[sc] + oc, flags, None)


_saved = {}


def install():
    if _saved:
        return
    _saved.update(CopyPlan=plans.CopyPlan, GemmPlan=plans.GemmPlan, EwPlan=plans.EwPlan, on_device=bk._on_device, check=bk._check,
                  native=yastn_backend._native, on_gpu=yastn_backend._on_gpu)
    plans.CopyPlan, plans.GemmPlan, plans.EwPlan = CpuCopyPlan, CpuGemmPlan, CpuEwPlan
    bk._on_device = lambda dev, launch: launch(None)

    def check(t, name):
        if t.dtype not in bk._DTYPE_CODE:
            raise TypeError(f"yastn_b200.{name}: dtype {t.dtype} not supported (float64 / complex128 only)")
    bk._check = check
    yastn_backend._native = lambda *ts: all(t.dtype in (torch.float64, torch.complex128) for t in ts)
    yastn_backend._on_gpu = yastn_backend._native
    from yastn_b200 import chain
    _saved.update(chain_runner=chain._NativeChain, chain_usable=chain._usable)
    chain._NativeChain, chain._usable = CpuChain, (lambda d: True)
    chain.clear()
    bk.clear_plan_cache()


def uninstall():
    if not _saved:
        return
    plans.CopyPlan, plans.GemmPlan, plans.EwPlan = _saved["CopyPlan"], _saved["GemmPlan"], _saved["EwPlan"]
    bk._on_device, bk._check, yastn_backend._native = _saved["on_device"], _saved["check"], _saved["native"]
    yastn_backend._on_gpu = _saved["on_gpu"]
    from yastn_b200 import chain
    chain._NativeChain, chain._usable = _saved["chain_runner"], _saved["chain_usable"]
    chain.clear()
    bk.clear_plan_cache()
    _saved.clear()
