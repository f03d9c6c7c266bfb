"""Translate YASTN's host metadata tuples into the flat int64 tables of the C ABI, and cache device plans.

The table builders are pure numpy (no CUDA) so that they are unit-tested on CPU against the oracle;
the Plan classes own the device-side plan handles created through the C ABI.

Meta tuple layouts (produced by the reference's lru-cached ``_meta_*`` functions):
  transpose_and_merge : meta_new = ((tn, Dn, (lo, hi)), ...), meta_mrg = ((tn, (lo, hi), Do, Dslc, Drsh), ...)
                        yastn/tensor/_merging.py:84-89,137-187,304-377
  unmerge             : ((sln, Dn, slo, Do, sub_slc), ...)            yastn/tensor/_merging.py:488-549
  transpose           : ((sln, Dn, slo, Do), ...)                      yastn/tensor/_single.py:316-346
  dot                 : ((slc, Dc, sla, Da, slb, Db), ...)             yastn/tensor/_contractions.py:281-346
  transpose_dot_sum   : meta_dot = ((sl, (Dl, Dr), ((ia, ib), ...)), ...), Areshape = ((sl, Di, Dl, Dr), ...)
                        yastn/tensor/_contractions.py:349-450
"""
import ctypes
from collections import OrderedDict

import numpy as np

from . import _lib

I64 = np.int64


def _cstrides(shape):
    """Row-major strides (in elements) of every row of an [n, r] shape table."""
    shape = np.asarray(shape, dtype=I64)
    st = np.ones_like(shape)
    if shape.shape[1] > 1:
        st[:, :-1] = np.cumprod(shape[:, :0:-1], axis=1)[:, ::-1]
    return st


def _pack(src_base, dst_base, ext, sstr, dstr):
    n, r = ext.shape
    recs = np.empty((n, 2 + 3 * r), dtype=I64)
    recs[:, 0] = src_base
    recs[:, 1] = dst_base
    recs[:, 2:2 + r] = ext
    recs[:, 2 + r:2 + 2 * r] = sstr
    recs[:, 2 + 2 * r:] = dstr
    return np.ascontiguousarray(recs), r


def merge_records(order, meta_new, meta_mrg):
    """Copy records of transpose_and_merge. Returns (recs, rank, covered_elements)."""
    n = len(meta_mrg)
    r = len(order)
    if n == 0:
        return np.zeros((0, 2 + 3 * max(r, 1)), dtype=I64), max(r, 1), 0
    target = {tn: (Dn, sln[0]) for tn, Dn, sln in meta_new}
    g = len(meta_mrg[0][3])
    Do = np.array([m[2] for m in meta_mrg], dtype=I64).reshape(n, r)
    lo = np.array([[s[0] for s in m[3]] for m in meta_mrg], dtype=I64).reshape(n, g)
    Drsh = np.array([m[4] for m in meta_mrg], dtype=I64).reshape(n, g)
    Dn = np.array([target[m[0]][0] for m in meta_mrg], dtype=I64).reshape(n, g)
    sln0 = np.array([target[m[0]][1] for m in meta_mrg], dtype=I64)
    slo0 = np.array([m[1][0] for m in meta_mrg], dtype=I64)
    if r == 0:  # rank-0 tensor: one element per record
        one = np.ones((n, 1), dtype=I64)
        recs, rank = _pack(slo0, sln0, one, one, one)
        return recs, rank, n
    order = list(order)
    P = Do[:, order]                                  # permuted extents == destination index space
    sstr = _cstrides(Do)[:, order]
    pstr = _cstrides(P)                               # strides of the permuted linear index
    rstr = _cstrides(Drsh)                            # strides of the reshaped (g-dim) index
    nstr = _cstrides(Dn)                              # strides inside the destination block
    # every permuted dim lies inside exactly one reshape group: rstr[q] <= pstr[k] and the dim fits in the group
    fits = (rstr[:, None, :] <= pstr[:, :, None]) & (pstr[:, :, None] * P[:, :, None] <= rstr[:, None, :] * Drsh[:, None, :])
    q = np.argmax(fits, axis=2)                       # [n, r]
    rows = np.arange(n)[:, None]
    dstr = (pstr // rstr[rows, q]) * nstr[rows, q]
    dstr = np.where(P > 1, dstr, 0)
    bad = (P > 1) & ~np.take_along_axis(fits, q[:, :, None], axis=2)[:, :, 0]
    if bad.any():
        raise ValueError("transpose_and_merge: reshape groups do not align with permuted dims")
    dst_base = sln0 + (lo * nstr).sum(axis=1)
    recs, rank = _pack(slo0, dst_base, P, sstr, dstr)
    return recs, rank, int(P.prod(axis=1).sum())


def unmerge_records(meta):
    """Copy records of unmerge: an N-d box of a fused block -> one contiguous output block."""
    n = len(meta)
    if n == 0:
        return np.zeros((0, 5), dtype=I64), 1
    g = len(meta[0][3])
    if g == 0:
        one = np.ones((n, 1), dtype=I64)
        return _pack(np.array([m[2][0] for m in meta], dtype=I64), np.array([m[0][0] for m in meta], dtype=I64), one, one, one)
    Do = np.array([m[3] for m in meta], dtype=I64).reshape(n, g)
    lo = np.array([[s[0] for s in m[4]] for m in meta], dtype=I64).reshape(n, g)
    hi = np.array([[s[1] for s in m[4]] for m in meta], dtype=I64).reshape(n, g)
    ext = hi - lo
    sstr = _cstrides(Do)
    src_base = np.array([m[2][0] for m in meta], dtype=I64) + (lo * sstr).sum(axis=1)
    dst_base = np.array([m[0][0] for m in meta], dtype=I64)
    return _pack(src_base, dst_base, ext, sstr, _cstrides(ext))


def transpose_records(axes, meta):
    """Copy records of transpose: out[sln].view(Dn) = in[slo].view(Do).permute(axes)."""
    n = len(meta)
    r = len(axes)
    if n == 0 or r == 0:
        one = np.ones((n, 1), dtype=I64)
        return _pack(np.array([m[2][0] for m in meta], dtype=I64), np.array([m[0][0] for m in meta], dtype=I64), one, one, one)
    Do = np.array([m[3] for m in meta], dtype=I64).reshape(n, r)
    axes = list(axes)
    ext = Do[:, axes]
    sstr = _cstrides(Do)[:, axes]
    return _pack(np.array([m[2][0] for m in meta], dtype=I64), np.array([m[0][0] for m in meta], dtype=I64),
                 ext, sstr, _cstrides(ext))


def reverse_records(recs, rank):
    """Adjoint copy: swap the source and destination roles of every record."""
    out = recs.copy()
    out[:, 0], out[:, 1] = recs[:, 1], recs[:, 0]
    out[:, 2 + rank:2 + 2 * rank] = recs[:, 2 + 2 * rank:]
    out[:, 2 + 2 * rank:] = recs[:, 2 + rank:2 + 2 * rank]
    return np.ascontiguousarray(out)


def dot_tables(meta_dot):
    """GEMM tables of backend.dot: one problem and one segment per record."""
    n = len(meta_dot)
    problems = np.empty((n, 6), dtype=I64)
    segments = np.empty((n, 7), dtype=I64)
    for i, (slc, Dc, sla, Da, slb, Db) in enumerate(meta_dot):
        M, K = Da
        N = Db[1]
        problems[i] = (M, N, slc[0], N, i, i + 1)
        segments[i] = (K, sla[0], K, 1, slb[0], N, 1)
    return problems, segments


def dot_backward_tables(meta_dot):
    """GEMM tables of the adjoints  A_b += C_b @ B^H  and  B_b += A^H @ C_b  (grouped by target block).

    Returns (probA, segA, probB, segB).  For the A-gradient GEMM operand "A" is C_b and operand "B" is the
    forward B read as B^H (contraction index = n, contiguous); for the B-gradient operand "A" is the forward A
    read as A^H (m contiguous) and operand "B" is C_b.
    """
    by_a, by_b = OrderedDict(), OrderedDict()
    for rec in meta_dot:
        by_a.setdefault((rec[2], rec[3]), []).append(rec)
        by_b.setdefault((rec[4], rec[5]), []).append(rec)
    probA, segA = [], []
    for (sla, Da), lst in by_a.items():
        M, K = Da
        s0 = len(segA)
        for (slc, Dc, _, _, slb, Db) in lst:
            N = Db[1]
            # out[m, k] += sum_n Cb[m, n] * conj(B[k, n]):  A-op = Cb (ld N), B-op(kc=n, x=k) at slb + k*N + n
            segA.append((N, slc[0], N, 1, slb[0], 1, N))
        probA.append((M, K, sla[0], K, s0, len(segA)))
    probB, segB = [], []
    for (slb, Db), lst in by_b.items():
        K, N = Db
        s0 = len(segB)
        for (slc, Dc, sla, Da, _, _) in lst:
            M = Da[0]
            # out[k, n] += sum_m conj(A[m, k]) * Cb[m, n]:  A-op(x=k, kc=m) at sla + m*K + k, B-op = Cb (ld N)
            segB.append((M, sla[0], 1, K, slc[0], N, 1))
        probB.append((K, N, slb[0], N, s0, len(segB)))
    f = lambda x, w: np.array(x, dtype=I64).reshape(len(x), w)
    return f(probA, 6), f(segA, 7), f(probB, 6), f(segB, 7)


def unmerge_scatter_tables(meta_dot, meta_unmerge):
    """Tables of the fused unmerge epilogue (include/yastn_b200.h, yb_gemm_plan_create_scatter).

    ``meta_unmerge`` (yastn/tensor/_merging.py:528-549) lists, for every merged C block ``slo`` of shape ``Do``,
    the rectangles ``((r0, r1), (c0, c1))`` that become the output blocks at ``sln``.  The rectangles of one merged
    block form a complete grid (row cuts x col cuts); the GEMM epilogue then writes every element straight to its
    output block.  Returns (scat_index[nprob], row_ptr, row_cuts, col_ptr, col_cuts, dst_ptr, dst).
    """
    by_src = OrderedDict()
    for sln, Dn, slo, Do, sub in meta_unmerge:
        by_src.setdefault(slo, []).append((sln, Do, sub))
    scat_index = np.full(len(meta_dot), -1, dtype=I64)
    row_ptr, col_ptr, dst_ptr = [0], [0], [0]
    row_cuts, col_cuts, dst = [], [], []
    used = 0
    for p, (slc, Dc, _, _, _, _) in enumerate(meta_dot):
        recs = by_src.get(slc)
        if recs is None:
            if Dc[0] * Dc[1] == 0:
                continue
            raise ValueError("unmerge meta does not cover every block produced by dot")
        used += 1
        rows = sorted({sub[0] for _, _, sub in recs})
        cols = sorted({sub[1] for _, _, sub in recs})
        if len(rows) * len(cols) != len(recs):
            raise ValueError("unmerge rectangles of a block do not form a grid")
        ri = {rc: i for i, rc in enumerate(rows)}
        ci = {cc: j for j, cc in enumerate(cols)}
        block = np.full((len(rows), len(cols)), -1, dtype=I64)
        for sln, Do, sub in recs:
            if tuple(Do) != tuple(Dc):
                raise ValueError("unmerge source shape differs from the dot block shape")
            block[ri[sub[0]], ci[sub[1]]] = sln[0]
        rc = [rows[0][0]] + [r[1] for r in rows]
        cc = [cols[0][0]] + [c[1] for c in cols]
        ok = (block >= 0).all() and rc[0] == 0 and cc[0] == 0 and rc[-1] == Dc[0] and cc[-1] == Dc[1] \
            and all(a[1] == b[0] for a, b in zip(rows, rows[1:])) and all(a[1] == b[0] for a, b in zip(cols, cols[1:]))
        if not ok:
            raise ValueError("unmerge rectangles do not tile the merged block")
        scat_index[p] = len(row_ptr) - 1
        row_cuts += rc
        col_cuts += cc
        dst += block.reshape(-1).tolist()
        row_ptr.append(len(row_cuts))
        col_ptr.append(len(col_cuts))
        dst_ptr.append(len(dst))
    if used != len(by_src):
        raise ValueError("unmerge meta references blocks that dot does not produce")
    f = lambda x: np.array(x, dtype=I64)
    return scat_index, f(row_ptr), f(row_cuts), f(col_ptr), f(col_cuts), f(dst_ptr), f(dst)


def _matrix_view(Di, order, Dl, Dr):
    """Strides (row, col) of block.reshape(Di).permute(order).reshape(Dl, Dr) if it is a strided matrix, else None."""
    ext = [Di[k] for k in order]
    st_full = [1] * len(Di)
    for k in range(len(Di) - 2, -1, -1):
        st_full[k] = st_full[k + 1] * Di[k + 1]
    st = [st_full[k] for k in order]
    # split point: leading dims whose product is Dl (unit dims at the boundary may fall on either side)
    split, prod = 0, 1
    while prod != Dl:
        prod *= ext[split]
        split += 1
    out = []
    for grp_e, grp_s in ((ext[:split], st[:split]), (ext[split:], st[split:])):
        dims = [(e, s) for e, s in zip(grp_e, grp_s) if e > 1]
        for (e0, s0), (e1, s1) in zip(dims, dims[1:]):
            if s0 != s1 * e1:
                return None
        out.append(dims[-1][1] if dims else 1)
    return out[0], out[1]


def tds_tables(meta_dot, Areshape, Breshape, Aorder, Border):
    """GEMM tables of backend.transpose_dot_sum (no_fusion): one problem per result block, one segment per pair.

    Returns (problems, segments, need_pack_a, need_pack_b): when an operand's permuted blocks are not plain
    strided matrices it must first be packed by a transpose copy (same slices, permuted layout), after which
    its blocks are row-major (Dl x Dr).
    """
    va = [_matrix_view(Di, Aorder, Dl, Dr) for (_, Di, Dl, Dr) in Areshape]
    vb = [_matrix_view(Di, Border, Dl, Dr) for (_, Di, Dl, Dr) in Breshape]
    pack_a = any(v is None for v in va)
    pack_b = any(v is None for v in vb)
    # a plan has one layout per operand: mixed unit-stride choices also force packing
    def consistent(views, shapes):
        kc = xc = True
        for v, (_, _, Dl, Dr) in zip(views, shapes):
            kc &= (v[1] == 1 or Dr <= 1)
            xc &= (v[0] == 1 or Dl <= 1)
        return kc or xc
    if not pack_a and not consistent(va, Areshape):
        pack_a = True
    if not pack_b and not consistent(vb, Breshape):
        pack_b = True
    problems, segments = [], []
    for (sl, (Dl, Dr), pairs) in meta_dot:
        s0 = len(segments)
        for ia, ib in pairs:
            sla, _, Ml, K = Areshape[ia]
            slb, _, Kb, N = Breshape[ib]
            sam, sak = (K, 1) if pack_a else va[ia]
            sbk, sbn = (N, 1) if pack_b else vb[ib]
            segments.append((K, sla[0], sam, sak, slb[0], sbk, sbn))
        problems.append((Dl, Dr, sl[0], Dr, s0, len(segments)))
    f = lambda x, w: np.array(x, dtype=I64).reshape(len(x), w)
    return f(problems, 6), f(segments, 7), pack_a, pack_b


def pack_records(reshape, order):
    """Copy records packing every block of a no_fusion operand into its permuted (Dl x Dr) row-major layout."""
    meta = tuple((sl, tuple(Di[k] for k in order), sl, Di) for (sl, Di, _, _) in reshape)
    return transpose_records(order, meta)


# -------------------------------------------------------------------------------------------------
# device plans
# -------------------------------------------------------------------------------------------------

def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class CopyPlan:
    """Device plan of one block-copy launch (owns the C handle)."""

    def __init__(self, recs, rank, itemsize, device, covered=None):
        lib = _lib.load()
        self._lib = lib
        self.handle = ctypes.c_void_p()
        recs = np.ascontiguousarray(recs, dtype=I64)
        self.nrec = recs.shape[0]
        _lib.check(lib.yb_copy_plan_create(_ptr(recs), self.nrec, rank, itemsize, device, ctypes.byref(self.handle)))
        self.covered = covered
        self.itemsize = itemsize
        self.device = device

    def info(self):
        out = (ctypes.c_int64 * 4)()
        _lib.check(self._lib.yb_copy_plan_info(self.handle, out))
        return {"items": out[0], "elements": out[1], "records": out[2], "tiled_records": out[3]}

    def run(self, src_ptr, dst_ptr, dst_elems, flags, stream):
        rc = self._lib.yb_copy_run(self.handle, src_ptr, dst_ptr, dst_elems, flags, stream)
        if rc:
            _lib.check(rc)

    def __del__(self):
        try:
            if self.handle:
                self._lib.yb_copy_plan_destroy(self.handle)
        except Exception:
            pass


class GemmPlan:
    """Device plan of one grouped-GEMM launch (owns the C handle)."""

    def __init__(self, problems, segments, dtype_code, device, scatter=None):
        lib = _lib.load()
        self._lib = lib
        self.handle = ctypes.c_void_p()
        problems = np.ascontiguousarray(problems, dtype=I64)
        segments = np.ascontiguousarray(segments, dtype=I64)
        if scatter is None:
            _lib.check(lib.yb_gemm_plan_create(_ptr(problems), problems.shape[0], _ptr(segments), segments.shape[0],
                                               dtype_code, device, ctypes.byref(self.handle)))
        else:
            tabs = [np.ascontiguousarray(t, dtype=I64) for t in scatter]
            _lib.check(lib.yb_gemm_plan_create_scatter(_ptr(problems), problems.shape[0], _ptr(segments), segments.shape[0],
                                                       _ptr(tabs[0]), tabs[1].shape[0] - 1, *[_ptr(t) for t in tabs[1:]],
                                                       dtype_code, device, ctypes.byref(self.handle)))
        self.device = device

    def info(self):
        out = (ctypes.c_int64 * 6)()
        _lib.check(self._lib.yb_gemm_plan_info(self.handle, out))
        return {"tiles": out[0], "macs": out[1], "big_tiles": out[2], "small_tiles": out[3], "grid": out[4], "split_ctas": out[5]}

    def run(self, a_ptr, b_ptr, c_ptr, flags, stream):
        rc = self._lib.yb_gemm_run(self.handle, a_ptr, b_ptr, c_ptr, flags, stream)
        if rc:
            _lib.check(rc)

    def __del__(self):
        try:
            if self.handle:
                self._lib.yb_gemm_plan_destroy(self.handle)
        except Exception:
            pass


class PlanCache:
    """LRU cache of device plans keyed by the *identity* of YASTN's cached meta objects.

    YASTN's ``_meta_*`` functions are lru_cached and hand back the same tuple object on every hit
    (yastn/tensor/_merging.py:137, _contractions.py:281), so ``id(meta)`` is a stable O(1) key; the cache
    keeps a strong reference to the meta so the id cannot be recycled while the entry lives.
    """

    def __init__(self, maxsize=4096):
        self.maxsize = maxsize
        self._d = OrderedDict()
        self.hits = 0
        self.misses = 0

    def get(self, key, anchor, build):
        ent = self._d.get(key)
        if ent is not None and ent[0] is anchor:
            self._d.move_to_end(key)
            self.hits += 1
            return ent[1]
        self.misses += 1
        val = build()
        self._d[key] = (anchor, val)
        if len(self._d) > self.maxsize:
            self._d.popitem(last=False)
        return val

    def clear(self):
        self._d.clear()
