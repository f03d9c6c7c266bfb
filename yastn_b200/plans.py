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
from ._flatten import flatten as _flatten_bytes      # CPython helper built by yastn_b200.build (csrc/yb_flatten.c)

I64 = np.int64


def _table(meta, n):
    """Nested meta tuples of n equally-shaped records -> [n, width] int64 array (depth-first flattening in C)."""
    flat = np.frombuffer(_flatten_bytes(meta), dtype=I64)
    if n == 0:
        return flat.reshape(0, 0)
    if flat.size % n:
        raise ValueError("meta records do not have a common width")
    return flat.reshape(n, flat.size // n)


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


SRC_ZERO = np.iinfo(np.int64).min      # YB_COPY_SRC_ZERO: a record without source, its destination box is filled with zeros
_ZERO_ROWS_MAX = 1 << 16               # above this many zero rows a memset of the whole destination is cheaper than the table


def _zero_records(grp, lo, hi, Dn_new, sln_new, rank):
    """Boxes of the merged blocks that no source block covers, as zero-fill records (src_base = SRC_ZERO).

    The reference allocates the merged tensor with zeros and scatters the source blocks into it
    (yastn/backend/_backend_torch_backwards.py:349-363); charge sectors that are absent from the tensor leave holes (20-30 % of
    the merges of a CTMRG / Hubbard run).  The source rectangles of one merged block lie on a grid (row segments x
    column segments; one segment list per fused leg in the N-d case): the holes are the grid cells no record occupies.
    Returns (records [m, 2 + 3 * rank] or None when a plain memset is the better plan, number of zero elements)."""
    g = lo.shape[1]
    vol = np.bincount(grp, weights=(hi - lo).prod(axis=1).astype(np.float64), minlength=Dn_new.shape[0])
    holes = np.nonzero(vol < Dn_new.prod(axis=1))[0]
    if holes.size == 0:
        return np.zeros((0, 2 + 3 * rank), dtype=I64), 0
    if g > rank:
        return None, 0
    order = np.argsort(grp, kind="stable")
    bounds = np.searchsorted(grp[order], np.arange(Dn_new.shape[0] + 1))
    out, rows, zeros = [], 0, 0
    for t in holes:
        idx = order[bounds[t]:bounds[t + 1]]
        Dn = Dn_new[t]
        cuts = [np.unique(np.concatenate(([0, Dn[d]], lo[idx, d], hi[idx, d]))) for d in range(g)]
        occ = np.zeros([c.size - 1 for c in cuts], dtype=bool)
        for i in idx:
            occ[tuple(slice(np.searchsorted(cuts[d], lo[i, d]), np.searchsorted(cuts[d], hi[i, d])) for d in range(g))] = True
        nstr = _cstrides(Dn[None, :])[0]
        free = np.argwhere(~occ)
        k = 0
        while k < free.shape[0]:                        # argwhere is row-major: neighbours along the last dim are consecutive
            c0 = free[k]
            e = k + 1
            while e < free.shape[0] and (free[e, :-1] == c0[:-1]).all() and free[e, -1] == free[e - 1, -1] + 1:
                e += 1
            b_lo = np.array([cuts[d][c0[d]] for d in range(g)], dtype=I64)
            b_hi = np.array([cuts[d][c0[d] + 1] for d in range(g)], dtype=I64)
            b_hi[-1] = cuts[-1][free[e - 1, -1] + 1]
            ext = b_hi - b_lo
            rec = np.zeros(2 + 3 * rank, dtype=I64)
            rec[0], rec[1] = SRC_ZERO, sln_new[t] + (b_lo * nstr).sum()
            rec[2:2 + rank] = 1
            rec[2:2 + g] = ext
            rec[2 + 2 * rank:2 + 2 * rank + g] = nstr
            out.append(rec)
            rows += int(ext[:-1].prod()) if g > 1 else 1
            zeros += int(ext.prod())
            k = e
    if rows > _ZERO_ROWS_MAX or zeros != int((Dn_new[holes].prod(axis=1) - vol[holes]).sum()):
        return None, 0                                   # too fragmented (or overlapping source boxes): memset instead
    return np.array(out, dtype=I64).reshape(len(out), 2 + 3 * rank), zeros


def merge_records_np(order, meta_new, meta_mrg, zero_records=True):
    """numpy statement of :func:`merge_records` (the C pass is checked against it on CPU).  Copy records of transpose_and_merge.  Returns (recs, rank, covered): ``covered`` = destination elements the records
    write, including the zero-fill records of uncovered cells (``zero_records``); a caller whose ``covered`` is below
    ``Dsize`` must clear the destination first."""
    n = len(meta_mrg)
    r = len(order)
    if n == 0:
        return np.zeros((0, 2 + 3 * max(r, 1)), dtype=I64), max(r, 1), 0
    g = len(meta_mrg[0][3])
    T = len(meta_mrg[0][0])
    mrg = _table(meta_mrg, n)                          # [tn(T), slo(2), Do(r), Dslc(2g), Drsh(g)]
    new = _table(meta_new, len(meta_new))              # [tn(T), Dn(g), sln(2)]
    if mrg.shape[1] != T + 2 + r + 3 * g or new.shape[1] != T + g + 2:
        raise ValueError("transpose_and_merge: unexpected meta layout")
    # meta_mrg is grouped by target charge in the order of meta_new (the reference relies on it: itertools.groupby in
    # _backend_torch_backwards.py:358); fall back to a dictionary when a caller hands over another order
    tn_m = mrg[:, :T]
    grp = np.zeros(n, dtype=I64)
    if n > 1 and T > 0:
        grp[1:] = np.cumsum((tn_m[1:] != tn_m[:-1]).any(axis=1))
    if grp[-1] >= new.shape[0] or not np.array_equal(new[grp, :T], tn_m):
        index = {tn: i for i, (tn, _, _) in enumerate(meta_new)}
        grp = np.array([index[m[0]] for m in meta_mrg], dtype=I64)
    Do = mrg[:, T + 2:T + 2 + r]
    lo = mrg[:, T + 2 + r:T + 2 + r + 2 * g:2]
    Drsh = mrg[:, T + 2 + r + 2 * g:]
    Dn = new[grp, T:T + g]
    sln0 = new[grp, T + g]
    slo0 = mrg[:, T]
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
    covered = int(P.prod(axis=1).sum())
    if zero_records:
        hi = mrg[:, T + 2 + r + 1:T + 2 + r + 2 * g:2]
        zrec, nzero = _zero_records(grp, lo, hi, new[:, T:T + g], new[:, T + g], rank)
        if zrec is not None and zrec.shape[0]:
            recs = np.ascontiguousarray(np.vstack([recs, zrec]))
            covered += nzero
    return recs, rank, covered


def unmerge_records(meta):
    """Copy records of unmerge: an N-d box of a fused block -> one contiguous output block."""
    n = len(meta)
    if n == 0:
        return np.zeros((0, 5), dtype=I64), 1
    g, gn = len(meta[0][3]), len(meta[0][1])
    tab = _table(meta, n)                              # [sln(2), Dn(gn), slo(2), Do(g), sub(2g)]
    if tab.shape[1] != 4 + gn + 3 * g:
        raise ValueError("unmerge: unexpected meta layout")
    sln0, slo0 = tab[:, 0], tab[:, 2 + gn]
    if g == 0:
        one = np.ones((n, 1), dtype=I64)
        return _pack(slo0, sln0, one, one, one)
    Do = tab[:, 4 + gn:4 + gn + g]
    lo = tab[:, 4 + gn + g::2]
    hi = tab[:, 5 + gn + g::2]
    ext = hi - lo
    sstr = _cstrides(Do)
    return _pack(slo0 + (lo * sstr).sum(axis=1), sln0, ext, sstr, _cstrides(ext))


def transpose_records(axes, meta):
    """Copy records of transpose: out[sln].view(Dn) = in[slo].view(Do).permute(axes)."""
    n = len(meta)
    r = len(axes)
    if n == 0:
        return np.zeros((0, 5), dtype=I64), 1
    tab = _table(meta, n)                              # [sln(2), Dn(r), slo(2), Do(r)]
    if tab.shape[1] != 4 + 2 * r:
        raise ValueError("transpose: unexpected meta layout")
    sln0, slo0 = tab[:, 0], tab[:, 2 + r]
    if r == 0:
        one = np.ones((n, 1), dtype=I64)
        return _pack(slo0, sln0, one, one, one)
    Do = tab[:, 4 + r:]
    axes = list(axes)
    ext = Do[:, axes]
    return _pack(slo0, sln0, ext, _cstrides(Do)[:, axes], _cstrides(ext))


def reverse_records(recs, rank):
    """Adjoint copy: swap the source and destination roles of every record (zero-fill records have no adjoint)."""
    recs = recs[recs[:, 0] != SRC_ZERO]
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
    if n == 0:
        return problems, segments
    tab = _table(meta_dot, n)                          # [slc(2), Dc(2), sla(2), Da(2), slb(2), Db(2)]
    if tab.shape[1] != 12:
        raise ValueError("dot: unexpected meta layout")
    M, K, N = tab[:, 6], tab[:, 7], tab[:, 11]
    idx = np.arange(n, dtype=I64)
    problems[:, 0], problems[:, 1], problems[:, 2], problems[:, 3], problems[:, 4], problems[:, 5] = M, N, tab[:, 0], N, idx, idx + 1
    segments[:, 0], segments[:, 1], segments[:, 2], segments[:, 3] = K, tab[:, 4], K, 1
    segments[:, 4], segments[:, 5], segments[:, 6] = tab[:, 8], N, 1
    return problems, segments


def vdot_tables(meta):
    """GEMM tables of backend.vdot (yastn/backend/backend_torch.py:537-546): one 1 x 1 problem per pair of common blocks,
    ``C[ii] = sum_k A[sla + k] * B[slb + k]``; the contraction index is contiguous in both operands."""
    n = len(meta)
    problems = np.empty((n, 6), dtype=I64)
    segments = np.empty((n, 7), dtype=I64)
    if n == 0:
        return problems, segments
    tab = _table(meta, n)                              # [sla(2), slb(2)]
    if tab.shape[1] != 4:
        raise ValueError("vdot: unexpected meta layout")
    K = tab[:, 1] - tab[:, 0]
    if (K != tab[:, 3] - tab[:, 2]).any():
        raise ValueError("vdot: blocks of different size")
    idx = np.arange(n, dtype=I64)
    problems[:, 0], problems[:, 1], problems[:, 2], problems[:, 3], problems[:, 4], problems[:, 5] = 1, 1, idx, 1, idx, idx + 1
    # M = N = 1: the row stride of A and the column stride of B are never used to address a valid element; 0 keeps any padded
    # row / column the loaders might touch inside the block (a joined slice can exceed the 2^24-element stride limit)
    segments[:, 0], segments[:, 1], segments[:, 2], segments[:, 3] = K, tab[:, 0], 0, 1
    segments[:, 4], segments[:, 5], segments[:, 6] = tab[:, 2], 1, 0
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


def unmerge_scatter_tables_np(meta_dot, meta_unmerge, dst_shift=None):
    """numpy statement of :func:`unmerge_scatter_tables` (the C pass is checked against it on CPU).  Tables of the fused unmerge epilogue (include/yastn_b200.h, yb_gemm_plan_create_scatter).

    ``meta_unmerge`` (yastn/tensor/_merging.py:528-549) lists, for every merged C block ``slo`` of shape ``Do``,
    the rectangles ``((r0, r1), (c0, c1))`` that become the output blocks at ``sln``.  The rectangles of one merged
    block form a complete grid (row cuts x col cuts); the GEMM epilogue then writes every element straight to its
    output block.  ``dst_shift`` (one int64 per record of ``meta_unmerge``) is added to the destination offset of that
    output block: with peer arenas it redirects the block into another rank's buffer (peer.PeerArena.shift).
    Returns (scat_index[nprob], row_ptr, row_cuts, col_ptr, col_cuts, dst_ptr, dst).
    """
    nprob, n = len(meta_dot), len(meta_unmerge)
    if n == 0:
        if any(rec[1][0] * rec[1][1] for rec in meta_dot):
            raise ValueError("unmerge meta does not cover every block produced by dot")
        z = np.zeros(1, dtype=I64)
        return np.full(nprob, -1, dtype=I64), z, np.zeros(0, dtype=I64), z, np.zeros(0, dtype=I64), z, np.zeros(0, dtype=I64)
    gn = len(meta_unmerge[0][1])
    if len(meta_unmerge[0][3]) != 2:
        raise ValueError("fused unmerge needs matrix-shaped source blocks")
    um = _table(meta_unmerge, n)                       # [sln(2), Dn(gn), slo(2), Do(2), r0, r1, c0, c1]
    md = _table(meta_dot, nprob)
    sln0, slo0 = um[:, 0], um[:, 2 + gn]
    DoM, DoN = um[:, 4 + gn], um[:, 5 + gn]
    r0, r1, c0, c1 = (um[:, 6 + gn + k] for k in range(4))
    src, grp = np.unique(slo0, return_inverse=True)    # merged blocks that are unmerged, and the group of every record
    ng = src.size
    # rank of every record's row / column cut inside its group (scalar keys group * big + cut keep np.unique 1-D)
    big = max(int(um[:, 6 + gn:].max()) + 1, 1)
    rk, ri = np.unique(grp * big + r0, return_inverse=True)       # sorted (group, r0) pairs
    ck, ci = np.unique(grp * big + c0, return_inverse=True)
    rkey = np.stack([rk // big, rk % big], axis=1)
    ckey = np.stack([ck // big, ck % big], axis=1)
    nrs = np.bincount(rkey[:, 0], minlength=ng)
    ncs = np.bincount(ckey[:, 0], minlength=ng)
    rstart = np.concatenate(([0], np.cumsum(nrs)))[:-1]
    cstart = np.concatenate(([0], np.cumsum(ncs)))[:-1]
    ri = ri - rstart[grp]
    ci = ci - cstart[grp]
    if not np.array_equal(np.bincount(grp, minlength=ng), nrs * ncs):
        raise ValueError("unmerge rectangles of a block do not form a grid")
    dst_ptr = np.concatenate(([0], np.cumsum(nrs * ncs)))
    dst = np.full(int(dst_ptr[-1]), -1, dtype=I64)
    placed = np.zeros(dst.size, dtype=bool)
    placed[dst_ptr[grp] + ri * ncs[grp] + ci] = True
    dst[dst_ptr[grp] + ri * ncs[grp] + ci] = sln0 if dst_shift is None else sln0 + np.asarray(dst_shift, dtype=I64)
    # cuts: the sorted starts of a group followed by the extent of the merged block; rectangles must tile it
    M_g = np.zeros(ng, dtype=I64); N_g = np.zeros(ng, dtype=I64)
    M_g[grp], N_g[grp] = DoM, DoN
    row_ptr = np.concatenate(([0], np.cumsum(nrs + 1)))
    col_ptr = np.concatenate(([0], np.cumsum(ncs + 1)))
    row_cuts = np.empty(int(row_ptr[-1]), dtype=I64)
    col_cuts = np.empty(int(col_ptr[-1]), dtype=I64)
    row_cuts[row_ptr[rkey[:, 0]] + (np.arange(rkey.shape[0]) - rstart[rkey[:, 0]])] = rkey[:, 1]
    col_cuts[col_ptr[ckey[:, 0]] + (np.arange(ckey.shape[0]) - cstart[ckey[:, 0]])] = ckey[:, 1]
    row_cuts[row_ptr[1:] - 1] = M_g
    col_cuts[col_ptr[1:] - 1] = N_g
    rnext = row_cuts[row_ptr[grp] + ri + 1]
    cnext = col_cuts[col_ptr[grp] + ci + 1]
    ok = placed.all() and (rnext == r1).all() and (cnext == c1).all() \
        and (row_cuts[row_ptr[:-1]] == 0).all() and (col_cuts[col_ptr[:-1]] == 0).all()
    if not ok:
        raise ValueError("unmerge rectangles do not tile the merged block")
    # problems -> groups
    slc0, Mp, Np = md[:, 0], md[:, 2], md[:, 3]
    pos = np.searchsorted(src, slc0)
    pos_c = np.minimum(pos, ng - 1)
    hit = src[pos_c] == slc0
    empty = (Mp * Np) == 0
    if (~hit & ~empty).any():
        raise ValueError("unmerge meta does not cover every block produced by dot")
    scat_index = np.where(hit & ~empty, pos_c, -1).astype(I64)
    live = scat_index >= 0
    if np.unique(scat_index[live]).size != ng:
        raise ValueError("unmerge meta references blocks that dot does not produce")
    if not (np.array_equal(M_g[scat_index[live]], Mp[live]) and np.array_equal(N_g[scat_index[live]], Np[live])):
        raise ValueError("unmerge source shape differs from the dot block shape")
    return scat_index, row_ptr.astype(I64), row_cuts, col_ptr.astype(I64), col_cuts, dst_ptr.astype(I64), dst


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


def bs_to_tds(a_t, a_slices, nout_a, nin_a, b_t, b_slices, nout_b, nin_b, c_t, c_slices):
    """Raw block tables of ``kernel_tensordot_bs`` -> the metas of ``transpose_dot_sum``.

    The single-call boundary (yastn/backend/backend_torch_cpp.py:173-188, called from yastn/tensor/_contractions.py:199-242)
    hands over only the block lists: flattened charges ``*_t`` and ``_slc`` records (offset range, shape ``D``) of the blocks
    of a and b that take part, the outgoing / contracted native axes, and the block list of the result.  Which blocks
    multiply is decided here: a pair (ia, ib) contributes when its contracted-leg charges agree, and lands in the result
    block whose charges are (outgoing charges of ia, outgoing charges of ib).  All joins are vectorised (sort-based), there
    is no Python loop over block pairs.  Returns (meta_dot, Areshape, Breshape) in the layout of
    yastn/tensor/_contractions.py:349-450.
    """
    na, nb, nc = len(a_t), len(b_t), len(c_t)
    ra, rb = len(nout_a) + len(nin_a), len(nout_b) + len(nin_b)
    if na == 0 or nb == 0 or nc == 0:
        return (), (), ()
    Da = np.array([s.D for s in a_slices], dtype=I64).reshape(na, ra)
    Db = np.array([s.D for s in b_slices], dtype=I64).reshape(nb, rb)
    w = len(a_t[0]) // ra                              # charges per leg (NSYM; 1 for the synthetic dense block)
    ta = np.array(a_t, dtype=I64).reshape(na, ra, w)
    tb = np.array(b_t, dtype=I64).reshape(nb, rb, w)
    nout_a, nin_a, nout_b, nin_b = list(nout_a), list(nin_a), list(nout_b), list(nin_b)
    ka = ta[:, nin_a, :].reshape(na, -1)
    kb = tb[:, nin_b, :].reshape(nb, -1)
    if ka.shape[1]:
        _, inv = np.unique(np.vstack([ka, kb]), axis=0, return_inverse=True)
        inv = inv.reshape(-1)
    else:                                              # outer product: every pair matches
        inv = np.zeros(na + nb, dtype=I64)
    ida, idb = inv[:na], inv[na:]
    nid = int(inv.max()) + 1
    oa, ob = np.argsort(ida, kind="stable"), np.argsort(idb, kind="stable")
    ca, cb = np.bincount(ida, minlength=nid), np.bincount(idb, minlength=nid)
    sa = np.concatenate(([0], np.cumsum(ca)))[:-1]
    sb = np.concatenate(([0], np.cumsum(cb)))[:-1]
    cnt = ca * cb
    first = np.concatenate(([0], np.cumsum(cnt)))[:-1]
    pid = np.repeat(np.arange(nid), cnt)
    pos = np.arange(int(cnt.sum())) - first[pid]
    ia = oa[sa[pid] + pos // np.maximum(cb[pid], 1)]
    ib = ob[sb[pid] + pos % np.maximum(cb[pid], 1)]
    if (Da[ia][:, nin_a] != Db[ib][:, nin_b]).any():
        raise ValueError("Bond dimensions do not match.")
    # result block of every pair
    tc = np.hstack([ta[ia][:, nout_a, :].reshape(ia.size, -1), tb[ib][:, nout_b, :].reshape(ib.size, -1)])
    ctab = np.array(c_t, dtype=I64).reshape(nc, -1)
    if ctab.shape[1]:
        _, inv2 = np.unique(np.vstack([ctab, tc]), axis=0, return_inverse=True)
        inv2 = inv2.reshape(-1)
        lut = np.full(int(inv2.max()) + 1, -1, dtype=I64)
        lut[inv2[:nc]] = np.arange(nc)
        ic = lut[inv2[nc:]]
    else:
        ic = np.zeros(ia.size, dtype=I64)
    if (ic < 0).any():
        raise ValueError("kernel_tensordot_bs: a block pair has no result block in c_struct_t")
    order = np.argsort(ic, kind="stable")
    ia, ib, ic = ia[order], ib[order], ic[order]
    Dal, Dar = Da[:, nout_a].prod(axis=1), Da[:, nin_a].prod(axis=1)
    Dbl, Dbr = Db[:, nin_b].prod(axis=1), Db[:, nout_b].prod(axis=1)
    Areshape = tuple((s.slcs[0], tuple(s.D), int(l), int(r)) for s, l, r in zip(a_slices, Dal, Dar))
    Breshape = tuple((s.slcs[0], tuple(s.D), int(l), int(r)) for s, l, r in zip(b_slices, Dbl, Dbr))
    bounds = np.searchsorted(ic, np.arange(nc + 1))
    pairs = list(zip(ia.tolist(), ib.tolist()))
    meta_dot = []
    for k in range(nc):
        lo, hi = int(bounds[k]), int(bounds[k + 1])
        sl = c_slices[k].slcs[0]
        if hi > lo:
            Dl, Dr = int(Dal[ia[lo]]), int(Dbr[ib[lo]])
        else:                                          # no contributing pair: the block is all zeros
            Dl, Dr = sl[1] - sl[0], 1
        if Dl * Dr != sl[1] - sl[0]:
            raise ValueError("kernel_tensordot_bs: result block size differs from the product of the outgoing dimensions")
        meta_dot.append((sl, (Dl, Dr), tuple(pairs[lo:hi])))
    return tuple(meta_dot), Areshape, Breshape


# -------------------------------------------------------------------------------------------------
# block-wise elementwise tables (include/yastn_b200.h, yb_ew_plan_create)
# -------------------------------------------------------------------------------------------------

EW_ABSENT = np.iinfo(np.int64).min
EW_LINCOMB, EW_DIAG, EW_GATHER, EW_SCATTER, EW_TRACE = range(5)
EW_SOURCES = 4


def _ew_rec(mode, dst, n, src=(), neg=0, aux=0, post=1, naxis=1, nfull=0):
    rec = [mode, dst, n] + list(src) + [EW_ABSENT] * (EW_SOURCES - len(src)) + [neg, aux, post, naxis, nfull, 0, 0, 0, 0]
    return rec


def _ew_array(recs):
    return np.array(recs, dtype=I64).reshape(len(recs), 16)


def add_tables_np(metas, signs=None):
    """Python statement of :func:`add_tables` (the C pass is checked against it on CPU).  Rounds of LINCOMB records of backend.add / sub (yastn/backend/backend_torch.py:518-534): ``new[sl_c] (+/-)= data_k[sl_a]``
    for every (sl_c, sl_a) of ``metas[k]``; output ranges no operand writes stay zero (``newdata = torch.zeros``).

    One launch adds up to four operands; with more, later rounds read the running sum as their first source.  Returns a list
    of (records, operand indices) — operand index -1 is the output itself."""
    n_ops = len(metas)
    signs = signs or (1,) * n_ops
    cuts = sorted({x for meta in metas for (sl_c, _) in meta for x in sl_c})
    # elementary output intervals and who writes them
    starts = np.array(cuts[:-1], dtype=I64) if len(cuts) > 1 else np.zeros(0, dtype=I64)
    writers = [[] for _ in range(max(len(cuts) - 1, 0))]
    for k, meta in enumerate(metas):
        for (c0, c1), (a0, a1) in meta:
            if c1 <= c0:
                continue
            i = int(np.searchsorted(starts, c0))
            while i < len(starts) and starts[i] < c1:
                writers[i].append((k, a0 + int(starts[i]) - c0))
                i += 1
    rounds = []
    ops_left = list(range(n_ops))
    first = True
    while ops_left:
        take = ops_left[:EW_SOURCES if first else EW_SOURCES - 1]
        ops_left = ops_left[len(take):]
        slots = ([] if first else [-1]) + take
        recs = []
        for i, w in enumerate(writers):
            lo, hi = cuts[i], cuts[i + 1]
            src = [EW_ABSENT] * len(slots)
            neg = 0
            hit = False
            for (k, off) in w:
                if k in take:
                    j = slots.index(k)
                    if src[j] != EW_ABSENT:
                        raise ValueError("add: an operand writes an output element twice")
                    src[j] = off
                    hit = True
                    if signs[k] < 0:
                        neg |= 1 << j
            if not first:
                if not hit:
                    continue                      # nothing new for this interval: the running sum stays
                src[0] = lo
            elif not w:
                continue                          # gap between blocks: belongs to no output block
            recs.append(_ew_rec(EW_LINCOMB, lo, hi - lo, src, neg))
        rounds.append((_ew_array(recs), slots))
        first = False
    return rounds


def negate_tables(slices, size):
    """LINCOMB records of negate_blocks (yastn/backend/_backend_torch_backwards.py:229-248): a copy of the data with the
    sign of the listed slices flipped."""
    recs, pos = [], 0
    for lo, hi in sorted(slices):
        if lo > pos:
            recs.append(_ew_rec(EW_LINCOMB, pos, lo - pos, [pos], 0))
        if hi > lo:
            recs.append(_ew_rec(EW_LINCOMB, lo, hi - lo, [lo], 1))
        pos = max(pos, hi)
    if size > pos:
        recs.append(_ew_rec(EW_LINCOMB, pos, size - pos, [pos], 0))
    return _ew_array(recs)


def dot_diag_tables(meta, axis, a_ndim):
    """DIAG records of backend.dot_diag (yastn/backend/backend_torch.py:557-564): ``new[sln].reshape(Db) = A[sla] (along axis) * B[slb].reshape(Db)``."""
    recs = []
    for sln, slb, Db, sla in meta:
        Db = (Db,) if isinstance(Db, int) else tuple(Db)
        ax = axis if a_ndim > 0 else 0
        post = int(np.prod(Db[ax + 1:], dtype=I64))
        recs.append(_ew_rec(EW_DIAG, sln[0], sln[1] - sln[0], [slb[0]], 0, sla[0], post, Db[ax], 0))
    return _ew_array(recs)


def mask_tables(meta, axis, ndim, scatter):
    """GATHER / SCATTER records of apply_mask / embed_mask (yastn/backend/_backend_torch_backwards.py:251-310).

    meta = ((sln, Dn, sla, Da, tm), ...): apply_mask takes ``A[sla].view(Da)[..., mask[tm], ...]`` into ``C[sln].view(Dn)``
    (scatter=False, the iteration space is the output block Dn, Da has the full extent); embed_mask puts ``A[sla].view(Da)``
    into ``C[sln].view(Dn)[..., mask[tm], ...]`` (scatter=True, the iteration space is the input block Da, Dn is full).
    Returns (records, tm order): the index vectors ``mask[tm]`` are concatenated in that order."""
    order, aux_of, pos = [], {}, 0
    recs = []
    for sln, Dn, sla, Da, tm in meta:
        small, full = (Da, Dn) if scatter else (Dn, Da)
        small = (small,) if isinstance(small, int) else tuple(small)
        full = (full,) if isinstance(full, int) else tuple(full)
        ax = axis if ndim > 0 else 0
        nsel = small[ax] if small else 1
        if tm not in aux_of:
            aux_of[tm] = pos
            order.append(tm)
            pos += nsel
        post = int(np.prod(small[ax + 1:], dtype=I64))
        n = int(np.prod(small, dtype=I64))
        recs.append(_ew_rec(EW_SCATTER if scatter else EW_GATHER, sln[0], n, [sla[0]], 0, aux_of[tm], post, nsel, full[ax] if full else 1))
    return _ew_array(recs), tuple(order)


def block_copy_tables(meta):
    """LINCOMB records of apply_mask / embed_mask whose mask is ``slice(None)`` for every charge (yastn.flip_charges,
    yastn/tensor/_single.py:220-225): every record moves a whole block, ``new[sln] = A[sla]``."""
    recs = []
    for sln, Dn, sla, Da, _ in meta:
        n = sln[1] - sln[0]
        if n != sla[1] - sla[0]:
            raise ValueError("mask slice(None): source and destination blocks differ in size")
        if n > 0:
            recs.append(_ew_rec(EW_LINCOMB, sln[0], n, [sla[0]], 0))
    return _ew_array(recs)


def trace_tables(order, meta):
    """TRACE records of backend.trace (yastn/backend/backend_torch.py:268-275):
    ``new[sln] += sum_i data[slo].reshape(Do).permute(order).reshape(D, D, rest)[i, i, :]``."""
    recs, traces = [], []
    order = list(order)
    for sln, lst in meta:
        first = len(traces)
        for slo, Do, Drsh in lst:
            st = _cstrides(np.array([Do], dtype=I64))[0]
            P = [Do[k] for k in order]
            ps = [int(st[k]) for k in order]
            # leading permuted dims form the two traced groups (products Drsh[0] == Drsh[1]); the rest is the output index
            split0, prod = 0, 1
            while prod != Drsh[0]:
                prod *= P[split0]
                split0 += 1
            split1, prod = split0, 1
            while prod != Drsh[1]:
                prod *= P[split1]
                split1 += 1
            g0 = [(e, s) for e, s in zip(P[:split0], ps[:split0]) if e > 1]
            g1 = [(e, s) for e, s in zip(P[split0:split1], ps[split0:split1]) if e > 1]
            if [e for e, _ in g0] != [e for e, _ in g1]:
                raise ValueError("trace: traced leg groups of different shape")
            rest = [(e, s) for e, s in zip(P[split1:], ps[split1:]) if e > 1]
            if len(rest) > 6:
                raise ValueError("trace: more than 6 remaining legs")
            # the diagonal of a multi-leg group is not a single stride: one trace row per index of all but the last traced leg
            outer = g0[:-1]
            outer1 = g1[:-1]
            D, ds = (g0[-1][0], g0[-1][1] + g1[-1][1]) if g0 else (1, 0)
            for idx in np.ndindex(*[e for e, _ in outer]) if outer else [()]:
                base = slo[0] + sum(i * (s0 + s1) for i, (_, s0), (_, s1) in zip(idx, outer, outer1))
                row = [base, D, ds, len(rest)] + [e for e, _ in rest] + [1] * (6 - len(rest)) + [s for _, s in rest] + [0] * (6 - len(rest))
                traces.append(row)
        recs.append(_ew_rec(EW_TRACE, sln[0], sln[1] - sln[0], [0], 0, first, 1, 1, len(traces) - first))
    tr = np.array(traces, dtype=I64).reshape(len(traces), 16)
    return _ew_array(recs), tr


# -------------------------------------------------------------------------------------------------
# meta pass in C (csrc/yb_tables.cu): the three builders a launch-bound sweep calls for every new block structure
# -------------------------------------------------------------------------------------------------

def _value_check(lib, rc):
    if rc:
        raise ValueError(lib.yb_last_error().decode())


def _fetch(lib):
    out = np.empty(lib.yb_tables_result_size(), dtype=I64)
    _lib.check(lib.yb_tables_result_fetch(_ptr(out), out.size))
    return out


def merge_records(order, meta_new, meta_mrg, zero_records=True):
    """Copy records of transpose_and_merge (one per source block, plus zero-fill records for the cells of the merged blocks
    that no source block covers).  Returns (recs, rank, covered): ``covered`` = destination elements the records write; a
    caller whose ``covered`` is below ``Dsize`` must clear the destination first."""
    n, r = len(meta_mrg), len(order)
    if n == 0:
        return np.zeros((0, 2 + 3 * max(r, 1)), dtype=I64), max(r, 1), 0
    lib = _lib.load()
    g, T = len(meta_mrg[0][3]), len(meta_mrg[0][0])
    mrg = _table(meta_mrg, n)
    new = _table(meta_new, len(meta_new))
    od = np.array(order, dtype=I64).reshape(-1)
    _value_check(lib, lib.yb_tables_merge(_ptr(mrg), n, mrg.shape[1], _ptr(new), new.shape[0], new.shape[1] if new.size else T + g + 2,
                                          _ptr(od), r, g, T, None, 1 if zero_records else 0))
    out = _fetch(lib)
    if out[0] == 1:      # records not grouped in the order of meta_new (the reference relies on it, _backend_torch_backwards.py:358)
        index = {tn: i for i, (tn, _, _) in enumerate(meta_new)}
        grp = np.array([index[m[0]] for m in meta_mrg], dtype=I64)
        _value_check(lib, lib.yb_tables_merge(_ptr(mrg), n, mrg.shape[1], _ptr(new), new.shape[0], new.shape[1], _ptr(od), r, g, T, _ptr(grp),
                                              1 if zero_records else 0))
        out = _fetch(lib)
    rank, covered, nrec = int(out[1]), int(out[2]), int(out[3])
    return out[4:].reshape(nrec, 2 + 3 * rank), rank, covered


def unmerge_scatter_tables(meta_dot, meta_unmerge, dst_shift=None):
    """Tables of the fused unmerge epilogue (include/yastn_b200.h, yb_gemm_plan_create_scatter): the rectangles
    ``((r0, r1), (c0, c1))`` of every merged C block (yastn/tensor/_merging.py:528-549) form a grid of row cuts x column cuts
    and the GEMM epilogue writes every element straight to its output block.  ``dst_shift`` (one int64 per record of
    ``meta_unmerge``) redirects a block into another rank's buffer (peer.PeerArena.shift).
    Returns (scat_index[nprob], row_ptr, row_cuts, col_ptr, col_cuts, dst_ptr, dst)."""
    nprob, n = len(meta_dot), len(meta_unmerge)
    if n == 0:
        return unmerge_scatter_tables_np(meta_dot, meta_unmerge, dst_shift)
    if len(meta_unmerge[0][3]) != 2:
        raise ValueError("fused unmerge needs matrix-shaped source blocks")
    lib = _lib.load()
    gn = len(meta_unmerge[0][1])
    um = _table(meta_unmerge, n)
    md = _table(meta_dot, nprob)
    if um.shape[1] != 10 + gn or (nprob and md.shape[1] != 12):
        raise ValueError("fused unmerge: unexpected meta layout")
    sh = None if dst_shift is None else np.ascontiguousarray(dst_shift, dtype=I64)
    _value_check(lib, lib.yb_tables_scatter(_ptr(um), n, gn, _ptr(md), nprob, None if sh is None else _ptr(sh)))
    out = _fetch(lib)
    nprob_, ng, nrow, ncol, ndst = (int(v) for v in out[:5])
    cuts = np.cumsum([5, nprob_, ng + 1, nrow, ng + 1, ncol, ng + 1, ndst])
    return tuple(out[cuts[k]:cuts[k + 1]] for k in range(7))


def add_tables(metas, signs=None):
    """Rounds of LINCOMB records of backend.add / sub (yastn/backend/backend_torch.py:518-534): ``new[sl_c] (+/-)= data_k[sl_a]``
    for every (sl_c, sl_a) of ``metas[k]``; output ranges no operand writes stay zero.  One launch adds up to four operands;
    with more, later rounds read the running sum as their first source.  Returns a list of (records, operand indices) —
    operand index -1 is the output itself."""
    n_ops = len(metas)
    if n_ops == 0:
        return []
    lib = _lib.load()
    rows = [(k, c[0], c[1], a[0]) for k, meta in enumerate(metas) for (c, a) in meta]
    ops = np.array(rows, dtype=I64).reshape(len(rows), 4)
    sg = np.array(signs or (1,) * n_ops, dtype=I64)
    _value_check(lib, lib.yb_tables_add(_ptr(ops), ops.shape[0], n_ops, _ptr(sg)))
    out = _fetch(lib)
    rounds, pos = [], 1
    for _ in range(int(out[0])):
        m, ns = int(out[pos]), int(out[pos + 1])
        slots = [int(v) for v in out[pos + 2:pos + 2 + ns]]
        rounds.append((out[pos + 6:pos + 6 + 16 * m].reshape(m, 16), slots))
        pos += 6 + 16 * m
    return rounds


# -------------------------------------------------------------------------------------------------
# device plans
# -------------------------------------------------------------------------------------------------

def _ptr(a):
    return a.ctypes.data          # plain address: every pointer argument is declared c_void_p (data_as costs 5 us per call)


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
        out = (ctypes.c_int64 * 5)()
        _lib.check(self._lib.yb_copy_plan_info(self.handle, out))
        return {"items": out[0], "elements": out[1], "records": out[2], "tiled_records": out[3], "runs": out[4]}

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


class EwPlan:
    """Device plan of one block-wise elementwise launch (owns the C handle)."""

    def __init__(self, recs, itemsize, device, traces=None):
        lib = _lib.load()
        self._lib = lib
        self.handle = ctypes.c_void_p()
        recs = np.ascontiguousarray(recs, dtype=I64).reshape(-1, 16)
        traces = np.zeros((0, 16), dtype=I64) if traces is None else np.ascontiguousarray(traces, dtype=I64).reshape(-1, 16)
        _lib.check(lib.yb_ew_plan_create(_ptr(recs), recs.shape[0], _ptr(traces), traces.shape[0], itemsize, device, ctypes.byref(self.handle)))
        self.itemsize, self.device = itemsize, device

    def info(self):
        out = (ctypes.c_int64 * 2)()
        _lib.check(self._lib.yb_ew_plan_info(self.handle, out))
        return {"pieces": out[0], "elements": out[1]}

    def run(self, dst_ptr, src_ptrs, aux_ptr, stream):
        p = list(src_ptrs) + [None] * (EW_SOURCES - len(src_ptrs))
        rc = self._lib.yb_ew_run(self.handle, dst_ptr, p[0], p[1], p[2], p[3], aux_ptr, stream)
        if rc:
            _lib.check(rc)

    def __del__(self):
        try:
            if self.handle:
                self._lib.yb_ew_plan_destroy(self.handle)
        except Exception:
            pass


class SvdPlan:
    """Device plan of one batched small-sector SVD launch (owns the C handle)."""

    def __init__(self, recs, itemsize, device):
        lib = _lib.load()
        self._lib = lib
        self.handle = ctypes.c_void_p()
        recs = np.ascontiguousarray(recs, dtype=I64).reshape(-1, 6)
        self.nrec = recs.shape[0]
        _lib.check(lib.yb_svd_plan_create(_ptr(recs), self.nrec, itemsize, device, ctypes.byref(self.handle)))

    def run(self, a_ptr, u_ptr, s_ptr, vh_ptr, status_ptr, max_sweeps, vectors, stream):
        rc = self._lib.yb_svd_run(self.handle, a_ptr, u_ptr, s_ptr, vh_ptr, status_ptr, max_sweeps, 1 if vectors else 0, stream)
        if rc:
            _lib.check(rc)

    def __del__(self):
        try:
            if self.handle:
                self._lib.yb_svd_plan_destroy(self.handle)
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
        out = (ctypes.c_int64 * 10)()
        _lib.check(self._lib.yb_gemm_plan_info(self.handle, out))
        return {"tiles": out[0], "macs": out[1], "big_tiles": out[2], "small_tiles": out[3], "grid": out[4], "split_ctas": out[5],
                "skinny_warps": out[6], "skinny_runs": out[7], "panel_units": out[8]}

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
    """LRU cache of device plans.

    Most entries are keyed by the *identity* of YASTN's cached meta objects: the ``_meta_*`` functions behind merge, unmerge,
    dot and vdot are lru_cached and hand back the same tuple object on every hit (yastn/tensor/_merging.py:137,
    _contractions.py:281), so ``id(meta)`` is a stable O(1) key; the entry keeps a strong reference to the meta (``anchor``)
    so the id cannot be recycled while it lives.  Metas that YASTN rebuilds on every call (``consume_transpose``,
    yastn/tensor/_single.py:343, is not cached) are keyed by content instead: the key holds the meta tuple itself and
    ``anchor`` is None.
    """

    def __init__(self, maxsize=4096):
        self.maxsize = maxsize
        self._d = OrderedDict()
        self.hits = 0
        self.misses = 0

    def get(self, key, anchor, build):
        ent = self._d.get(key)
        if ent is not None and (anchor is None or ent[0] is anchor):
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
