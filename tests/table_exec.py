"""CPU interpreters of the C-ABI tables (test infrastructure): they execute copy records / GEMM tables with
numpy exactly as include/yastn_b200.h specifies, so plans.py is verified against the oracle without a GPU."""
import numpy as np


def exec_copy(recs, rank, src, dst):
    for rec in recs:
        sb, db = int(rec[0]), int(rec[1])
        ext = rec[2:2 + rank]
        ss = rec[2 + rank:2 + 2 * rank]
        ds = rec[2 + 2 * rank:2 + 3 * rank]
        if np.prod(ext) == 0:
            continue
        idx = np.indices(tuple(int(e) for e in ext)).reshape(rank, -1)
        do = db + (idx * ds[:, None]).sum(axis=0)
        if sb == np.iinfo(np.int64).min:      # YB_COPY_SRC_ZERO: zero-fill record
            dst[do] = 0
            continue
        so = sb + (idx * ss[:, None]).sum(axis=0)
        dst[do] = src[so]
    return dst


def exec_gemm(problems, segments, A, B, C, conj_a=False, conj_b=False, scatter=None):
    for p, (M, N, offC, ldc, s0, s1) in enumerate(problems):
        acc = np.zeros((M, N), dtype=C.dtype)
        m = np.arange(M)[:, None]
        n = np.arange(N)[None, :]
        for (K, offA, sAm, sAk, offB, sBk, sBn) in segments[s0:s1]:
            k = np.arange(K)
            a = A[offA + m * sAm + k[None, :] * sAk]
            b = B[offB + k[:, None] * sBk + n * sBn]
            if conj_a:
                a = a.conj()
            if conj_b:
                b = b.conj()
            acc += a @ b
        if scatter is not None and scatter[0][p] >= 0:
            scat_index, row_ptr, row_cuts, col_ptr, col_cuts, dst_ptr, dst = scatter
            s = scat_index[p]
            rc = row_cuts[row_ptr[s]:row_ptr[s + 1]]
            cc = col_cuts[col_ptr[s]:col_ptr[s + 1]]
            d = dst[dst_ptr[s]:dst_ptr[s + 1]].reshape(len(rc) - 1, len(cc) - 1)
            for i in range(len(rc) - 1):
                for j in range(len(cc) - 1):
                    blk = acc[rc[i]:rc[i + 1], cc[j]:cc[j + 1]]
                    C[d[i, j]:d[i, j] + blk.size] = blk.reshape(-1)
        else:
            C[offC + m * ldc + n] = acc
    return C


def exec_ew(recs, traces, dst, srcs, aux):
    """Block-wise elementwise records (yb_ew_plan_create in include/yastn_b200.h)."""
    ABSENT = np.iinfo(np.int64).min
    for rec in recs:
        mode, d, n = int(rec[0]), int(rec[1]), int(rec[2])
        soff, neg, a, post, naxis, nfull = rec[3:7], int(rec[7]), int(rec[8]), max(int(rec[9]), 1), max(int(rec[10]), 1), int(rec[11])
        if n == 0:
            continue
        e = np.arange(n)
        if mode == 0:
            acc = np.zeros(n, dtype=dst.dtype)
            for k in range(4):
                if soff[k] != ABSENT:
                    x = srcs[k][int(soff[k]):int(soff[k]) + n]
                    acc = acc - x if (neg >> k) & 1 else acc + x
            dst[d:d + n] = acc
        elif mode == 1:
            j = (e // post) % naxis
            dst[d:d + n] = srcs[0][int(soff[0]):int(soff[0]) + n] * aux[a + j]
        elif mode in (2, 3):
            q, t = e % post, e // post
            j, p = t % naxis, t // naxis
            other = (p * nfull + aux[a + j]) * post + q
            if mode == 2:
                dst[d:d + n] = srcs[0][int(soff[0]) + other]
            else:
                dst[d + other] = srcs[0][int(soff[0]):int(soff[0]) + n]
        else:
            acc = np.zeros(n, dtype=dst.dtype)
            for row in traces[a:a + nfull]:
                base, D, ds, nd = int(row[0]), int(row[1]), int(row[2]), int(row[3])
                ext, st = row[4:4 + nd], row[10:10 + nd]
                rem, off = e.copy(), np.full(n, base, dtype=np.int64)
                for k in range(nd - 1, -1, -1):
                    off += (rem % ext[k]) * st[k]
                    rem //= ext[k]
                for i in range(D):
                    acc += srcs[0][off + i * ds]
            dst[d:d + n] = acc
    return dst
