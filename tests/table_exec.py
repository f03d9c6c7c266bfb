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
