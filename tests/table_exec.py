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
        so = sb + (idx * ss[:, None]).sum(axis=0)
        do = db + (idx * ds[:, None]).sum(axis=0)
        dst[do] = src[so]
    return dst


def exec_gemm(problems, segments, A, B, C, conj_a=False, conj_b=False):
    for (M, N, offC, ldc, s0, s1) in problems:
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
        C[offC + m * ldc + n] = acc
    return C
