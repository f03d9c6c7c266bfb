#!/usr/bin/env python
"""Tall sectors: cuSOLVER gesvdp on A directly against QR first (A = QR, SVD of the small R, U = Q U_R)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yastn_b200 import cusolver_svdp  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


def main():
    dt = torch.complex128
    torch.linalg.qr(torch.eye(4, dtype=dt, device="cuda"))
    for m, n in ((1304, 326), (2608, 652), (652, 163), (1304, 652), (978, 652), (652, 652), (5216, 652)):
        A = torch.randn(m, n, dtype=dt, device="cuda")

        def direct():
            return cusolver_svdp.svd(A)[:3]

        def qr_first():
            Q, R = torch.linalg.qr(A)
            U, S, Vh, _ = cusolver_svdp.svd(R)
            return Q @ U, S, Vh
        t_d, (U, S, Vh) = timed(direct)
        t_q, (U2, S2, Vh2) = timed(qr_first)
        t_g, _ = timed(lambda: torch.linalg.svd(A, full_matrices=False, driver="gesvd"), reps=1)
        row = {"m": m, "n": n, "gesvdp_ms": round(t_d, 2), "qr_first_ms": round(t_q, 2), "gesvd_ms": round(t_g, 2),
               "dS": float((S - S2).abs().max() / S.max()),
               "rec_qr_first": float(torch.linalg.norm(U2 * S2.to(dt) @ Vh2 - A) / torch.linalg.norm(A)),
               "orth_qr_first": float(torch.linalg.norm(U2.conj().t() @ U2 - torch.eye(n, dtype=dt, device="cuda")))}
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
