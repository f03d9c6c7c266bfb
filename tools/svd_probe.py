#!/usr/bin/env python
"""Per-sector SVD cost on the GPU by driver and size (measurement tool): what bounds the critical path of the sector pool.

    python tools/svd_probe.py [--dtype complex128] [--sizes 64 128 256 384 512 640 768]
"""
import argparse
import json
import time

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="complex128")
    ap.add_argument("--sizes", type=int, nargs="+", default=[64, 128, 256, 384, 512, 640, 768])
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dt = getattr(torch, args.dtype)
    torch.linalg.svd(torch.eye(4, dtype=dt, device="cuda"))
    rows = []
    for n in args.sizes:
        A = torch.randn(n, n, dtype=dt, device="cuda")
        row = {"n": n, "dtype": args.dtype}
        for drv in ("gesvd", "gesvdj", "gesvda"):
            try:
                torch.linalg.svd(A, full_matrices=False, driver=drv)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(3):
                    U, S, Vh = torch.linalg.svd(A, full_matrices=False, driver=drv)
                torch.cuda.synchronize()
                row[drv + "_ms"] = (time.perf_counter() - t0) / 3 * 1e3
                row[drv + "_err"] = float(torch.linalg.norm(U * S.to(dt) @ Vh - A) / torch.linalg.norm(A))
            except Exception as e:   # a driver may reject a dtype / shape
                row[drv + "_ms"] = None
                row[drv + "_err"] = str(e)[:80]
        try:     # cuSOLVER's polar-decomposition SVD, bound through ctypes (yastn_b200/cusolver_svdp.py)
            import os, sys
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            from yastn_b200 import cusolver_svdp
            cusolver_svdp.svd(A)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                U, S, Vh, err = cusolver_svdp.svd(A)
            torch.cuda.synchronize()
            row["gesvdp_ms"] = (time.perf_counter() - t0) / 3 * 1e3
            row["gesvdp_err"] = float(torch.linalg.norm(U * S.to(dt) @ Vh - A) / torch.linalg.norm(A))
            row["gesvdp_err_sigma"] = err
            S0 = torch.linalg.svdvals(A)
            row["gesvdp_dS_over_Smax"] = float((S - S0).abs().max() / S0.max())
            row["gesvdp_orthU"] = float(torch.linalg.norm(U.conj().t() @ U - torch.eye(U.shape[1], dtype=dt, device="cuda")))
            # graded spectrum over ten decades (a DMRG two-site tensor): relative accuracy of the small singular values
            Q1, _ = torch.linalg.qr(torch.randn(n, n, dtype=dt, device="cuda"))
            Q2, _ = torch.linalg.qr(torch.randn(n, n, dtype=dt, device="cuda"))
            sg = torch.logspace(0, -10, n, dtype=torch.float64, device="cuda")
            G = (Q1 * sg.to(dt)) @ Q2
            Ug, Sg, Vg, errg = cusolver_svdp.svd(G)
            Sr = torch.linalg.svdvals(G)
            row["graded_gesvdp_dS_over_Smax"] = float((Sg - sg).abs().max())
            row["graded_gesvd_dS_over_Smax"] = float((Sr - sg).abs().max())
            row["graded_gesvdp_rec"] = float(torch.linalg.norm(Ug * Sg.to(dt) @ Vg - G) / torch.linalg.norm(G))
            row["graded_gesvdp_kept_1e-8"] = int((Sg > 1e-8).sum()), int((sg > 1e-8).sum())
        except Exception as e:
            row["gesvdp_ms"] = None
            row["gesvdp_err"] = f"{type(e).__name__}: {str(e)[:120]}"
        Ac = A.cpu()
        t0 = time.perf_counter()
        torch.linalg.svd(Ac, full_matrices=False)
        row["cpu_gesdd_ms"] = (time.perf_counter() - t0) * 1e3
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "a") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
