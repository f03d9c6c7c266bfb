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
