#!/usr/bin/env python
"""Panel products (an MPO block times a very long matrix) in isolation: C[M x N] = A[M x K] B[K x N], M, K <= 4, N ~ 10^6,
complex128, B and C row-major (the layout fuse_to_matrix gives Heff2's MPO applications).  Prints GB/s of algorithmic traffic."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from yastn_b200 import plans, _lib  # noqa: E402


def run(shapes, N, cplx=True, reps=5, check=True):
    dt = torch.complex128 if cplx else torch.float64
    isz = 16 if cplx else 8
    probs, segs, offA, offB, offC = [], [], 0, 0, 0
    for i, (M, K) in enumerate(shapes):
        probs.append([M, N, offC, N, i, i + 1])
        segs.append([K, offA, K, 1, offB, N, 1])
        offA += M * K
        offB += K * N
        offC += M * N
    A = torch.randn(offA, dtype=dt, device="cuda")
    B = torch.randn(offB, dtype=dt, device="cuda")
    C = torch.empty(offC, dtype=dt, device="cuda")
    plan = plans.GemmPlan(np.array(probs), np.array(segs), _lib.YB_C128 if cplx else _lib.YB_F64, 0)
    st = torch.cuda.current_stream().cuda_stream
    plan.run(A.data_ptr(), B.data_ptr(), C.data_ptr(), 0, st)
    torch.cuda.synchronize()
    if check:
        M, K = shapes[0]
        ref = A[:M * K].view(M, K) @ B[:K * N].view(K, N)
        err = float((C[:M * N].view(M, N) - ref).abs().max())
        assert err < 1e-10, err
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run(A.data_ptr(), B.data_ptr(), C.data_ptr(), 0, st)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = min(ms)
    gb = (offB + offC) * isz * 1e-9
    return {"shapes": shapes, "N": N, "cplx": cplx, "ms": round(t, 3), "GB": round(gb, 3), "GBps": round(gb / (t * 1e-3), 1), "info": plan.info()}


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    for shapes in ([(4, 4)] * 8, [(1, 1)] * 16, [(2, 2)] * 12, [(1, 1), (2, 2), (4, 4), (1, 2), (2, 1), (4, 2), (2, 4), (1, 4)] * 2, [(1, 1)] * 15 + [(4, 4)]):
        print(json.dumps(run(shapes, N)), flush=True)
    print(json.dumps(run([(4, 4)] * 8, N, cplx=False)), flush=True)
