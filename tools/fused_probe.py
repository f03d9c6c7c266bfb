#!/usr/bin/env python
"""Kernel time of dot + unmerge (two launches) vs dot_unmerge (scatter epilogue) on recorded structures (GPU box)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_io import bench_structs  # noqa: E402
from yastn_b200 import backend_b200 as bk  # noqa: E402


def ev_time(fn, reps=7):
    fn(); fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for name in sys.argv[1:] or ["U1xU1_D4096_P1", "U1_D1024_P1", "U1_D4096_P1", "U1_D16384_P1", "Z2_D512_P1", "U1_D64_P1"]:
    st = bench_structs()[name]["f2m"]
    md, um = st["dot"]["meta_dot"], st["unmerge"]["meta"]
    na = max(r[2][1] for r in md); nb = max(r[4][1] for r in md)
    for dt in (torch.float64, torch.complex128):
        A = torch.rand(na, dtype=torch.float64, device="cuda").to(dt); B = torch.rand(nb, dtype=torch.float64, device="cuda").to(dt)
        t_dot = ev_time(lambda: bk.dot(A, B, md, st["dot"]["Dsize"]))
        C = bk.dot(A, B, md, st["dot"]["Dsize"])
        t_unm = ev_time(lambda: bk.unmerge(C, um))
        t_fused = ev_time(lambda: bk.dot_unmerge(A, B, md, st["dot"]["Dsize"], um))
        print(json.dumps({"case": name, "dtype": str(dt).split(".")[1], "dot_ms": round(t_dot, 4), "unmerge_ms": round(t_unm, 4),
                          "two_launch_ms": round(t_dot + t_unm, 4), "fused_ms": round(t_fused, 4)}), flush=True)
