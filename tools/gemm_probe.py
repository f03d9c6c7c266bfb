#!/usr/bin/env python
"""Time bk.dot on synthetic sector lists (measurement tool, GPU box only): isolates alignment, shape and count effects.

    python tools/gemm_probe.py [--dtype f64|c128]
Each line: one meta_dot of `count` identical sectors M x K x N (row-major operands packed back to back, offsets as given).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yastn_b200 import backend_b200 as bk  # noqa: E402
from yastn_b200 import plans  # noqa: E402


def meta_of(shapes, pad=0):
    recs, oa, ob, oc = [], pad, pad, 0
    for (M, K, N) in shapes:
        recs.append(((oc, oc + M * N), (M, N), (oa, oa + M * K), (M, K), (ob, ob + K * N), (K, N)))
        oa += M * K + pad; ob += K * N + pad; oc += M * N
    return tuple(recs), oa, ob, oc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--out", default="gpurun_out/gemm_probe.jsonl")
    args = ap.parse_args()
    cplx = args.dtype == "c128"
    tdt = torch.complex128 if cplx else torch.float64
    cases = [
        ("aligned 4096x1088x5328", [(4096, 1088, 5328)], 0),
        ("odd     4097x1089x5329", [(4097, 1089, 5329)], 0),
        ("aligned shapes, odd offset", [(4096, 1088, 5328)], 1),
        ("cube 8192 aligned", [(8192, 8192, 8192)], 0),
        ("cube 4096 aligned", [(4096, 4096, 4096)], 0),
        ("47 x 1024x1024x1024", [(1024, 1024, 1024)] * 47, 0),
        ("47 x 1023x1023x1023", [(1023, 1023, 1023)] * 47, 0),
        ("149 x 512x160x512", [(512, 160, 512)] * 149, 0),
        ("149 x 511x163x511", [(511, 163, 511)] * 149, 0),
        ("600 x 128x64x128", [(128, 64, 128)] * 600, 0),
        ("2000 x 32x32x32", [(32, 32, 32)] * 2000, 0),
        ("skinny 8x1000000x8", [(8, 1000000, 8)], 0),
        ("2 x 1024x320x1024 (Z2 D512)", [(1024, 320, 1024), (1024, 192, 1024)], 0),
    ]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = open(args.out, "a")
    for name, shapes, pad in cases:
        md, na, nb, nc = meta_of(shapes, pad)
        A = torch.rand(na, dtype=torch.float64, device="cuda").to(tdt)
        B = torch.rand(nb, dtype=torch.float64, device="cuda").to(tdt)
        for _ in range(3):
            C = bk.dot(A, B, md, nc)
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); C = bk.dot(A, B, md, nc); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        fl = sum((8 if cplx else 2) * M * K * N for M, K, N in shapes)
        problems, segments = plans.dot_tables(md)
        info = plans.GemmPlan(problems, segments, 1 if cplx else 0, 0).info()
        row = {"case": name, "dtype": args.dtype, "ms_min": min(ts), "ms_med": sorted(ts)[3], "TFLOPs": fl / min(ts) * 1e-9, "GFLOP": fl * 1e-9, **info}
        print(json.dumps(row), flush=True)
        out.write(json.dumps(row) + "\n")
        del A, B, C
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
