#!/usr/bin/env python
"""Per-kernel timing table of the hot path (measurement tool, GPU box only).

For every benchmark structure (tests/golden/structs_bench.json.gz) and dtype, time each stage of the
fuse_to_matrix pipeline on its own with CUDA events (L2 flushed between repetitions) and print
  merge_a / merge_b / unmerge : algorithmic GB/s (itemsize * (elements read + elements written)) vs HBM peak
  dot                         : algorithmic TFLOP/s (2 or 8 * M*K*N) vs the FP64 DMMA pipe peak
Writes one JSON document to the path given by --out.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_io import bench_structs  # noqa: E402
from yastn_b200 import backend_b200 as bk  # noqa: E402


def timed(fn, flush, reps):
    best = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1))
    best.sort()
    return best[0], best[len(best) // 2], out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/kernel_table.json")
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--names", nargs="*", default=None)
    ap.add_argument("--dtypes", nargs="*", default=["f64", "c128"])
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    fp64 = 37.1
    st = bench_structs()
    names = args.names or [n for n in st if not n.endswith("D64_P1") and not n.endswith("D64_P2") and not n.endswith("D64_P3")]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for name in names:
        case = st[name]
        for dt in args.dtypes:
            cplx = dt == "c128"
            isz = 16 if cplx else 8
            tdt = torch.complex128 if cplx else torch.float64
            s = case["f2m"]
            need = (case["a"]["size"] + case["b"]["size"] + 3 * s["dot"]["Dsize"]) * isz
            if need > 120e9:
                continue
            A = torch.rand(case["a"]["size"], dtype=torch.float64, device="cuda") * 2 - 1
            B = torch.rand(case["b"]["size"], dtype=torch.float64, device="cuda") * 2 - 1
            if cplx:
                A = torch.complex(A, A.flip(0)); B = torch.complex(B, B.flip(0))
            row = {"case": name, "dtype": dt}
            Am, Bm = A, B
            for key, X in (("merge_a", A), ("merge_b", B)):
                m = s[key]
                if m is None:
                    continue
                f = lambda: bk.transpose_and_merge(X, m["order"], m["meta_new"], m["meta_mrg"], m["Dsize"])
                f(); f()
                t, tmed, out = timed(f, flush, args.reps)
                src = sum(lo_hi[1] - lo_hi[0] for (_, lo_hi, _, _, _) in m["meta_mrg"])
                gb = isz * (src + m["Dsize"]) * 1e-9
                row[key] = {"ms": t, "ms_median": tmed, "GB": gb, "GBps": gb / (t * 1e-3), "frac_hbm": gb / (t * 1e-3) / hbm,
                            "blocks": len(m["meta_mrg"]), "order": list(m["order"])}
                if key == "merge_a":
                    Am = out
                else:
                    Bm = out
            md = s["dot"]["meta_dot"]
            f = lambda: bk.dot(Am, Bm, md, s["dot"]["Dsize"])
            f(); f()
            t, tmed, C = timed(f, flush, args.reps)
            fl = sum((8 if cplx else 2) * Da[0] * Da[1] * Db[1] for (_, _, _, Da, _, Db) in md)
            row["dot"] = {"ms": t, "ms_median": tmed, "GFLOP": fl * 1e-9, "TFLOPs": fl / (t * 1e-3) * 1e-12, "frac_dmma": fl / (t * 1e-3) * 1e-12 / fp64,
                          "sectors": len(md)}
            if s["unmerge"] is not None:
                um = s["unmerge"]["meta"]
                f = lambda: bk.unmerge(C, um)
                f(); f()
                t, tmed, _ = timed(f, flush, args.reps)
                gb = 2 * isz * C.numel() * 1e-9
                row["unmerge"] = {"ms": t, "ms_median": tmed, "GB": gb, "GBps": gb / (t * 1e-3), "frac_hbm": gb / (t * 1e-3) / hbm, "blocks": len(um)}
            rows.append(row)
            print(json.dumps(row), flush=True)
            del A, B, Am, Bm, C
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"hbm_peak_gbs": hbm, "fp64_dmma_peak_tflops": fp64, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
