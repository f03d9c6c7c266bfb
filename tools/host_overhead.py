#!/usr/bin/env python
"""Host-side cost of a hot backend call (measurement tool, GPU box): plan-cache miss (table build + plan creation) vs hit
(lookup + launch), per function, on recorded benchmark structures.  Times are wall-clock per call with the device idle
(synchronised before each measurement, not after: what the Python thread pays)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_io import bench_structs  # noqa: E402
from yastn_b200 import backend_b200 as bk  # noqa: E402


def fresh(x):
    """A structurally equal but new meta object (plan cache is keyed on identity): simulates YASTN's lru miss."""
    return tuple(fresh(y) for y in x) if isinstance(x, tuple) else x


def timeit(fn, reps, args=None):
    """fn() repeated, or fn(x) for every x in args (pre-built fresh metas, so that their construction is not timed)."""
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if args is None:
        for _ in range(reps):
            fn()
    else:
        reps = len(args)
        for x in args:
            fn(x)
    dt = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    return dt * 1e6


def main():
    names = sys.argv[1:] or ["U1_D64_P1", "U1_D1024_P1", "U1xU1_D4096_P1", "U1_D16384_P1"]
    for name in names:
        case = bench_structs()[name]
        st = case["f2m"]
        for dt in (torch.float64, torch.complex128):
            A = torch.rand(case["a"]["size"], dtype=torch.float64, device="cuda").to(dt)
            B = torch.rand(case["b"]["size"], dtype=torch.float64, device="cuda").to(dt)
            ma, mb, md, um = st["merge_a"], st["merge_b"], st["dot"], st["unmerge"]
            Am = bk.transpose_and_merge(A, ma["order"], ma["meta_new"], ma["meta_mrg"], ma["Dsize"]) if ma else A
            Bm = bk.transpose_and_merge(B, mb["order"], mb["meta_new"], mb["meta_mrg"], mb["Dsize"]) if mb else B
            C = bk.dot(Am, Bm, md["meta_dot"], md["Dsize"])
            row = {"case": name, "dtype": str(dt).split(".")[1], "blocks": {"merge_a": len(ma["meta_mrg"]) if ma else 0, "dot": len(md["meta_dot"]),
                                                                           "unmerge": len(um["meta"]) if um else 0}}
            if ma:
                row["merge_hit_us"] = timeit(lambda: bk.transpose_and_merge(A, ma["order"], ma["meta_new"], ma["meta_mrg"], ma["Dsize"]), 20)
                row["merge_miss_us"] = timeit(lambda m: bk.transpose_and_merge(A, ma["order"], ma["meta_new"], m, ma["Dsize"]), 5, [fresh(ma["meta_mrg"]) for _ in range(5)])
            row["dot_hit_us"] = timeit(lambda: bk.dot(Am, Bm, md["meta_dot"], md["Dsize"]), 20)
            row["dot_miss_us"] = timeit(lambda m: bk.dot(Am, Bm, m, md["Dsize"]), 5, [fresh(md["meta_dot"]) for _ in range(5)])
            if um:
                row["unmerge_hit_us"] = timeit(lambda: bk.unmerge(C, um["meta"]), 20)
                row["unmerge_miss_us"] = timeit(lambda m: bk.unmerge(C, m), 5, [fresh(um["meta"]) for _ in range(5)])
                bk.dot_unmerge(Am, Bm, md["meta_dot"], md["Dsize"], um["meta"])
                row["dot_unmerge_hit_us"] = timeit(lambda: bk.dot_unmerge(Am, Bm, md["meta_dot"], md["Dsize"], um["meta"]), 20)
                row["dot_unmerge_miss_us"] = timeit(lambda m: bk.dot_unmerge(Am, Bm, m, md["Dsize"], um["meta"]), 5, [fresh(md["meta_dot"]) for _ in range(5)])
            print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in row.items()}), flush=True)
            del A, B, Am, Bm, C
            bk.clear_plan_cache()


if __name__ == "__main__":
    main()
