#!/usr/bin/env python
"""Host cost of creating device plans for small block structures (what a launch-bound DMRG sweep pays per new structure).

    python tools/plan_create_probe.py [--reps 300]

Times, per call: the table builders (C meta pass), CopyPlan / GemmPlan creation through the C ABI (one pool block + one pinned
upload), and a warm run call.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=300)
    ap.add_argument("--case", default="U1_D64_P1")
    args = ap.parse_args()
    import torch
    from golden_io import bench_structs
    from yastn_b200 import plans, _lib
    st = bench_structs()[args.case]["f2m"]
    m, md, um = st["merge_a"], st["dot"]["meta_dot"], st["unmerge"]["meta"]
    torch.zeros(1, device="cuda")
    out = {"case": args.case, "merge_blocks": len(m["meta_mrg"]), "sectors": len(md), "reps": args.reps}

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        keep = [fn() for _ in range(args.reps)]
        dt = (time.perf_counter() - t) / args.reps * 1e6
        del keep
        return round(dt, 1)
    out["merge_records_us"] = timed(lambda: plans.merge_records(m["order"], m["meta_new"], m["meta_mrg"]))
    out["merge_records_numpy_us"] = timed(lambda: plans.merge_records_np(m["order"], m["meta_new"], m["meta_mrg"]))
    out["scatter_tables_us"] = timed(lambda: plans.unmerge_scatter_tables(md, um))
    out["scatter_tables_numpy_us"] = timed(lambda: plans.unmerge_scatter_tables_np(md, um))
    recs, rank, cov = plans.merge_records(m["order"], m["meta_new"], m["meta_mrg"])
    out["copy_plan_create_us"] = timed(lambda: plans.CopyPlan(recs, rank, 8, 0, cov))
    pr, sg = plans.dot_tables(md)
    sc = plans.unmerge_scatter_tables(md, um)
    out["gemm_plan_create_us"] = timed(lambda: plans.GemmPlan(pr, sg, _lib.YB_F64, 0))
    out["gemm_scatter_plan_create_us"] = timed(lambda: plans.GemmPlan(pr, sg, _lib.YB_F64, 0, sc))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
