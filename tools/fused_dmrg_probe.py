#!/usr/bin/env python
"""Where the time of the fused dot+unmerge path goes inside a DMRG run (measurement tool): table building, plan creation
(C ABI) and the launch itself are timed separately (device-synchronised) by wrapping yastn_b200.plans."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402
from yastn_loader import load_yastn  # noqa: E402

yastn = load_yastn(allow_reference_checkout=False)
import yastn.tn.mps as mps  # noqa: E402
from yastn_b200 import plans, yastn_backend  # noqa: E402
from dmrg_bench import build  # noqa: E402

T = {"tables_s": 0.0, "tables_n": 0, "create_s": 0.0, "create_n": 0, "run_s": 0.0, "run_n": 0, "records_max": 0, "nscat_max": 0,
     "slowest_run_ms": 0.0, "slowest_run_shape": None}


def timed(name, fn, sync=False):
    def f(*a, **k):
        if sync:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(*a, **k)
        if sync:
            torch.cuda.synchronize()
        T[name + "_s"] += time.perf_counter() - t0
        T[name + "_n"] += 1
        return out
    return f


orig_tables = plans.unmerge_scatter_tables


def tables(meta_dot, meta_unmerge, dst_shift=None):
    T["records_max"] = max(T["records_max"], len(meta_unmerge))
    return orig_tables(meta_dot, meta_unmerge, dst_shift)


plans.unmerge_scatter_tables = timed("tables", tables)
orig_init, orig_run = plans.GemmPlan.__init__, plans.GemmPlan.run


def init(self, problems, segments, dtype_code, device, scatter=None):
    t0 = time.perf_counter()
    orig_init(self, problems, segments, dtype_code, device, scatter)
    self._scat = scatter is not None
    self._shape = (len(problems), None if scatter is None else int(len(scatter[6])))
    dt = time.perf_counter() - t0
    if scatter is not None:
        T["create_s"] += dt
        T["create_n"] += 1
        T["nscat_max"] = max(T["nscat_max"], int(len(scatter[1])) - 1)
    else:
        tiles = self.info()["tiles"]
        kind = "plain_create_many_tiles" if tiles > 50000 else "plain_create"
        T[kind + "_s"] = T.get(kind + "_s", 0.0) + dt
        T[kind + "_n"] = T.get(kind + "_n", 0) + 1
        T["tiles_max"] = max(T.get("tiles_max", 0), tiles)


SK = {"skinny_s": 0.0, "skinny_n": 0, "other_s": 0.0, "other_n": 0, "skinny_macs": 0, "other_macs": 0}


def run(self, a, b, c, flags, stream):
    if not self._scat:
        # plain grouped GEMM: split the time by arithmetic density (multiply-adds per tile; a full complex128 64x64 tile with
        # K = 64 has 262144) to see what the tall-and-skinny products of the MPO application cost
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        orig_run(self, a, b, c, flags, stream)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        info = self.info()
        kind = "skinny" if info["tiles"] and info["macs"] / info["tiles"] < 4096 else "other"
        SK[kind + "_s"] += dt; SK[kind + "_n"] += 1; SK[kind + "_macs"] += info["macs"]
        return
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    orig_run(self, a, b, c, flags, stream)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    T["run_s"] += dt
    T["run_n"] += 1
    if dt * 1e3 > T["slowest_run_ms"]:
        T["slowest_run_ms"], T["slowest_run_shape"] = dt * 1e3, (self._shape, self.info())


plans.GemmPlan.__init__, plans.GemmPlan.run = init, run

N, D = int(sys.argv[1]) if len(sys.argv) > 1 else 14, int(sys.argv[2]) if len(sys.argv) > 2 else 4096
backend = yastn_backend.module()
if "--no-fused" not in sys.argv:
    yastn_backend.enable_fused_tensordot()
cfg_kw = dict(backend=backend, default_device="cuda", tensordot_policy="fuse_to_matrix", default_dtype="complex128")
ops, I, H, n_total = build("hubbard", N, cfg_kw, yastn, mps)
ops.random_seed(seed=0)
psi = mps.random_mps(I, n=n_total, D_total=D, dtype="complex128")
t0 = time.perf_counter()
for out in mps.dmrg_(psi, H, method="2site", max_sweeps=1, opts_svd={"tol": 1e-10, "D_total": D}, iterator=True):
    pass
torch.cuda.synchronize()
T["sweep_s"] = time.perf_counter() - t0
T.update(SK)
print(json.dumps(T))
