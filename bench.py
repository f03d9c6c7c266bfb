#!/usr/bin/env python
"""Benchmark of the block-sparse tensordot hot path (BASELINE.json metric: block-sparse tensordot GFLOP/s, fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--dtype f64|c128] [--sizes ...]

Workload (SURVEY.md 8d "config 5"): synthetic U(1) rank-4 tensors A[L*,p,p,L], F[L*,w,L], B4[L*,p*,p*,L] with
gaussian-distributed sector dimensions, total leg dimension D in {1024, 2048, 4096, 8192, 16384} (7..47 charge
sectors, 24..184 blocks), contractions P1 = tensordot(A, F, (3, 0)), P2 = tensordot(A, B4, ((1,2,3),(2,1,0))) and
P3 = tensordot(A.transpose((2,0,3,1)), B4, ((1,2),(3,0))) (both big legs contracted: K ~ 1e7, result blocks <= 4 x 4).
The block structure and all backend metas come from the reference (tests/golden/structs_bench.json.gz, recorded
from yastn's own _meta_* functions); data is uniform(-1, 1) generated at run time.  One "step" is one pass over
the whole sweep (every size x pattern: merge A, merge B, grouped GEMM, unmerge).  Inputs exceed L2 (126 MB) for
the sizes that carry the FLOPs and the sweep cycles through >10 GB between reuses, so no explicit L2 flush is needed.

value  = algorithmic GFLOP/s (sum over meta_dot of 2*M*K*N) with inputs resident in HBM, CUDA-event timed.
e2e    = same metric through the backend API with HOST (pinned) operands: H2D of A and B, the backend calls,
         D2H of the result, every step; copies run on their own streams so that the transfers of neighbouring
         contractions overlap the kernels (full-duplex PCIe).
N > 1  = the charge sectors of every contraction are sharded FLOP-balanced over the ranks (no collective on
         the data path); value = total FLOPs / max-over-ranks time ("strong" scaling).  In the e2e leg every rank moves
         only the operand blocks its sectors read and the result blocks it writes; the byte counts are job totals.
--impl reference = the UNMODIFIED reference (yastn from baseline/_ref, numpy backend, OpenBLAS on all host cores) running
         yastn.tensordot on the SAME sizes, patterns and dtype: tensors built with the same recipe the golden structures were
         recorded from (tests/golden/make_golden.py), every step one full pass.
gpu_baseline = the reference's stock torch backend on the same GPU (loop of cuBLAS calls and slice copies,
         yastn/backend/_backend_torch_backwards.py:100-109,340-364,397-408) through yastn.tensordot, and our backend module
         through the very same yastn.tensordot calls.
dmrg_sweep_s = one 2-site DMRG sweep of the U(1)xU(1) Hubbard chain at D=4096 complex128 (BASELINE config 3 shape, N=20 sites)
         on the unmodified YASTN with our backend module (fused dot+unmerge, recorded tensordot chains) and with the stock torch
         backend on the same GPU.  N > 1 adds spmd_dmrg_sweep: the same sweep run SPMD on all ranks (yastn_b200.spmd).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv[1:]:
    # the reference arm uses every host core: torchrun exports OMP_NUM_THREADS=1 to its ranks, which would throttle OpenBLAS
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count())

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FP64_PEAK_TFLOPS = 37.1   # measured DMMA.8x8x4 pipe peak on this pool's B200 (profiles/fp64_peaks_r01.json);
                          # MEASURED_PEAKS.json carries no FP64 figure (bf16 only)
try:
    HBM_PEAK_GBS = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    HBM_FROM_FILE = True
except Exception:
    HBM_PEAK_GBS, HBM_FROM_FILE = 6650.0, False
DEFAULT_SIZES = (1024, 2048, 4096, 8192, 16384)
PATTERNS = ("P1", "P2", "P3")
SIGMA = {64: 0.7, 1024: 1.0, 2048: 1.5, 4096: 2.5, 8192: 4.0, 16384: 6.0}    # gaussian_leg widths of the recorded structures


def load_cases(sizes):
    from golden_io import bench_structs
    st = bench_structs()
    return [(f"U1_D{d}_{p}", st[f"U1_D{d}_{p}"]) for d in sizes for p in PATTERNS]


def case_flops(stage, cplx):
    return sum((8 if cplx else 2) * Da[0] * Da[1] * Db[1] for (_, _, _, Da, _, Db) in stage["dot"]["meta_dot"])


def _coalesce(slices):
    out = []
    for lo, hi in sorted(s for s in slices if s[1] > s[0]):
        if out and lo <= out[-1][1]:
            out[-1][1] = max(out[-1][1], hi)
        else:
            out.append([lo, hi])
    return [tuple(x) for x in out]


def shard_ranges(stage):
    """Storage ranges a rank touches for one sharded contraction: (A blocks read, B blocks read, result blocks written)."""
    md = stage["dot"]["meta_dot"]
    a = [r[1] for r in stage["merge_a"]["meta_mrg"]] if stage["merge_a"] is not None else [r[2] for r in md]
    b = [r[1] for r in stage["merge_b"]["meta_mrg"]] if stage["merge_b"] is not None else [r[4] for r in md]
    c = [r[0] for r in stage["unmerge"]["meta"]] if stage["unmerge"] is not None else [r[0] for r in md]
    return _coalesce(a), _coalesce(b), _coalesce(c)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
# reference arm / cpu baseline / gpu baseline: the unmodified reference (baseline/_ref) through yastn.tensordot
# -------------------------------------------------------------------------------------------------

def workload_config(sizes, world):
    """The `config` object of the JSON line; identical for both arms (same workload, same sizes)."""
    cases = load_cases(tuple(sizes))
    return {"workload": f"synthetic U(1) rank-4 block-sparse tensordot sweep (SURVEY 8d config 5): D={list(sizes)}, "
                        f"P1=A.F axes (3,0), P2=A.B4 axes ((1,2,3),(2,1,0)), P3=A^T.B4 axes ((1,2),(3,0)); fuse_to_matrix policy "
                        f"(merge, merge, sector GEMMs, unmerge)",
            "contractions_per_step": len(cases),
            "l2": "operands + results of one step (>10 GB) exceed L2 and the host caches; no flush needed",
            "sharding": "charge sectors FLOP-balanced over ranks, no collective" if world > 1 else "single GPU"}


def load_reference():
    from yastn_loader import load_yastn
    return load_yastn(allow_reference_checkout=False)


def yastn_workload(yastn, cfg, sizes, cplx):
    """The synthetic tensors of the sweep built by the reference itself (recipe of SURVEY 8d = tests/golden/make_golden.py:
    gaussian_leg + rand); returns [(name, a, b, axes)] in the order of load_cases()."""
    dtype = "complex128" if cplx else "float64"
    cfg.backend.random_seed(0)
    out = []
    for D in sizes:
        L = yastn.gaussian_leg(cfg, s=1, n=0, sigma=SIGMA[D], D_total=D, method="round")
        p = yastn.Leg(cfg, s=1, t=(-1, 1), D=(1, 1))
        w = yastn.Leg(cfg, s=1, t=(-2, 0, 2), D=(1, 3, 1))
        A = yastn.rand(cfg, legs=[L.conj(), p, p, L], n=0, dtype=dtype)
        F = yastn.rand(cfg, legs=[L.conj(), w, L], n=0, dtype=dtype)
        B4 = yastn.rand(cfg, legs=[L.conj(), p.conj(), p.conj(), L], n=0, dtype=dtype)
        ops = {"P1": (A, F, (3, 0)), "P2": (A, B4, ((1, 2, 3), (2, 1, 0))), "P3": (A.transpose((2, 0, 3, 1)), B4, ((1, 2), (3, 0)))}
        for pat in PATTERNS:
            out.append((f"U1_D{D}_{pat}",) + ops[pat])
    return out


def check_same_structure(work, cases):
    """The tensors the reference built have exactly the recorded block structure (same config on both arms)."""
    for (name, a, b, _), (cname, case) in zip(work, cases):
        assert name == cname and a.size == case["a"]["size"] and b.size == case["b"]["size"], (name, a.size, case["a"]["size"])


def time_yastn_sweep(yastn, work, passes, sync=None):
    """Seconds per pass of yastn.tensordot over the sweep (1 untimed pass warms YASTN's meta caches, thread pools, pages)."""
    for _, a, b, axes in work:
        yastn.tensordot(a, b, axes=axes)
    if sync:
        sync()
    t0 = time.perf_counter()
    for _ in range(passes):
        for _, a, b, axes in work:
            yastn.tensordot(a, b, axes=axes)
    if sync:
        sync()
    return (time.perf_counter() - t0) / passes


def cpu_time_port(cases, cplx, reps):
    """Fallback when baseline/_ref is absent: the oracle port of the numpy backend (kind "port")."""
    from oracle import backend_oracle as orc
    rng = np.random.default_rng(0)
    secs = 0.0
    for name, case in cases:
        A = rng.uniform(-1, 1, case["a"]["size"]); B = rng.uniform(-1, 1, case["b"]["size"])
        if cplx:
            A = A + 1j * rng.uniform(-1, 1, A.size); B = B + 1j * rng.uniform(-1, 1, B.size)
        orc.tensordot_f2m(A, B, case)
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            orc.tensordot_f2m(A, B, case)
            best = min(best, time.perf_counter() - t0)
        secs += best
    return secs


def cpu_reference_seconds(sizes, cplx, passes):
    """(seconds per pass, kind, description) of the reference's CPU path on this box."""
    cases = load_cases(tuple(sizes))
    yastn = load_reference()
    if yastn is None:
        return cpu_time_port(cases, cplx, max(1, passes)), "port", "oracle/backend_oracle.py (numpy restatement; baseline/_ref not present)"
    cfg = yastn.make_config(sym="U1", backend="np", tensordot_policy="fuse_to_matrix")
    work = yastn_workload(yastn, cfg, sizes, cplx)
    check_same_structure(work, cases)
    return time_yastn_sweep(yastn, work, passes), "reference", "unmodified yastn (baseline/_ref) numpy backend, yastn.tensordot, OpenBLAS"


def cpu_baseline(cplx, sizes):
    cases = load_cases(tuple(sizes))
    flops = sum(case_flops(c["f2m"], cplx) for _, c in cases)
    secs, kind, what = cpu_reference_seconds(sizes, cplx, passes=1)
    return {"value": flops / secs * 1e-9, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": kind,
            "sample": f"{what}; one full pass of the sweep D={list(sizes)} x P1,P2,P3 after one warm-up pass, {flops * 1e-9:.1f} GFLOP in {secs:.2f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cplx = args.dtype == "c128"
    sizes = tuple(args.sizes)
    cases = load_cases(sizes)
    flops = sum(case_flops(c["f2m"], cplx) for _, c in cases)
    t0 = time.perf_counter()
    yastn = load_reference()
    if yastn is not None:
        cfg = yastn.make_config(sym="U1", backend="np", tensordot_policy="fuse_to_matrix")
        work = yastn_workload(yastn, cfg, sizes, cplx)
        check_same_structure(work, cases)
        secs = time_yastn_sweep(yastn, work, args.steps)
        kind, what = "reference", "unmodified yastn (baseline/_ref), numpy backend, yastn.tensordot, OpenBLAS on all host cores"
    else:
        secs = cpu_time_port(cases, cplx, 1)
        kind, what = "port", "oracle port of the reference numpy backend (baseline/_ref not present)"
    val = flops / secs * 1e-9
    cfg_line = workload_config(sizes, 1)
    cfg_line["gflop_per_step"] = flops * 1e-9
    line = {"impl": "reference", "metric": "block-sparse tensordot GFLOP/s", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": cfg_line,
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": kind,
                             "sample": f"{what}; {args.steps} full passes of the sweep after one warm-up pass",
                             "threads_env": os.environ.get("OMP_NUM_THREADS")},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


def gpu_baseline(sizes, cplx, dev, flops):
    """Same contractions through yastn.tensordot on the GPU: the reference's stock torch backend, and our backend module."""
    import torch
    yastn = load_reference()
    if yastn is None:
        return {"unavailable": "baseline/_ref not present"}
    from yastn_b200 import yastn_backend
    cases = load_cases(tuple(sizes))
    out = {"unit": "GFLOP/s", "api": "yastn.tensordot, fuse_to_matrix, tensors resident on the GPU, wall clock around 3 passes with device sync"}
    for key, backend in (("stock_torch", "torch"), ("ours", yastn_backend.module())):
        cfg = yastn.make_config(sym="U1", backend=backend, default_device=str(dev), tensordot_policy="fuse_to_matrix")
        work = yastn_workload(yastn, cfg, sizes, cplx)
        check_same_structure(work, cases)
        secs = time_yastn_sweep(yastn, work, 3, sync=lambda: torch.cuda.synchronize(dev))
        out[key] = flops / secs * 1e-9
        out[key + "_ms_per_step"] = secs * 1e3
        del work
        torch.cuda.empty_cache()
    out["value"] = out["stock_torch"]
    out["kind"] = "unmodified yastn (baseline/_ref) stock backend_torch on the same GPU"
    return out


def dmrg_sweep_spmd(rank, world, timeout=300):
    """The same DMRG sweep run SPMD on all ranks of this job (yastn_b200.spmd: contractions sharded by row panels, SVD sectors
    dealt to the ranks, results completed by NCCL all-reduces over NVLink).  Every bench rank starts one child rank (fresh
    process, own rendezvous port); rank 0 returns the child's line."""
    # own rendezvous: a fresh TCP store on another port, started by the child of rank 0 (torchrun's agent store must not be reused:
    # with TORCHELASTIC_USE_AGENT_STORE the child ranks would all wait as clients of a server nobody starts)
    env = {k: v for k, v in os.environ.items() if not k.startswith("TORCHELASTIC_")}
    env["MASTER_PORT"] = str(int(env.get("MASTER_PORT", "29500")) + 23)
    cmd = [sys.executable, os.path.join(ROOT, "tools", "dmrg_bench.py"), "--model", "hubbard", "--N", "20", "--D", "4096", "--D0", "4096",
           "--sweeps", "1", "--dtype", "complex128", "--backend", "b200", "--fused", "--chains", "--spmd"]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        if rank != 0:
            return None
        d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
        return {"sweep_s": d["sweep_s"][0], "energy": d["energy"][0], "ranks": world, "spmd": d.get("spmd"), "decomp": d.get("decomp_stats"),
                "config": "U(1)xU(1) Hubbard N=20, 2-site DMRG, D=4096, complex128, one sweep; every rank runs the unmodified yastn program, "
                          "tensordots above 4 GFLOP sharded by row panels + peer-memory panel exchange, SVD sectors sharded + all-reduce"}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"} if rank == 0 else None


def dmrg_sweep(which, timeout=900):
    """One 2-site DMRG sweep, U(1)xU(1) Hubbard N=20 D=4096 complex128, in a fresh process (tools/dmrg_bench.py)."""
    cmd = [sys.executable, os.path.join(ROOT, "tools", "dmrg_bench.py"), "--model", "hubbard", "--N", "20", "--D", "4096", "--D0", "4096",
           "--sweeps", "1", "--dtype", "complex128", "--backend", which] + (["--fused", "--chains", "--gemm-roofline"] if which == "b200" else [])
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
        d = json.loads(line)
        if "sweep_s" not in d:
            return d
        out = {"sweep_s": d["sweep_s"][0], "energy": d["energy"][0]}
        if d.get("gemm_roofline"):
            out["gemm"] = d["gemm_roofline"]
        return out
    except Exception as e:   # a missing figure must not take the bench line down
        return {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}


# -------------------------------------------------------------------------------------------------
# B200 arm
# -------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "c128"])
    ap.add_argument("--sizes", type=int, nargs="+", default=list(DEFAULT_SIZES))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-torch-backend-on-GPU leg")
    ap.add_argument("--no-dmrg", action="store_true", help="skip the DMRG D=4096 sweep legs (about 2.5 minutes)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fuse", action="store_true", help="run unmerge as its own launch instead of the GEMM scatter epilogue")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly one JSON line: everything else that libraries print there (NCCL's version banner is written
    # to fd 1 whatever NCCL_DEBUG_FILE says) is sent to stderr by swapping the descriptors for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from yastn_b200 import backend_b200 as bk
    from yastn_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cplx = args.dtype == "c128"
    tdt = torch.complex128 if cplx else torch.float64

    cases = load_cases(tuple(args.sizes))
    work = []   # per contraction: sharded stage metas, device operands, flops
    total_flops = 0
    gen = torch.Generator(device=dev).manual_seed(1234)

    def rnd(n):
        x = torch.rand(n, dtype=torch.float64, device=dev, generator=gen) * 2 - 1
        if cplx:
            x = torch.complex(x, torch.rand(n, dtype=torch.float64, device=dev, generator=gen) * 2 - 1)
        return x
    # the three contractions of one size share their operands, as in the reference arm: P1 = A.F, P2 = A.B4, P3 = A^T.B4
    # (the transpose is lazy: same storage)
    operands = {}
    for name, case in cases:
        size_tag, pat = name.rsplit("_", 1)
        akey, bkey = (size_tag, "A"), (size_tag, "F" if pat == "P1" else "B4")
        for key, n in ((akey, case["a"]["size"]), (bkey, case["b"]["size"])):
            if key not in operands:
                operands[key] = rnd(n)
            assert operands[key].numel() == n
        stage = case["f2m"]
        total_flops += case_flops(stage, cplx)
        if world > 1:
            stage, _ = sharding.shard_f2m(stage, rank, world)
        work.append({"name": name, "stage": stage, "A": operands[akey], "B": operands[bkey], "akey": akey, "bkey": bkey,
                     "flops": case_flops(stage, cplx)})

    launches = [0]

    fuse = not args.no_fuse

    def count_merge(data, m):
        """Kernels one merge launches: copy_kernel and / or tiled_kernel (counted once per plan, outside the timed region)."""
        info = bk._merge_plans(data, m["order"], m["meta_new"], m["meta_mrg"], m["Dsize"])["fwd"].info()
        return (1 if info["records"] > info["tiled_records"] else 0) + (1 if info["tiled_records"] > 0 else 0)

    # N > 1: a rank merges only the blocks its panels read; the rest of the merged buffer is never read, so it is not cleared
    # either (the full-size call would memset the whole destination on every rank: merge time did not shrink with N in round 1)
    merge = bk.transpose_and_merge_partial if world > 1 else bk.transpose_and_merge

    def contract(w, A, B, ev=None):
        """One fuse_to_matrix tensordot: merge A, merge B, grouped GEMM whose epilogue scatters into the unmerged
        block layout (3 launches; with --no-fuse the unmerge is a 4th launch, exactly the reference's call sequence)."""
        st = w["stage"]
        ma, mb = st["merge_a"], st["merge_b"]
        Am, Bm = A, B
        if ev is not None:
            ev[2].record()
        if ma is not None:
            Am = merge(A, ma["order"], ma["meta_new"], ma["meta_mrg"], ma["Dsize"]); launches[0] += w["launches_a"]
        if mb is not None:
            Bm = merge(B, mb["order"], mb["meta_new"], mb["meta_mrg"], mb["Dsize"]); launches[0] += w["launches_b"]
        if ev is not None:
            ev[0].record()
        if fuse and st["unmerge"] is not None:
            C = bk.dot_unmerge(Am, Bm, st["dot"]["meta_dot"], st["dot"]["Dsize"], st["unmerge"]["meta"]); launches[0] += 1
            if ev is not None:
                ev[1].record()
            return C
        C = bk.dot(Am, Bm, st["dot"]["meta_dot"], st["dot"]["Dsize"]); launches[0] += 1
        if ev is not None:
            ev[1].record()
        if st["unmerge"] is not None:
            C = bk.unmerge(C, st["unmerge"]["meta"]); launches[0] += 1
        return C

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the metas come from JSON (fresh tuples): plans are cached on their identity, i.e. built once in warm-up
    for w in work:
        w["launches_a"] = w["launches_b"] = 1
    for it in range(max(args.warmup, 3)):
        for w in work:
            contract(w, w["A"], w["B"])
            if it == 0:
                st = w["stage"]
                w["launches_a"] = count_merge(w["A"], st["merge_a"]) if st["merge_a"] is not None else 0
                w["launches_b"] = count_merge(w["B"], st["merge_b"]) if st["merge_b"] is not None else 0
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    gemm_events = [[tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in work] for _ in range(args.steps)]
    launches[0] = 0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        for w, ev in zip(work, gemm_events[s]):
            contract(w, w["A"], w["B"], ev)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    n_launch = launches[0]
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel shares from the events recorded inside the timed region.  P3 contractions run on the skinny (HBM-bound)
    # kernel, P1 / P2 on the DMMA tile kernel: they are reported against different rooflines.
    skinny = [w["name"].endswith("P3") for w in work]
    gemm_ms = sum(a.elapsed_time(b) for row in gemm_events for (a, b, _), sk in zip(row, skinny) if not sk)
    skinny_ms = sum(a.elapsed_time(b) for row in gemm_events for (a, b, _), sk in zip(row, skinny) if sk)
    merge_ms = sum(c.elapsed_time(a) for row in gemm_events for (a, _, c) in row)
    isz = 16 if cplx else 8
    merge_bytes = 0
    skinny_bytes = 0
    for w, sk in zip(work, skinny):
        for key in ("merge_a", "merge_b"):
            m = w["stage"][key]
            if m is not None:
                # read: the source blocks; written: the merged blocks of this rank (all of them at N = 1: Dsize)
                merge_bytes += isz * (sum(x[1][1] - x[1][0] for x in m["meta_mrg"]) + sum(x[2][1] - x[2][0] for x in m["meta_new"]))
        if sk:   # every element of the two merged operands is read once; the result is a few numbers
            skinny_bytes += isz * sum(Da[0] * Da[1] + Db[0] * Db[1] for (_, _, _, Da, _, Db) in w["stage"]["dot"]["meta_dot"])
    own_flops = sum(w["flops"] for w in work)
    tile_flops = sum(w["flops"] for w, sk in zip(work, skinny) if not sk)

    # end-to-end: host operands (pinned) -> H2D -> backend calls -> D2H of every result, every step.  The three contractions of
    # one size share their operands, so a step uploads each of A, F, B4 once per size (what a user of the API does as well)
    e2e = None
    if not args.no_e2e:
        host_ops = {k: v.cpu().pin_memory() for k, v in operands.items()}
        out_host = [torch.empty(w["stage"]["dot"]["Dsize"], dtype=tdt).pin_memory() for w in work]
        # N > 1: a rank moves only the operand blocks its sectors read and the result blocks it produces (coalesced ranges of
        # the 1-D storage); N = 1: the whole tensors
        ranges = [shard_ranges(w["stage"]) if world > 1 else None for w in work]
        isz_ = 16 if cplx else 8
        groups = {}          # size tag -> contractions of that size
        for i, w in enumerate(work):
            groups.setdefault(w["akey"][0], []).append(i)
        need = {}            # operand -> ranges to upload (None: everything)
        for i, w in enumerate(work):
            for key, which in ((w["akey"], 0), (w["bkey"], 1)):
                if ranges[i] is None:
                    need[key] = None
                else:
                    need[key] = _coalesce(list(need.get(key) or []) + list(ranges[i][which]))
        if world > 1:
            h2d = sum(isz_ * sum(hi - lo for lo, hi in r) for r in need.values())
            d2h = sum(isz_ * sum(hi - lo for lo, hi in r[2]) for r in ranges)
        else:
            h2d = sum(v.numel() * v.element_size() for v in host_ops.values())
            d2h = sum(o.numel() * o.element_size() for o in out_host)
        e2e_steps = max(1, min(args.steps, 5))
        # three streams: H2D of the next size and D2H of the previous results overlap the kernels of the current size
        # (PCIe is full duplex); every step still does H2D -> merge/merge/GEMM+unmerge -> D2H inside the timed region
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        # largest size first: its D2H (2 GB at D=16384) then overlaps the H2D and kernels of the others; consecutive
        # steps are not joined (the copy streams are ordered by events only), the timed region ends when the last D2H lands
        group_order = sorted(groups, key=lambda g: -max(out_host[i].numel() for i in groups[g]))

        def e2e_pass():
            cur = torch.cuda.current_stream(dev)
            for gname in group_order:
                devops = {}
                with torch.cuda.stream(s_in):
                    for key in sorted({k for i in groups[gname] for k in (work[i]["akey"], work[i]["bkey"])}):
                        h = host_ops[key]
                        if need[key] is None:
                            devops[key] = h.to(dev, non_blocking=True)
                        else:
                            t = torch.empty(h.numel(), dtype=tdt, device=dev)
                            for lo, hi in need[key]:
                                t[lo:hi].copy_(h[lo:hi], non_blocking=True)
                            devops[key] = t
                    ready = torch.cuda.Event(); ready.record(s_in)
                cur.wait_event(ready)
                for t in devops.values():
                    t.record_stream(cur)
                for i in sorted(groups[gname], key=lambda i: -out_host[i].numel()):
                    w, ho, rg = work[i], out_host[i], ranges[i]
                    C = contract(w, devops[w["akey"]], devops[w["bkey"]])
                    done = torch.cuda.Event(); done.record(cur)
                    s_out.wait_event(done)
                    with torch.cuda.stream(s_out):
                        if rg is None:
                            ho.copy_(C, non_blocking=True)
                        else:
                            for lo, hi in rg[2]:
                                ho[lo:hi].copy_(C[lo:hi], non_blocking=True)
                    C.record_stream(s_out)

        def e2e_join():
            cur = torch.cuda.current_stream(dev)
            cur.wait_stream(s_in)
            cur.wait_stream(s_out)
        e2e_pass()
        e2e_join()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        s_in.wait_stream(torch.cuda.current_stream(dev))
        s_out.wait_stream(torch.cuda.current_stream(dev))
        for _ in range(e2e_steps):
            e2e_pass()
        e2e_join()
        f1.record()
        barrier()
        e2e_ms = f0.elapsed_time(f1) / e2e_steps
    if world > 1:
        t = torch.tensor([ms, gemm_ms, float(own_flops), e2e_ms if not args.no_e2e else 0.0], dtype=torch.float64, device=dev)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms, e2e_max = float(tmax[0]), float(tmax[3])
        if not args.no_e2e:       # bytes moved by the whole job
            tb = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device=dev)
            dist.all_reduce(tb)
            h2d, d2h = int(tb[0]), int(tb[1])
    else:
        e2e_max = e2e_ms if not args.no_e2e else 0.0
    if not args.no_e2e:
        e2e = {"value": total_flops / (e2e_max * 1e-3) * 1e-9, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_max}

    spmd_dmrg = None
    if world > 1 and not args.no_dmrg:
        # second workload at N > 1: a whole DMRG sweep on all ranks, with the NCCL exchange inside the timed sweep
        del work, operands
        if not args.no_e2e:
            del host_ops, out_host
        torch.cuda.empty_cache()
        barrier()
        spmd_dmrg = dmrg_sweep_spmd(rank, world)
    if rank == 0:
        # DRAM bytes of the dominant launch from the committed ncu --set full capture (not measurable outside a profiler)
        traffic = traffic_of = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["gemm"]
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            traffic_of = f"{tr['launch']}; algorithmic bytes of that launch {tr['algorithmic_bytes']:.3e}; {tr['source']}"
        except Exception:
            pass
        ms_step = ms / args.steps
        gflops = total_flops / (ms_step * 1e-3) * 1e-9
        gemm_tflops = tile_flops * args.steps / (gemm_ms * 1e-3) * 1e-12 if gemm_ms > 0 else 0.0
        cfg_line = workload_config(args.sizes, world)
        cfg_line["gflop_per_step"] = total_flops * 1e-9
        line = {"metric": "block-sparse tensordot GFLOP/s", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": cfg_line,
                "gpu_launches": n_launch,
                "roofline": {"kernel": "yb::gemm_kernel (grouped DMMA.8x8x4 block GEMM; P1 and P2 contractions)", "bound": "tensor", "achieved": gemm_tflops,
                             "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": gemm_tflops / FP64_PEAK_TFLOPS, "traffic": traffic, "traffic_of": traffic_of,
                             "peak_source": "measured FP64 DMMA pipe peak, tools/microbench/fp64_pipes.cu (profiles/fp64_peaks_r01.json); cuBLAS DGEMM 8192^3 = 35.5; "
                                            "MEASURED_PEAKS.json has no FP64 figure",
                             "gemm_share_of_step": gemm_ms / ms if ms > 0 else None,
                             "epilogue": "fused unmerge scatter" if fuse else "plain store + separate unmerge launch"},
                "roofline_merge": {"kernel": "yb::copy_kernel (transpose_and_merge of A and B)", "bound": "hbm",
                                   "achieved": merge_bytes * args.steps / (merge_ms * 1e-3) * 1e-9 if merge_ms > 0 else None,
                                   "peak": HBM_PEAK_GBS, "unit": "GB/s",
                                   "frac": merge_bytes * args.steps / (merge_ms * 1e-3) * 1e-9 / HBM_PEAK_GBS if merge_ms > 0 else None,
                                   "share_of_step": merge_ms / ms if ms > 0 else None,
                                   "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if HBM_FROM_FILE else "fallback 6650 GB/s (B200_PROFILING.md)"},
                "roofline_skinny": {"kernel": "yb::skinny_kernel (P3 contractions: K ~ 1e7, result blocks <= 4 x 4)", "bound": "hbm",
                                    "achieved": skinny_bytes * args.steps / (skinny_ms * 1e-3) * 1e-9 if skinny_ms > 0 else None,
                                    "peak": HBM_PEAK_GBS, "unit": "GB/s",
                                    "frac": skinny_bytes * args.steps / (skinny_ms * 1e-3) * 1e-9 / HBM_PEAK_GBS if skinny_ms > 0 else None,
                                    "share_of_step": skinny_ms / ms if ms > 0 else None},
                "clocks": clocks}
        if e2e is not None:
            line["e2e"] = e2e
        if spmd_dmrg is not None:
            line["spmd_dmrg_sweep"] = spmd_dmrg
        if world == 1:
            del work, operands
            torch.cuda.empty_cache()
            if not args.no_cpu_baseline:
                line["cpu_baseline"] = cpu_baseline(cplx, args.sizes)
            if not args.no_gpu_baseline:
                try:
                    line["gpu_baseline"] = gpu_baseline(args.sizes, cplx, dev, total_flops)
                except Exception as e:
                    line["gpu_baseline"] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
            if not args.no_dmrg:
                torch.cuda.empty_cache()
                ours, stock = dmrg_sweep("b200"), dmrg_sweep("torch")
                line["dmrg_sweep_s"] = {"config": "U(1)xU(1) Hubbard chain N=20, 2-site DMRG, D=4096, complex128, one sweep from a random D=4096 MPS "
                                                  "(BASELINE config 3 shape), unmodified yastn (baseline/_ref), fuse_to_matrix",
                                        "ours": ours, "stock_torch_same_gpu": stock}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
