#!/usr/bin/env python
"""Benchmark of the block-sparse tensordot hot path (BASELINE.json metric: block-sparse tensordot GFLOP/s, fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--dtype f64|c128] [--sizes ...]

Workload (SURVEY.md 8d "config 5"): synthetic U(1) rank-4 tensors A[L*,p,p,L], F[L*,w,L], B4[L*,p*,p*,L] with
gaussian-distributed sector dimensions, total leg dimension D in {1024, 2048, 4096, 8192, 16384} (7..47 charge
sectors, 24..184 blocks), contractions P1 = tensordot(A, F, (3, 0)) and P2 = tensordot(A, B4, ((1,2,3),(2,1,0))).
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
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

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
PATTERNS = ("P1", "P2")


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
# reference arm / cpu baseline: the oracle port (numpy + OpenBLAS) on the host cores
# -------------------------------------------------------------------------------------------------

def cpu_time_cases(cases, cplx, reps):
    from oracle import backend_oracle as orc
    rng = np.random.default_rng(0)
    flops, secs = 0, 0.0
    for name, case in cases:
        A = rng.uniform(-1, 1, case["a"]["size"]); B = rng.uniform(-1, 1, case["b"]["size"])
        if cplx:
            A = A + 1j * rng.uniform(-1, 1, A.size); B = B + 1j * rng.uniform(-1, 1, B.size)
        orc.tensordot_f2m(A, B, case)   # warm (thread pools, page faults)
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            orc.tensordot_f2m(A, B, case)
            best = min(best, time.perf_counter() - t0)
        flops += case_flops(case["f2m"], cplx)
        secs += best
    return flops, secs


def cpu_baseline(cplx, sizes=(1024, 2048, 4096, 8192)):
    cases = load_cases(sizes)
    flops, secs = cpu_time_cases(cases, cplx, reps=2)
    return {"value": flops / secs * 1e-9, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"oracle/backend_oracle.py (numpy+OpenBLAS restatement of the reference numpy backend) on the D={list(sizes)} "
                      f"P1+P2 subset of the sweep, best of 2 per contraction, {flops * 1e-9:.1f} GFLOP"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cplx = args.dtype == "c128"
    sizes = tuple(s for s in args.sizes if s <= 8192) or (min(args.sizes),)
    cases = load_cases(sizes)
    for _ in range(max(args.warmup, 0) and 1):
        cpu_time_cases(cases[:2], cplx, reps=1)
    t0 = time.perf_counter()
    flops_total, secs_total = 0, 0.0
    for _ in range(args.steps):
        f, s = cpu_time_cases(cases, cplx, reps=1)
        flops_total += f
        secs_total += s
    val = flops_total / secs_total * 1e-9
    line = {"impl": "reference", "metric": "block-sparse tensordot GFLOP/s", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs_total / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"synthetic U(1) rank-4 tensordot sweep, D={list(sizes)} (bounded sample of D={list(args.sizes)}), P1+P2, fuse_to_matrix"},
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"oracle port of the reference numpy backend, D={list(sizes)} P1+P2, {args.steps} passes"},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


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
    for name, case in cases:
        stage = case["f2m"]
        total_flops += case_flops(stage, cplx)
        if world > 1:
            stage, _ = sharding.shard_f2m(stage, rank, world)
        def rnd(n):
            x = torch.rand(n, dtype=torch.float64, device=dev, generator=gen) * 2 - 1
            if cplx:
                x = torch.complex(x, torch.rand(n, dtype=torch.float64, device=dev, generator=gen) * 2 - 1)
            return x
        work.append({"name": name, "stage": stage, "A": rnd(case["a"]["size"]), "B": rnd(case["b"]["size"]),
                     "flops": case_flops(stage, cplx)})

    launches = [0]

    fuse = not args.no_fuse

    def contract(w, A, B, ev=None):
        """One fuse_to_matrix tensordot: merge A, merge B, grouped GEMM whose epilogue scatters into the unmerged
        block layout (3 launches; with --no-fuse the unmerge is a 4th launch, exactly the reference's call sequence)."""
        st = w["stage"]
        ma, mb = st["merge_a"], st["merge_b"]
        Am, Bm = A, B
        if ev is not None:
            ev[2].record()
        if ma is not None:
            Am = bk.transpose_and_merge(A, ma["order"], ma["meta_new"], ma["meta_mrg"], ma["Dsize"]); launches[0] += 1
        if mb is not None:
            Bm = bk.transpose_and_merge(B, mb["order"], mb["meta_new"], mb["meta_mrg"], mb["Dsize"]); launches[0] += 1
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
    for _ in range(max(args.warmup, 3)):
        for w in work:
            contract(w, w["A"], w["B"])
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
    gemm_ms = sum(a.elapsed_time(b) for row in gemm_events for (a, b, _) in row)
    merge_ms = sum(c.elapsed_time(a) for row in gemm_events for (a, _, c) in row)
    isz = 16 if cplx else 8
    merge_bytes = 0
    for w in work:
        for key in ("merge_a", "merge_b"):
            m = w["stage"][key]
            if m is not None:
                merge_bytes += isz * (sum(x[1][1] - x[1][0] for x in m["meta_mrg"]) + m["Dsize"])
    own_flops = sum(w["flops"] for w in work)

    # end-to-end: host operands (pinned) -> H2D -> 4 backend calls -> D2H of the result
    e2e = None
    if not args.no_e2e:
        host = [(w["A"].cpu().pin_memory(), w["B"].cpu().pin_memory()) for w in work]
        out_host = [torch.empty(w["stage"]["dot"]["Dsize"], dtype=tdt).pin_memory() for w in work]
        # N > 1: a rank moves only the operand blocks its sectors read and the result blocks it produces (coalesced ranges of
        # the 1-D storage); N = 1: the whole tensors
        ranges = [shard_ranges(w["stage"]) if world > 1 else None for w in work]
        isz_ = 16 if cplx else 8
        if world > 1:
            h2d = sum(isz_ * sum(hi - lo for lo, hi in r[0] + r[1]) for r in ranges)
            d2h = sum(isz_ * sum(hi - lo for lo, hi in r[2]) for r in ranges)
        else:
            h2d = sum(a.numel() * a.element_size() + b.numel() * b.element_size() for a, b in host)
            d2h = sum(o.numel() * o.element_size() for o in out_host)
        e2e_steps = max(1, min(args.steps, 5))
        # three streams: H2D of the next contraction and D2H of the previous one overlap the kernels of the current one
        # (PCIe is full duplex); every contraction still does H2D -> merge/merge/GEMM+unmerge -> D2H inside the timed region
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        def e2e_pass():
            cur = torch.cuda.current_stream(dev)
            for w, (ha, hb), ho, rg in e2e_order:
                with torch.cuda.stream(s_in):
                    if rg is None:
                        A = ha.to(dev, non_blocking=True); B = hb.to(dev, non_blocking=True)
                    else:
                        A = torch.empty(ha.numel(), dtype=tdt, device=dev); B = torch.empty(hb.numel(), dtype=tdt, device=dev)
                        for lo, hi in rg[0]:
                            A[lo:hi].copy_(ha[lo:hi], non_blocking=True)
                        for lo, hi in rg[1]:
                            B[lo:hi].copy_(hb[lo:hi], non_blocking=True)
                    ready = torch.cuda.Event(); ready.record(s_in)
                cur.wait_event(ready)
                A.record_stream(cur); B.record_stream(cur)
                C = contract(w, A, B)
                done = torch.cuda.Event(); done.record(cur)
                s_out.wait_event(done)
                with torch.cuda.stream(s_out):
                    if rg is None:
                        ho.copy_(C, non_blocking=True)
                    else:
                        for lo, hi in rg[2]:
                            ho[lo:hi].copy_(C[lo:hi], non_blocking=True)
                C.record_stream(s_out)
        # largest contraction first: its D2H (2 GB at D=16384) then overlaps the H2D and kernels of the others; consecutive
        # steps are not joined (the copy streams are ordered by events only), the timed region ends when the last D2H lands
        e2e_order = sorted(zip(work, host, out_host, ranges), key=lambda t: -t[2].numel())
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
        gemm_tflops = own_flops * args.steps / (gemm_ms * 1e-3) * 1e-12 if gemm_ms > 0 else 0.0
        line = {"metric": "block-sparse tensordot GFLOP/s", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": f"synthetic U(1) rank-4 block-sparse tensordot sweep (SURVEY 8d config 5): D={list(args.sizes)}, "
                                       f"P1=A.F axes (3,0) and P2=A.B4 axes ((1,2,3),(2,1,0)), fuse_to_matrix pipeline merge/merge/GEMM+unmerge",
                           "contractions_per_step": len(work), "gflop_per_step": total_flops * 1e-9,
                           "l2": "inputs+outputs of the sweep (>10 GB) exceed L2; no flush needed",
                           "sharding": "charge sectors FLOP-balanced over ranks, no collective" if world > 1 else "single GPU"},
                "gpu_launches": n_launch,
                "roofline": {"kernel": "yb::gemm_kernel (grouped DMMA.8x8x4 block GEMM)", "bound": "tensor", "achieved": gemm_tflops,
                             "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": gemm_tflops / FP64_PEAK_TFLOPS, "traffic": traffic, "traffic_of": traffic_of,
                             "peak_source": "measured FP64 DMMA pipe peak, tools/microbench/fp64_pipes.cu (profiles/fp64_peaks_r01.json); cuBLAS DGEMM 8192^3 = 35.5",
                             "gemm_share_of_step": gemm_ms / ms if ms > 0 else None,
                             "epilogue": "fused unmerge scatter" if fuse else "plain store + separate unmerge launch"},
                "roofline_merge": {"kernel": "yb::copy_kernel (transpose_and_merge of A and B)", "bound": "hbm",
                                   "achieved": merge_bytes * args.steps / (merge_ms * 1e-3) * 1e-9 if merge_ms > 0 else None,
                                   "peak": HBM_PEAK_GBS, "unit": "GB/s",
                                   "frac": merge_bytes * args.steps / (merge_ms * 1e-3) * 1e-9 / HBM_PEAK_GBS if merge_ms > 0 else None,
                                   "share_of_step": merge_ms / ms if ms > 0 else None,
                                   "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if HBM_FROM_FILE else "fallback 6650 GB/s (B200_PROFILING.md)"},
                "clocks": clocks}
        if e2e is not None:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(cplx)
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
