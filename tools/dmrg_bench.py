#!/usr/bin/env python
"""End-to-end DMRG on the unmodified YASTN (baseline/_ref) with a chosen backend — BASELINE.json configs 1-3 (measurement tool).

    python tools/dmrg_bench.py --model heisenberg|fermions|hubbard --N 32 --D 64 --sweeps 3 --backend b200|torch|np [--device cuda]

heisenberg : U(1) spin-1/2 Heisenberg chain (config 1)         float64
fermions   : Z2 spinless-fermion hopping chain (config 2)      float64
hubbard    : U(1)xU(1) Hubbard chain (config 3)                complex128 (--dtype) / float64
Runs two-site DMRG (yastn/tn/mps/_dmrg.py:42-249) sweep by sweep and prints one JSON line: seconds per sweep (device
synchronised), energy per sweep, number of hot backend calls handled by the B200 kernels, bond dimension reached.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build(model, N, cfg_kw, yastn, mps):
    if model == "heisenberg":
        ops = yastn.operators.Spin12(sym="U1", **cfg_kw)
        I = mps.product_mpo(ops.I(), N)
        terms = []
        for n in range(N - 1):
            terms += [mps.Hterm(1.0, [n, n + 1], [ops.sz(), ops.sz()]), mps.Hterm(0.5, [n, n + 1], [ops.sp(), ops.sm()]),
                      mps.Hterm(0.5, [n, n + 1], [ops.sm(), ops.sp()])]
        n_total = 0
    elif model == "fermions":
        ops = yastn.operators.SpinlessFermions(sym="Z2", **cfg_kw)
        I = mps.product_mpo(ops.I(), N)
        terms = []
        for n in range(N - 1):
            terms += [mps.Hterm(-1.0, [n, n + 1], [ops.cp(), ops.c()]), mps.Hterm(-1.0, [n + 1, n], [ops.cp(), ops.c()])]
        for n in range(N):
            terms.append(mps.Hterm(0.2 * ((n % 3) - 1), [n], [ops.n()]))
        n_total = (N // 2) % 2
    else:
        ops = yastn.operators.SpinfulFermions(sym="U1xU1", **cfg_kw)
        I = mps.product_mpo(ops.I(), N)
        terms = []
        for n in range(N - 1):
            for s in ("u", "d"):
                terms += [mps.Hterm(-1.0, [n, n + 1], [ops.cp(s), ops.c(s)]), mps.Hterm(-1.0, [n + 1, n], [ops.cp(s), ops.c(s)])]
        for n in range(N):
            terms.append(mps.Hterm(4.0, [n], [ops.n("u") @ ops.n("d")]))
        n_total = (N // 2, N // 2)
    H = mps.generate_mpo(I, terms)
    return ops, I, H, n_total


def profiled(backend, device, prof):
    """Copy of the backend module whose public functions are timed (device-synchronised) into ``prof``."""
    import importlib
    import types
    import torch
    if isinstance(backend, str):
        backend = importlib.import_module("yastn.backend.backend_" + backend)
    mod = types.ModuleType("profiled_" + backend.__name__.split(".")[-1])
    for name in dir(backend):
        if not name.startswith("__"):
            setattr(mod, name, getattr(backend, name))

    def wrap(name, fn):
        def f(*a, **k):
            if device != "cpu":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn(*a, **k)
            if device != "cpu":
                torch.cuda.synchronize()
            e = prof.setdefault(name, [0, 0.0])
            e[0] += 1
            e[1] += time.perf_counter() - t0
            return out
        return f
    for name in list(vars(mod)):
        fn = getattr(mod, name)
        if isinstance(fn, types.FunctionType) and not name.startswith("_"):
            setattr(mod, name, wrap(name, fn))
    return mod


def gemm_roofline(backend, cplx_hint):
    """Copy of the backend module whose GEMM entry points (dot / dot_unmerge) are bracketed by CUDA events on the current
    stream (no synchronisation inside the sweep) and whose algorithmic FLOPs are summed from meta_dot
    (2*M*K*N per record, x4 for complex128).  ``report()`` returns TFLOP/s inside the kernel launches."""
    import types
    import torch
    mod = types.ModuleType("roofline_" + backend.__name__)
    for name in dir(backend):
        if not name.startswith("__"):
            setattr(mod, name, getattr(backend, name))
    events, flops_of = [], {}

    def flops(meta_dot, cplx):
        key = id(meta_dot)
        if key not in flops_of:
            flops_of[key] = (meta_dot, sum(2 * Da[0] * Da[1] * Db[1] for (_, _, _, Da, _, Db) in meta_dot))
        return flops_of[key][1] * (4 if cplx else 1)

    # CUDA events right around the launch (yastn_b200.backend_b200._gemm_hook): building the plan of a new structure is host
    # time during which the device may idle, it must not be counted as GEMM time.  A replayed chain launches its GEMMs inside
    # one library call and is not seen here: run without --chains for the complete count.
    from yastn_b200 import backend_b200 as _bk
    current = {"flops": 0, "meta": None}

    def hook(token):
        if token is None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
            return e0
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        if current["meta"] is not None:      # dot / dot_unmerge only (vdot launches through the same entry point)
            events.append((token, e1, current["flops"], current["meta"]))
        current["flops"] = 0          # a call that launches twice (fallbacks) counts its FLOPs once
        return None
    _bk._gemm_hook = hook

    def wrap(fn):
        def f(Adata, Bdata, meta_dot, *rest):
            current["flops"] = flops(meta_dot, Adata.is_complex() or Bdata.is_complex())
            current["meta"] = meta_dot
            try:
                return fn(Adata, Bdata, meta_dot, *rest)
            finally:
                current["meta"] = None
        return f
    for name in ("dot", "dot_unmerge"):
        if hasattr(backend, name):
            setattr(mod, name, wrap(getattr(backend, name)))

    def report():
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b, _, _ in events)
        fl = sum(f for _, _, f, _ in events)
        big = [(a.elapsed_time(b), f) for a, b, f, _ in events if f >= 1e9]
        # time by shape class: (problems, max M, max K, max N) of the call, rounded to powers of two
        import math
        classes = {}
        for a, b, f, md in events:
            if not md:
                continue
            lg = lambda v: 1 << max(0, math.ceil(math.log2(max(v, 1))))
            key = (lg(len(md)), lg(max(r[3][0] for r in md)), lg(max(r[3][1] for r in md)), lg(max(r[5][1] for r in md)))
            e = classes.setdefault(key, [0, 0.0, 0.0])
            e[0] += 1
            e[1] += a.elapsed_time(b)
            e[2] += f
        top = sorted(classes.items(), key=lambda kv: -kv[1][1])[:14]
        by_class = [{"nprob<=": k[0], "M<=": k[1], "K<=": k[2], "N<=": k[3], "calls": v[0], "ms": round(v[1], 2), "gflop": round(v[2] * 1e-9, 1),
                     "tflops": round(v[2] / (v[1] * 1e-3) * 1e-12, 2) if v[1] > 0 else None} for k, v in top]
        ms_big, fl_big = sum(x for x, _ in big), sum(f for _, f in big)
        return {"timed": "CUDA events around every grouped-GEMM launch of dot / dot_unmerge (plan creation excluded; launches inside replayed chains are not seen)",
                "calls": len(events), "gflop": fl * 1e-9, "gemm_s": ms * 1e-3, "tflops": fl / (ms * 1e-3) * 1e-12 if ms > 0 else None,
                "frac_of_37.1": fl / (ms * 1e-3) * 1e-12 / 37.1 if ms > 0 else None,
                "calls_over_1gflop": len(big), "tflops_over_1gflop": fl_big / (ms_big * 1e-3) * 1e-12 if ms_big > 0 else None,
                "share_of_flops_over_1gflop": fl_big / fl if fl else None, "by_shape_class": by_class}
    return mod, report


def device_is_cuda(args):
    return args.device != "cpu" and not args.shim


def time_f2m(prof):
    """Device-synchronised time of every fuse_to_matrix tensordot (whatever implementation is installed: stock, fused, sharded)."""
    import torch
    import yastn.tensor._contractions as C
    inner = C._tensordot_f2m

    def timed(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = inner(*a, **k)
        torch.cuda.synchronize()
        e = prof.setdefault("tensordot_f2m (whole call)", [0, 0.0])
        e[0] += 1
        e[1] += time.perf_counter() - t0
        return out
    C._tensordot_f2m = timed


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="heisenberg")
    ap.add_argument("--N", type=int, default=32)
    ap.add_argument("--D", type=int, default=64)
    ap.add_argument("--sweeps", type=int, default=3)
    ap.add_argument("--backend", default="b200", choices=["b200", "torch", "np"])
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--policy", default="fuse_to_matrix")
    ap.add_argument("--out", default=None)
    ap.add_argument("--D0", type=int, default=None, help="bond dimension of the random initial MPS (default min(D, 32))")
    ap.add_argument("--shim", action="store_true", help="CPU table interpreter instead of the kernels (host-logic check, tests/cpu_shim.py)")
    ap.add_argument("--fused", action="store_true", help="b200 only: fuse dot+unmerge into one launch (yastn_backend.enable_fused_tensordot)")
    ap.add_argument("--chains", action="store_true", help="b200 only: record / replay the tensordot chains of Heff and the environment updates (yastn_b200.chain)")
    ap.add_argument("--spmd", action="store_true", help="b200 only, under torchrun: every rank runs the same sweep, large contractions and the sectors of the "
                    "decompositions are sharded over the ranks (yastn_b200.spmd); rank 0 prints")
    ap.add_argument("--spmd-min-flops", type=float, default=None)
    ap.add_argument("--spmd-nccl", action="store_true", help="exchange the result panels with NCCL broadcasts instead of peer-memory stores")
    ap.add_argument("--gc-freeze", action="store_true", help="gc.freeze() after the model is built (keeps Python's cyclic GC off YASTN's cached metadata)")
    ap.add_argument("--decomp-workers", type=int, default=None, help="b200 only: sector streams of svd/qr/eigh (1 = the reference's serial loop)")
    ap.add_argument("--profile", action="store_true", help="time every backend function (device-synchronised: perturbs the totals)")
    ap.add_argument("--gemm-roofline", action="store_true", help="CUDA-event time and FLOP count of every dot / dot_unmerge launch (no sync inside the sweep)")
    args = ap.parse_args()
    from yastn_loader import load_yastn
    rank, world = 0, 1
    if args.spmd:
        import torch
        import torch.distributed as dist
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        if args.device != "cpu" and not args.shim:
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("gloo" if (args.shim or args.device == "cpu") else "nccl", rank=rank, world_size=world)
    yastn = load_yastn(allow_reference_checkout=False)
    if yastn is None:
        print(json.dumps({"unavailable": "yastn not importable: run tools/install_reference.sh"}))
        return
    import yastn.tn.mps as mps
    counts = None
    if args.backend == "b200":
        from yastn_b200 import yastn_backend
        if args.shim:
            import cpu_shim
            cpu_shim.install()
            args.device = "cpu"
        backend = yastn_backend.module()
        if args.decomp_workers is not None:
            from yastn_b200 import decomp
            decomp.set_workers(args.decomp_workers)
        counts = yastn_backend.call_counts
        if args.fused:
            yastn_backend.enable_fused_tensordot()
        if args.chains:
            from yastn_b200 import chain
            chain.enable()
        if args.spmd:
            from yastn_b200 import spmd
            if args.shim:
                spmd._bk_usable = lambda d: True
            spmd.enable(min_flops=args.spmd_min_flops, peer_arena=not args.spmd_nccl)
            spmd.set_profile(args.profile)
            if device_is_cuda(args):
                import torch.distributed as dist
                warm = torch.zeros(1 << 20, dtype=torch.float64, device="cuda")
                for _ in range(3):                       # NCCL builds its communicator on the first collective: not part of a sweep
                    dist.all_reduce(warm)
                torch.cuda.synchronize()
    else:
        backend = args.backend
    device = "cpu" if args.backend == "np" else args.device
    prof = {}
    roof = None
    if args.gemm_roofline and device != "cpu" and not isinstance(backend, str):
        backend, roof = gemm_roofline(backend, args.dtype == "complex128")
    if args.profile:
        backend = profiled(backend, device, prof)
        if device != "cpu":
            time_f2m(prof)
    cfg_kw = dict(backend=backend, default_device=device, tensordot_policy=args.policy, default_dtype=args.dtype)
    ops, I, H, n_total = build(args.model, args.N, cfg_kw, yastn, mps)
    ops.random_seed(seed=0)
    psi = mps.random_mps(I, n=n_total, D_total=args.D0 or min(args.D, 32), dtype=args.dtype)
    sync = (lambda: None)
    if device != "cpu":
        import torch
        sync = torch.cuda.synchronize
    if args.gc_freeze:
        import gc
        gc.collect()
        gc.freeze()
    times, energies = [], []
    opts_svd = {"tol": 1e-10, "D_total": args.D}
    t_all = time.perf_counter()
    for out in mps.dmrg_(psi, H, method="2site", max_sweeps=args.sweeps, opts_svd=opts_svd, iterator=True):
        sync()
        times.append(time.perf_counter() - t_all - sum(times))
        energies.append(float(out.energy))
    line = {"model": args.model, "N": args.N, "D": args.D, "dtype": args.dtype, "backend": args.backend + ("+fused" if args.fused else ""), "device": device, "policy": args.policy,
            "sweep_s": times, "energy": energies, "bond_dims": max(psi.get_bond_dimensions()),
            "hot_calls": counts() if counts else None, "decomp_workers": args.decomp_workers}
    if args.backend == "b200" and args.chains:
        line["backend"] += "+chains"
        line["chains"] = chain.stats()
    if args.spmd:
        line["spmd"] = dict(spmd.stats(), world=world)
        line["backend"] += f"+spmd{world}"
        from yastn_b200 import decomp
        line["decomp_stats"] = decomp.stats()
    if roof is not None:
        line["gemm_roofline"] = roof()
    if args.profile:
        top = sorted(prof.items(), key=lambda kv: -kv[1][1])[:14]
        line["backend_profile"] = {k: {"calls": v[0], "s": round(v[1], 3)} for k, v in top}
        line["backend_total_s"] = round(sum(v[1] for v in prof.values()), 3)
    if args.spmd:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return
    print(json.dumps(line))
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
