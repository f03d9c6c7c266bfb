#!/usr/bin/env python
"""End-to-end CTMRG on the unmodified YASTN (baseline/_ref) with a chosen backend — BASELINE.json config 4 (measurement tool).

    python tools/ctmrg_bench.py --D 5 --chi 256 --sweeps 4 --backend b200|torch|np [--device cuda] [--profile]

Inputs follow SURVEY.md 8(d): a U(1)-symmetric iPEPS tensor on a 1x1 infinite square lattice,
``rand(legs=[lv.conj(), lv, lv, lv.conj(), lp])`` with ``lv = gaussian_leg(D_total=D)`` and the spin-1/2 physical leg,
``EnvCTM(psi, init='eye').ctmrg_(opts_svd={'D_total': chi})`` (yastn/tn/fpeps/envs/_env_ctm.py:862-1000).
Prints one JSON line: seconds per CTM sweep (device synchronised), the environment bond dimension reached, the corner
singular-value drift the reference reports, and how many hot backend calls ran on the B200 kernels.
"""
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
    ap.add_argument("--D", type=int, default=5)
    ap.add_argument("--chi", type=int, default=256)
    ap.add_argument("--sweeps", type=int, default=4)
    ap.add_argument("--backend", default="b200", choices=["b200", "torch", "np"])
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--policy", default="fuse_to_matrix")
    ap.add_argument("--fused", action="store_true")
    ap.add_argument("--chains", action="store_true", help="b200 only: record / replay the double-layer contractions (yastn_b200.chain.enable_peps)")
    ap.add_argument("--decomp-workers", type=int, default=None, help="b200 only: sector streams of svd/qr/eigh (1 = the reference's serial loop)")
    ap.add_argument("--profile", action="store_true", help="time every backend function (device-synchronised: perturbs the totals)")
    ap.add_argument("--shim", action="store_true", help="CPU table interpreter instead of the kernels (host-logic check, tests/cpu_shim.py)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from yastn_loader import load_yastn
    yastn = load_yastn(allow_reference_checkout=False)
    if yastn is None:
        print(json.dumps({"unavailable": "yastn not importable: run tools/install_reference.sh"}))
        return
    import yastn.tn.fpeps as fpeps
    counts = None
    device = "cpu" if args.backend == "np" else args.device
    if args.backend == "b200":
        from yastn_b200 import yastn_backend
        if args.shim:
            import cpu_shim
            cpu_shim.install()
            device = "cpu"
        backend = yastn_backend.module()
        if args.decomp_workers is not None:
            from yastn_b200 import decomp
            decomp.set_workers(args.decomp_workers)
        counts = yastn_backend.call_counts
        if args.fused:
            yastn_backend.enable_fused_tensordot()
        if args.chains:
            from yastn_b200 import chain
            chain.enable_peps()
    else:
        backend = args.backend
    prof = {}
    if args.profile:
        from dmrg_bench import profiled
        backend = profiled(backend, device, prof)
    cfg = yastn.make_config(sym="U1", backend=backend, default_device=device, tensordot_policy=args.policy, default_dtype=args.dtype)
    cfg.backend.random_seed(0)
    lv = yastn.gaussian_leg(cfg, s=1, n=0, sigma=1.0, D_total=args.D, method="round")
    lp = yastn.Leg(cfg, s=1, t=(-1, 1), D=(1, 1))
    A = yastn.rand(cfg, legs=[lv.conj(), lv, lv, lv.conj(), lp], n=0, dtype=args.dtype)
    A = A / A.norm()
    geometry = fpeps.SquareLattice(dims=(1, 1), boundary="infinite")
    psi = fpeps.Peps(geometry=geometry, tensors={(0, 0): A})
    env = fpeps.EnvCTM(psi, init="eye")
    sync = (lambda: None)
    if device != "cpu":
        import torch
        sync = torch.cuda.synchronize
    times, dsv = [], []
    sync()
    t_all = time.perf_counter()
    for info in env.ctmrg_(opts_svd={"D_total": args.chi, "tol": 1e-12}, max_sweeps=args.sweeps, corner_tol=1e-14, iterator=True):
        sync()
        times.append(time.perf_counter() - t_all - sum(times))
        dsv.append(float(info.max_dsv) if info.max_dsv is not None else None)
    chi_reached = max(max(env[(0, 0)].tl.get_shape()), max(env[(0, 0)].t.get_shape()))
    line = {"model": "ctmrg_U1", "D": args.D, "chi": args.chi, "chi_reached": int(chi_reached), "dtype": args.dtype,
            "backend": args.backend + ("+fused" if args.fused else "") + ("+chains" if args.chains else "") + ("+shim" if args.shim else ""), "device": device,
            "policy": args.policy, "sweep_s": times, "max_dsv": dsv, "hot_calls": counts() if counts else None, "decomp_workers": args.decomp_workers}
    if args.backend == "b200" and args.chains:
        line["chains"] = chain.stats()
    if args.profile:
        top = sorted(prof.items(), key=lambda kv: -kv[1][1])[:16]
        line["backend_profile"] = {k: {"calls": v[0], "s": round(v[1], 3)} for k, v in top}
        line["backend_total_s"] = round(sum(v[1] for v in prof.values()), 3)
    print(json.dumps(line))
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
