#!/usr/bin/env python
"""Run the reference's OWN hot-path tests (unmodified, from baseline/_ref/ref_tests, see tools/install_reference.sh) against the
B200 kernels: install mode B (`yastn_backend.activate()`) + the reference's conftest options `--backend torch --device cuda`.

    python tools/run_reference_tests.py [--policies fuse_to_matrix fuse_contracted no_fusion] [--shim] [--files ...] [--out FILE]

--shim runs on CPU with the numpy table interpreter standing in for the device (host-logic check, tests/cpu_shim.py).
Prints one JSON summary line per policy (passed / failed / skipped, native vs delegated hot-call counts).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DEFAULT_FILES = ["tensor/test_tensordot.py", "tensor/test_ncon_einsum.py", "tensor/test_fuse_hard.py", "tensor/test_transpose.py",
                 "tensor/test_tensordot_ad.py", "tensor/test_fuse_meta.py", "tensor/test_cache.py", "mps/test_dmrg.py", "mps/test_env.py",
                 "mps/test_tdvp.py",
                 # decompositions (activate() also installs the sector-parallel svd / qr / eigh of yastn_b200.decomp)
                 "tensor/test_svd.py", "tensor/test_qr.py", "tensor/test_eigh.py",
                 # vdot is one grouped-GEMM launch (backend_b200.vdot)
                 "tensor/test_vdot.py"]


class Tally:
    def __init__(self):
        self.c = {"passed": 0, "failed": 0, "skipped": 0}
        self.failed = []

    def pytest_runtest_logreport(self, report):
        if report.when == "call" or (report.when == "setup" and report.outcome != "passed"):
            self.c[report.outcome] = self.c.get(report.outcome, 0) + 1
            if report.outcome == "failed":
                self.failed.append(report.nodeid)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--policies", nargs="+", default=["fuse_to_matrix", "fuse_contracted", "no_fusion"])
    ap.add_argument("--files", nargs="+", default=DEFAULT_FILES)
    ap.add_argument("--shim", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("-k", default=None)
    ap.add_argument("--chains", action="store_true", help="also enable the fused tensordot and the recorded tensordot chains (yastn_b200.chain)")
    args = ap.parse_args()
    import pytest
    from yastn_loader import load_yastn
    if load_yastn(allow_reference_checkout=False) is None:
        print(json.dumps({"unavailable": "yastn not importable: run tools/install_reference.sh first"}))
        return 0
    tests_root = os.path.join(ROOT, "baseline", "_ref", "ref_tests")
    files = [os.path.join(tests_root, f) for f in args.files if os.path.exists(os.path.join(tests_root, f))]
    from yastn_b200 import yastn_backend
    device = "cuda"
    if args.shim:
        import cpu_shim
        cpu_shim.install()
        device = "cpu"
    yastn_backend.activate()
    if args.chains:
        from yastn_b200 import chain
        yastn_backend.enable_fused_tensordot()
        chain.enable()
        chain.enable_peps()
    lines = []
    for policy in args.policies:
        before = yastn_backend.call_counts()
        tally = Tally()
        opts = ["--rootdir", tests_root, "-c", os.devnull, "-q", "-p", "no:cacheprovider", "--backend", "torch", "--device", device,
                "--tensordot_policy", policy, "--no-header", "--tb=short", "-x" if False else "-q"]
        if args.k:
            opts += ["-k", args.k]
        rc = pytest.main(opts + files, plugins=[tally])
        after = yastn_backend.call_counts()
        delta = {kind: {k: after[kind][k] - before[kind][k] for k in after[kind]} for kind in after}
        line = {"policy": policy, "device": device, "rc": int(rc), **tally.c, "failed_ids": tally.failed[:20], "hot_calls": delta, "files": args.files}
        if args.chains:
            line["chains"] = chain.stats()
        print(json.dumps(line), flush=True)
        lines.append(line)
    if args.out:
        with open(args.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
