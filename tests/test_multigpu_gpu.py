"""Multi-GPU data path on real devices (needs >= 2 GPUs on the box, skipped otherwise): two sharded contractions with the block
exchange in between — NCCL send/recv, the peer-memory copy kernel and the fused GEMM epilogue — each checked inside
tools/multigpu_chain.py against the unsharded chain (rel. Frobenius error <= 1e-12).  The host logic of the same path runs on CPU
in tests/test_peer_tables.py and tests/test_sharding_gloo.py."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["f64", "c128"])
def test_sharded_chain_with_peer_exchange(dtype):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "multigpu_chain.py"), "--case", "U1_D1024_chain", "--dtype", dtype, "--iters", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert {ln["variant"] for ln in lines} >= {"nccl", "exchange", "fused"}
    for ln in lines:
        if "rel_err_E" in ln:
            assert ln["rel_err_E"] <= 1e-12 and ln["rel_err_C_blocks"] <= 1e-12


@pytest.mark.gpu
def test_spmd_dmrg_two_gpus_same_energy():
    """An unmodified YASTN 2-site DMRG (Z2 fermions, D=256) run SPMD on two GPUs through yastn_b200.spmd — contractions sharded by
    row panels, SVD sectors dealt to the ranks, results completed by NCCL all-reduces — reaches the single-GPU energy to 1e-12."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    tool = os.path.join(ROOT, "tools", "dmrg_bench.py")
    common = ["--model", "fermions", "--N", "16", "--D", "256", "--D0", "64", "--sweeps", "2", "--backend", "b200", "--fused"]
    one = subprocess.run([sys.executable, tool] + common, capture_output=True, text=True, timeout=600)
    assert one.returncode == 0, one.stdout[-2000:] + one.stderr[-2000:]
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), tool] + common + ["--spmd", "--spmd-min-flops", "1e6"], capture_output=True, text=True, timeout=600)
    assert two.returncode == 0, two.stdout[-2000:] + two.stderr[-2000:]
    a = [json.loads(ln) for ln in one.stdout.splitlines() if ln.startswith("{")][-1]
    b = [json.loads(ln) for ln in two.stdout.splitlines() if ln.startswith("{")][-1]
    assert b["spmd"]["sharded"] > 0 and b["spmd"]["world"] == 2
    for x, y in zip(a["energy"], b["energy"]):
        assert abs(x - y) <= 1e-12 * abs(x)
