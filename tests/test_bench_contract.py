"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the keys the driver reads, under
torchrun-style env only rank 0 prints, and the B200 arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--sizes", "1024"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "block-sparse tensordot GFLOP/s" and d["unit"] == "GFLOP/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "yastn"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")     # the real yastn + numpy backend when installed
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "P3" in d["config"]["workload"] and d["config"]["contractions_per_step"] == 3
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_uses_all_host_threads_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1 to its ranks; the reference arm must not inherit that (round-1 SCALE ratios were
    inflated 3x by it)."""
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--sizes", "1024", "--gpus", "2"],
             env={"RANK": "0", "WORLD_SIZE": "2", "OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][0])
    assert d["cpu_baseline"]["threads_env"] == str(os.cpu_count())


def test_both_arms_describe_the_same_config():
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    a = bench.workload_config([1024, 16384], 1)
    assert a == bench.workload_config((1024, 16384), 1) and "16384" in a["workload"]


def test_reference_arm_other_ranks_stay_silent():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--sizes", "1024", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1"])
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr


def test_spmd_dmrg_children_get_their_own_rendezvous(monkeypatch):
    """The SPMD DMRG leg of N > 1 starts one child rank per bench rank: fresh port, and none of torchrun's TORCHELASTIC_* variables
    (with TORCHELASTIC_USE_AGENT_STORE the children would all wait as clients of a store nobody starts: a 15-minute hang once)."""
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    seen = {}

    class Done:
        stdout = json.dumps({"sweep_s": [1.5], "energy": [-2.0], "spmd": {"sharded": 3}, "decomp_stats": {}}) + "\n"

    def fake_run(cmd, capture_output, text, timeout, env):
        seen.update(cmd=cmd, env=env, timeout=timeout)
        return Done()
    monkeypatch.setattr(bench.subprocess, "run", fake_run)
    monkeypatch.setenv("MASTER_PORT", "29500")
    monkeypatch.setenv("TORCHELASTIC_USE_AGENT_STORE", "True")
    monkeypatch.setenv("TORCHELASTIC_RUN_ID", "x")
    monkeypatch.setenv("RANK", "0")
    out = bench.dmrg_sweep_spmd(0, 4)
    assert out["sweep_s"] == 1.5 and out["ranks"] == 4 and out["spmd"] == {"sharded": 3}
    assert seen["env"]["MASTER_PORT"] == "29523" and seen["env"]["RANK"] == "0"
    assert not any(k.startswith("TORCHELASTIC_") for k in seen["env"])
    assert "--spmd" in seen["cmd"] and seen["timeout"] <= 300
    assert bench.dmrg_sweep_spmd(1, 4) is None          # only rank 0 reports
