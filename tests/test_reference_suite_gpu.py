"""The reference's OWN test files (unmodified, from baseline/_ref/ref_tests) run on the B200 kernels inside `pytest -m gpu`,
so that the driver — not only the builder — sees them: install mode B (yastn_backend.activate()) plus the reference's conftest
options ``--backend torch --device cuda``.  tools/run_reference_tests.py does the run in a subprocess (its own pytest session);
here its JSON summary is checked: no failures, and the hot calls really went through the native kernels.

fuse_to_matrix (the policy this backend recommends) runs the whole list including the CTMRG tests (tests/peps/test_ctmrg.py,
SURVEY 2.2 #31); the other two policies run the contraction tests only, to keep the driver's GPU session short."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "baseline", "_ref", "ref_tests")

pytestmark = pytest.mark.gpu

FULL = ["tensor/test_tensordot.py", "tensor/test_ncon_einsum.py", "tensor/test_fuse_hard.py", "tensor/test_transpose.py",
        "tensor/test_tensordot_ad.py", "tensor/test_fuse_meta.py", "tensor/test_cache.py", "tensor/test_vdot.py",
        "tensor/test_svd.py", "tensor/test_qr.py", "tensor/test_eigh.py", "tensor/test_trace.py", "tensor/test_broadcast.py",
        "tensor/test_mask.py", "tensor/test_algebra.py", "tensor/test_swap_gate.py",
        "mps/test_dmrg.py", "mps/test_env.py", "peps/test_ctmrg.py"]
CONTRACTIONS = ["tensor/test_tensordot.py", "tensor/test_ncon_einsum.py", "tensor/test_fuse_hard.py", "tensor/test_vdot.py"]


def _run(policy, files, extra=()):
    if not os.path.isdir(REF_TESTS):
        pytest.skip("baseline/_ref/ref_tests missing (tools/install_reference.sh was not run in the authoring container)")
    files = [f for f in files if os.path.exists(os.path.join(REF_TESTS, f))]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_tests.py"), "--policies", policy, *extra, "--files"] + files,
                       capture_output=True, text=True, timeout=2400)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, r.stdout[-3000:] + r.stderr[-3000:]
    return json.loads(lines[-1]), files


def test_reference_tests_fuse_to_matrix_including_ctmrg():
    d, files = _run("fuse_to_matrix", FULL)
    assert d["failed"] == 0, d["failed_ids"]
    assert d["passed"] >= 100 and d["device"] == "cuda"
    native = d["hot_calls"]["native"]
    assert native["dot"] > 10000 and native["transpose_and_merge"] > 10000 and native["unmerge"] > 1000
    assert "peps/test_ctmrg.py" in files


@pytest.mark.parametrize("policy", ["fuse_contracted", "no_fusion"])
def test_reference_contraction_tests_other_policies(policy):
    d, _ = _run(policy, CONTRACTIONS)
    assert d["failed"] == 0, d["failed_ids"]
    assert d["passed"] >= 40
    native = d["hot_calls"]["native"]
    assert (native["transpose_dot_sum"] if policy == "no_fusion" else native["dot"]) > 100


def test_reference_mps_tests_with_recorded_chains():
    """The reference's DMRG / TDVP / environment / CTMRG tests with Heff1, Heff2, the environment updates and the double-layer
    PEPS contractions replayed from recorded chains (yastn_b200.chain) and dot + unmerge fused: no failures, and the chains
    really replay."""
    d, _ = _run("fuse_to_matrix", ["mps/test_dmrg.py", "mps/test_tdvp.py", "mps/test_env.py", "mps/test_environment.py", "mps/test_measurement.py",
                                   "peps/test_ctmrg.py"], extra=("--chains",))
    assert d["failed"] == 0, d["failed_ids"]
    assert d["passed"] >= 10
    assert d["chains"]["replayed"] > d["chains"]["recorded"] > 0
