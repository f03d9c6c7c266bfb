"""Tensordot chains (yastn_b200.chain): Heff2 / Heff1 / environment updates of the reference's MPS environment
(yastn/tn/mps/_env.py:496-518) recorded once and replayed by one yb_chain_run.

  * ``shim`` (CPU): recording, data-flow analysis, arena layout and the step table, executed by the numpy table interpreter;
  * ``cuda`` (gpu): the same through the C ABI on the B200, plus the launch count of a replayed Heff2.
A replay must give the bits of the unchained call (same plans, same launch order)."""
import ctypes

import numpy as np
import pytest
import torch

from yastn_loader import load_yastn

yastn = load_yastn()
if yastn is None:
    pytest.skip("yastn not importable (no baseline/_ref, no reference checkout)", allow_module_level=True)

import yastn.tn.mps as mps  # noqa: E402
from yastn_b200 import yastn_backend, chain, _lib  # noqa: E402
from yastn_b200 import backend_b200 as bk  # noqa: E402
import cpu_shim  # noqa: E402


@pytest.fixture(params=["shim", pytest.param("cuda", marks=pytest.mark.gpu)])
def device(request):
    if request.param == "shim":
        cpu_shim.install()
        yield "cpu"
        cpu_shim.uninstall()
    else:
        cpu_shim.uninstall()
        assert torch.cuda.is_available()
        yield "cuda"
    chain.disable()
    yastn_backend.disable_fused_tensordot()


def _model(device, sym, dtype, N):
    kw = dict(backend=yastn_backend.module(), default_device=device, tensordot_policy="fuse_to_matrix", default_dtype=dtype)
    if sym == "U1":
        ops = yastn.operators.Spin12(sym="U1", **kw)
        terms = []
        for n in range(N - 1):
            terms += [mps.Hterm(1.0, [n, n + 1], [ops.sz(), ops.sz()]), mps.Hterm(0.5, [n, n + 1], [ops.sp(), ops.sm()]),
                      mps.Hterm(0.5, [n, n + 1], [ops.sm(), ops.sp()])]
        n_total = 0
    else:
        ops = yastn.operators.SpinfulFermions(sym="U1xU1", **kw)
        terms = []
        for n in range(N - 1):
            for s in ("u", "d"):
                terms += [mps.Hterm(-1.0, [n, n + 1], [ops.cp(s), ops.c(s)]), mps.Hterm(-1.0, [n + 1, n], [ops.cp(s), ops.c(s)])]
        for n in range(N):
            terms.append(mps.Hterm(4.0, [n], [ops.n("u") @ ops.n("d")]))
        n_total = (N // 2, N // 2)
    I = mps.product_mpo(ops.I(), N)
    H = mps.generate_mpo(I, terms)
    ops.random_seed(seed=0)
    psi = mps.random_mps(I, n=n_total, D_total=12, dtype=dtype)
    return psi, H


def _bits(t):
    d = t._data
    return d.resolve_conj().cpu().numpy().copy() if d.is_conj() else d.cpu().numpy().copy()


@pytest.mark.parametrize("sym,dtype", [("U1", "float64"), ("U1xU1", "complex128")])
@pytest.mark.parametrize("fused", [False, True])
def test_heff_and_env_updates_replay_bit_exact(device, sym, dtype, fused):
    psi, H = _model(device, sym, dtype, 6)
    psi.canonize_(to="first")
    env = mps.Env(psi, [H, psi])
    env.setup_(to="first")
    if fused:
        yastn_backend.enable_fused_tensordot()
    bd = (2, 3)
    env.update_env_(0, to="last")
    env.update_env_(1, to="last")
    AA = psi.pre_2site(bd)
    ref = {"Heff2": env.Heff2(AA, bd), "Heff1": env.Heff1(psi.A[2], 2),
           "to_last": env.update_env_to_last(env.F[1, 2], 2), "to_first": env.update_env_to_first(env.F[4, 3], 3)}
    chain.enable()
    chain.clear()
    s0 = chain.stats()
    for rep in range(3):            # first pass records, the others replay
        out = {"Heff2": env.Heff2(AA, bd), "Heff1": env.Heff1(psi.A[2], 2),
               "to_last": env.update_env_to_last(env.F[1, 2], 2), "to_first": env.update_env_to_first(env.F[4, 3], 3)}
        for k in ref:
            assert out[k].struct == ref[k].struct and out[k].slices == ref[k].slices and out[k].hfs == ref[k].hfs
            assert np.array_equal(_bits(out[k]), _bits(ref[k])), (k, rep)
            assert out[k].is_consistent() and out[k].are_independent(ref[k])
    s1 = chain.stats()
    assert s1["recorded"] - s0["recorded"] == 4 and s1["replayed"] - s0["replayed"] == 8 and s1["rejected"] == s0["rejected"]
    # another operand structure -> a new recording, not a wrong replay
    AA2 = psi.pre_2site((1, 2))
    r2 = env.Heff2(AA2, (1, 2))
    chain.disable()
    assert np.array_equal(_bits(r2), _bits(env.Heff2(AA2, (1, 2))))


def test_replayed_heff2_is_at_most_twelve_launches(device):
    """One replayed Heff2 = at most 12 launches (4 tensordots x (merge, merge, dot+unmerge)), issued by a single library call."""
    psi, H = _model(device, "U1", "float64", 6)
    psi.canonize_(to="first")
    env = mps.Env(psi, [H, psi])
    env.setup_(to="first")
    env.update_env_(0, to="last")
    env.update_env_(1, to="last")
    yastn_backend.enable_fused_tensordot()
    chain.enable()
    chain.clear()
    AA = psi.pre_2site((2, 3))
    env.Heff2(AA, (2, 3))
    c0 = yastn_backend.call_counts()["native"]
    s0 = chain.stats()
    env.Heff2(AA, (2, 3))
    s1 = chain.stats()
    assert yastn_backend.call_counts()["native"] == c0          # no per-tensordot backend call on the replay
    assert s1["replayed"] - s0["replayed"] == 1
    assert 4 <= s1["launches_replayed"] - s0["launches_replayed"] <= 12


def test_dmrg_energy_unchanged_by_chains(device):
    energies, stats = [], None
    for use in (False, True):
        psi, H = _model(device, "U1", "float64", 8)
        if use:
            chain.enable()
            chain.clear()
        out = mps.dmrg_(psi, H, method="2site", max_sweeps=3, opts_svd={"tol": 1e-10, "D_total": 16})
        energies.append(float(out.energy))
        if use:
            stats = chain.stats()
            chain.disable()
    assert energies[0] == energies[1]
    assert stats["replayed"] > stats["recorded"] > 0


def test_traces_with_foreign_operations_are_never_replayed(device):
    """A torch operation between the launches breaks the recorded data flow: such a call is run as written, every time."""
    psi, H = _model(device, "U1", "float64", 6)
    a, b = psi.A[2], psi.A[3]

    def body(x, y):
        t = yastn.tensordot(x, y, axes=(2, 0))
        return yastn.tensordot(t * 2.0, y.conj(), axes=((2, 3), (1, 2)))      # `* 2.0` is a plain torch multiply
    r0 = body(a, b)
    s0 = chain.stats()
    for _ in range(3):
        r = chain.trace("foreign", body, (a, b))
        assert np.array_equal(_bits(r), _bits(r0))
    s1 = chain.stats()
    assert s1["rejected"] - s0["rejected"] == 1 and s1["replayed"] == s0["replayed"]
    a.requires_grad_(True)
    chain.trace("grad", lambda x, y: yastn.tensordot(x, y, axes=(2, 0)), (a, b))
    assert chain.stats()["bypassed"] == s1["bypassed"] + 1


def test_operands_sharing_data_are_part_of_the_key(device):
    """<psi|H|psi> passes the same tensor as bra and ket; a later call with two different tensors of the same structure (a
    compression environment) must not replay that recording with the ket read from the bra's buffer."""
    psi, H = _model(device, "U1", "float64", 6)
    psi.canonize_(to="first")
    env = mps.Env(psi, [H, psi.shallow_copy()])          # the copy shares its tensors with psi: bra.A[n] is ket.A[n]
    env.setup_(to="first")
    vecL, A, W = env.F[-1, 0], psi.A[0], H.A[0]
    other = A * 1.0
    other._data.mul_(torch.arange(1, other._data.numel() + 1, dtype=other._data.dtype, device=other._data.device))
    plain_same = env.update_env_to_last(vecL, 0)
    env.ket.A[0] = other
    plain_diff = env.update_env_to_last(vecL, 0)
    assert not np.array_equal(_bits(plain_same), _bits(plain_diff))
    chain.enable()
    chain.clear()
    env.ket.A[0] = A
    for _ in range(2):
        assert np.array_equal(_bits(env.update_env_to_last(vecL, 0)), _bits(plain_same))
    env.ket.A[0] = other
    for _ in range(2):
        assert np.array_equal(_bits(env.update_env_to_last(vecL, 0)), _bits(plain_diff))
    env.ket.A[0] = A
    assert np.array_equal(_bits(env.update_env_to_last(vecL, 0)), _bits(plain_same))
    assert chain.stats()["chains"] == 2


def test_ctmrg_double_layer_contractions_replay_bit_exact(device):
    """chain.enable_peps(): append_vec_* of the CTMRG double-layer contractions (yastn/tn/fpeps/envs/_env_contractions.py:211-365)
    traced as they stand; a CTM run with chains reproduces the environment of the plain run bit for bit."""
    import yastn.tn.fpeps as fpeps

    def run(use):
        cfg = yastn.make_config(sym="U1", backend=yastn_backend.module(), default_device=device, tensordot_policy="fuse_to_matrix")
        cfg.backend.random_seed(0)
        lv = yastn.gaussian_leg(cfg, s=1, n=0, sigma=1.0, D_total=3, method="round")
        lp = yastn.Leg(cfg, s=1, t=(-1, 1), D=(1, 1))
        A = yastn.rand(cfg, legs=[lv.conj(), lv, lv, lv.conj(), lp], n=0)
        A = A / A.norm()
        psi = fpeps.Peps(geometry=fpeps.SquareLattice(dims=(1, 1), boundary="infinite"), tensors={(0, 0): A})
        env = fpeps.EnvCTM(psi, init="eye")
        if use:
            chain.enable_peps()
            chain.clear()
        try:
            env.ctmrg_(opts_svd={"D_total": 8, "tol": 1e-12}, max_sweeps=5, corner_tol=1e-14)
            st = chain.stats()
        finally:
            chain.disable()
        e = env[(0, 0)]
        return [_bits(getattr(e, k)) for k in ("tl", "t", "tr", "r", "br", "b", "bl", "l")], st
    yastn_backend.enable_fused_tensordot()
    s0 = chain.stats()
    plain, _ = run(False)
    chained, st = run(True)
    assert st["replayed"] - s0["replayed"] >= 8 and st["rejected"] == s0["rejected"]
    for x, y in zip(plain, chained):
        assert np.array_equal(x, y)


def test_degenerate_traces_are_rejected_not_replayed(device):
    """A traced function that launches nothing (returns an operand, or a lazy transpose of it), or whose result is empty (no
    matching charges), has no replayable data flow: it is run as written every time and gives the reference's answer."""
    cfg = yastn.make_config(sym="U1", backend=yastn_backend.module(), default_device=device, tensordot_policy="fuse_to_matrix")
    cfg.backend.random_seed(1)
    a = yastn.rand(config=cfg, s=(-1, 1, 1), t=((0, 1), (0, 1), (0, 1)), D=((2, 3), (2, 3), (2, 3)))
    b = yastn.rand(config=cfg, s=(-1, 1), t=((5, 6), (5, 6)), D=((2, 3), (2, 3)))       # shares no charge with the legs of a
    s0 = chain.stats()
    for _ in range(3):
        r = chain.trace("identity", lambda x: x.transpose(axes=(1, 0, 2)), (a,))
        assert r.get_shape() == a.transpose(axes=(1, 0, 2)).get_shape()
        e = chain.trace("empty", lambda x, y: yastn.tensordot(x, y, axes=(2, 0)), (a, b))
        assert e.size == 0
    s1 = chain.stats()
    assert s1["replayed"] == s0["replayed"] and s1["rejected"] - s0["rejected"] == 2


def test_chain_abi_rejects_malformed_steps():
    lib = _lib.load()
    h = ctypes.c_void_p()
    good = np.array([[_lib.YB_CHAIN_COPY, 0, 0x1000, 0, 0, 0, 0, 2, 0, 16]], dtype=np.int64)
    assert lib.yb_chain_create(good.ctypes.data, 1, 3, ctypes.byref(h)) == 0 and lib.yb_chain_steps(h) == 1
    lib.yb_chain_destroy(h)
    for col, val in ((0, 7), (2, 0), (3, 5), (7, -1)):
        bad = good.copy()
        bad[0, col] = val
        assert lib.yb_chain_create(bad.ctypes.data, 1, 3, ctypes.byref(h)) != 0
        assert b"malformed" in lib.yb_last_error()
    gem = good.copy()
    gem[0, 0], gem[0, 5] = _lib.YB_CHAIN_GEMM, 9
    assert lib.yb_chain_create(gem.ctypes.data, 1, 3, ctypes.byref(h)) != 0
