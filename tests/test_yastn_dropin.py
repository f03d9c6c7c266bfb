"""Drop-in tests: the real YASTN drives our backend module; results against YASTN's stock numpy backend.

Block structure (struct, slices, hfs, mfs, trans) must be equal tuple-for-tuple ("bit-exact metadata", north_star),
data within rel. Frobenius error 1e-12.  Two variants of every test:
  * ``shim`` (CPU, not marked gpu): the device side of the C ABI is replaced by the numpy table interpreter
    (tests/cpu_shim.py), so only the host logic is under test — plan tables, caching, promotion, conj, autograd;
  * ``cuda`` (marked gpu): the real kernels on the B200 through the C ABI.
The test bodies mirror the reference's own tests (tests/tensor/test_tensordot.py:29-60 ``tensordot_vs_numpy``,
test_fuse_hard.py, test_transpose.py, test_ncon_einsum.py, tests/mps/test_dmrg.py).
"""
import numpy as np
import pytest
import torch

from yastn_loader import load_yastn

yastn = load_yastn()
if yastn is None:
    pytest.skip("yastn not importable (no baseline/_ref, no reference checkout)", allow_module_level=True)

from yastn_b200 import yastn_backend  # noqa: E402
import cpu_shim  # noqa: E402

TOL = 1e-12
POLICIES = ("fuse_to_matrix", "fuse_contracted", "no_fusion")


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


def cfgs(sym, policy, device, fusion="hard"):
    """(reference config on the numpy backend, our config) with otherwise identical settings."""
    ref = yastn.make_config(sym=sym, backend="np", tensordot_policy=policy, default_fusion=fusion)
    our = yastn.make_config(sym=sym, backend=yastn_backend.module(), tensordot_policy=policy, default_fusion=fusion,
                            default_device=device)
    return ref, our


def mirror(x, cfg):
    """Same tensor (identical bits) on another config."""
    return yastn.Tensor.from_dict(x.to_dict(level=2), config=cfg)


def same_structure(c, r):
    assert c.struct == r.struct
    assert c.slices == r.slices
    assert c.hfs == r.hfs and c.mfs == r.mfs and c.trans == r.trans
    assert c.is_consistent()


def close(c, r, tol=TOL):
    same_structure(c, r)
    x = c.to_numpy() if c.size else np.zeros(0)
    y = r.to_numpy() if r.size else np.zeros(0)
    nrm = np.linalg.norm(y)
    assert np.linalg.norm(x - y) <= tol * max(nrm, 1e-300) or (nrm == 0 and np.linalg.norm(x) == 0)


def u1_operands(cfg, dtype):
    a = yastn.rand(config=cfg, s=(-1, 1, 1, -1), t=((-1, 1, 2), (-1, 1, 2), (-1, 1, 2), (-1, 1, 2)),
                   D=((1, 2, 3), (4, 5, 6), (7, 8, 9), (10, 11, 12)), dtype=dtype)
    b = yastn.rand(config=cfg, s=(1, -1, 1), t=((-1, 1, 2), (-1, 1, 2), (-1, 0, 1)),
                   D=((1, 2, 3), (4, 5, 6), (10, 7, 11)), dtype=dtype)
    return a, b


@pytest.mark.parametrize("policy", POLICIES)
@pytest.mark.parametrize("dtype", ["float64", "complex128"])
def test_tensordot_u1(device, policy, dtype):
    ref, our = cfgs("U1", policy, device)
    ref.backend.random_seed(3)
    a, b = u1_operands(ref, dtype)
    A, B = mirror(a, our), mirror(b, our)
    before = yastn_backend.call_counts()["native"]["dot"] + yastn_backend.call_counts()["native"]["transpose_dot_sum"]
    for axes, conj in ((((0, 1), (0, 1)), (0, 0)), ((0, 0), (0, 0)), (((1, 0), (1, 0)), (1, 1)), (((), ()), (0, 0)),
                       (((0, 1), (0, 1)), (1, 0)), (((0, 1), (0, 1)), (0, 1))):
        if conj in ((1, 0), (0, 1)):
            # signatures must match after conjugating one side: contract a with conj(a)-like partner
            r = yastn.tensordot(a, a, axes=((0, 1, 2), (0, 1, 2)), conj=conj)
            c = yastn.tensordot(A, A, axes=((0, 1, 2), (0, 1, 2)), conj=conj)
        else:
            r = yastn.tensordot(a, b, axes=axes, conj=conj)
            c = yastn.tensordot(A, B, axes=axes, conj=conj)
        close(c, r)
        assert yastn.are_independent(c, A) and yastn.are_independent(c, B)
    # lazy transposes folded into the merge order
    r = yastn.tensordot(a.transpose((3, 1, 0, 2)), b.transpose((2, 0, 1)), axes=((2, 1), (1, 2)))
    c = yastn.tensordot(A.transpose((3, 1, 0, 2)), B.transpose((2, 0, 1)), axes=((2, 1), (1, 2)))
    close(c, r)
    after = yastn_backend.call_counts()["native"]["dot"] + yastn_backend.call_counts()["native"]["transpose_dot_sum"]
    assert after - before >= 7      # every contraction went through our functions


@pytest.mark.parametrize("policy", POLICIES)
@pytest.mark.parametrize("sym", ["dense", "Z2", "U1xU1", "Z2xU1"])
def test_tensordot_other_symmetries(device, policy, sym):
    symarg = {"Z2xU1": yastn.sym.sym_Z2xU1, "dense": "dense"}.get(sym, sym)
    ref, our = cfgs(symarg, policy, device)
    ref.backend.random_seed(5)
    if sym == "dense":
        a = yastn.rand(config=ref, s=(-1, 1, 1, -1), D=(2, 3, 4, 5))
        b = yastn.rand(config=ref, s=(1, -1, 1), D=(2, 3, 5), dtype="complex128")
        pairs = [(((0, 3), (0, 2)), (0, 0)), (((), ()), (0, 0))]
    elif sym == "Z2":
        L = yastn.Leg(ref, s=1, t=(0, 1), D=(7, 9)); p = yastn.Leg(ref, s=1, t=(0, 1), D=(1, 1))
        a = yastn.rand(ref, legs=[L.conj(), p, p, L], n=0)
        b = yastn.rand(ref, legs=[L.conj(), p, L], n=1)
        pairs = [((3, 0), (0, 0)), (((0, 1), (0, 1)), (0, 1))]
    elif sym == "U1xU1":
        L = yastn.gaussian_leg(ref, s=1, n=(0, 0), sigma=1.0, D_total=24, method="round")
        p = yastn.Leg(ref, s=1, t=((0, 0), (1, 0), (0, 1), (1, 1)), D=(1, 1, 1, 1))
        a = yastn.rand(ref, legs=[L.conj(), p, L], n=(0, 0), dtype="complex128")
        b = yastn.rand(ref, legs=[L.conj(), p.conj(), p, L], n=(0, 0), dtype="complex128")
        pairs = [((2, 0), (0, 0)), (((1, 2), (1, 0)), (0, 0))]
    else:
        t1 = ((0, -1), (0, 1), (1, -1), (1, 1))
        a = yastn.rand(config=ref, s=(-1, 1, 1, -1), t=(t1, t1, t1, t1), D=((1, 2, 2, 4), (9, 4, 3, 2), (5, 6, 7, 8), (7, 8, 9, 10)))
        b = yastn.rand(config=ref, s=(1, -1, 1), t=(t1, t1, t1), D=((1, 2, 2, 4), (9, 4, 3, 2), (4, 5, 6, 3)))
        pairs = [(((0, 1), (0, 1)), (0, 0)), ((0, 0), (0, 0))]
    A, B = mirror(a, our), mirror(b, our)
    for axes, conj in pairs:
        close(yastn.tensordot(A, B, axes=axes, conj=conj), yastn.tensordot(a, b, axes=axes, conj=conj))


@pytest.mark.parametrize("policy", POLICIES)
def test_missing_charges_and_empty_result(device, policy):
    """Blocks without a partner are filtered (_common_inds), merges pad with zeros, empty result has size 0
    (reference: tests/tensor/test_tensordot.py:140-170)."""
    ref, our = cfgs("U1", policy, device)
    ref.backend.random_seed(7)
    a = yastn.rand(config=ref, s=(-1, 1, 1), t=((-2, 0, 2), (-1, 1), (-3, -1, 1, 3)), D=((2, 3, 4), (5, 6), (1, 2, 3, 4)))
    b = yastn.rand(config=ref, s=(-1, 1, 1), t=((-3, -1, 5), (0, 1), (-1, 0, 1)), D=((1, 2, 7), (3, 2), (2, 3, 4)))
    A, B = mirror(a, our), mirror(b, our)
    close(yastn.tensordot(A, B, axes=(2, 0)), yastn.tensordot(a, b, axes=(2, 0)))
    close(yastn.tensordot(B, A, axes=(0, 2)), yastn.tensordot(b, a, axes=(0, 2)))
    a3 = yastn.rand(config=ref, s=(-1, 1), t=((0,), (0,)), D=((2,), (3,)))
    b3 = yastn.rand(config=ref, s=(-1, 1), t=((1,), (1,)), D=((4,), (5,)))
    c = yastn.tensordot(mirror(a3, our), mirror(b3, our), axes=(1, 0))
    r = yastn.tensordot(a3, b3, axes=(1, 0))
    same_structure(c, r)
    assert c.size == 0


def test_fuse_unfuse_transpose(device):
    """fuse_legs(mode='hard') = transpose_and_merge with an N-d target, unfuse_legs = unmerge, consume_transpose =
    transpose (reference: yastn/tensor/_merging.py:283-301,466-468, _single.py:316-346).  Pure data movement: exact."""
    ref, our = cfgs("U1", "fuse_to_matrix", device)
    ref.backend.random_seed(9)
    a = yastn.rand(config=ref, s=(-1, 1, 1, -1, 1), t=((-1, 0, 1), (0, 1), (-1, 1), (0, 1, 2), (-1, 0)),
                   D=((2, 3, 2), (3, 2), (2, 4), (1, 2, 3), (2, 2)), dtype="complex128")
    A = mirror(a, our)
    f, F = a.fuse_legs(axes=((0, 2), 1, (4, 3)), mode="hard"), A.fuse_legs(axes=((0, 2), 1, (4, 3)), mode="hard")
    close(F, f, tol=0)
    close(F.unfuse_legs(axes=(0, 2)), f.unfuse_legs(axes=(0, 2)), tol=0)
    f2, F2 = f.fuse_legs(axes=((0, 1), 2), mode="hard"), F.fuse_legs(axes=((0, 1), 2), mode="hard")
    close(F2, f2, tol=0)
    close(F2.unfuse_legs(axes=0).unfuse_legs(axes=(0, 2)), f2.unfuse_legs(axes=0).unfuse_legs(axes=(0, 2)), tol=0)
    close(A.transpose((4, 2, 0, 3, 1)).consume_transpose(), a.transpose((4, 2, 0, 3, 1)).consume_transpose(), tol=0)
    close(A.conj().transpose((1, 0, 2, 4, 3)).consume_transpose(), a.conj().transpose((1, 0, 2, 4, 3)).consume_transpose(), tol=0)
    # hard-fused operands contracted over the fused leg
    b = yastn.rand(config=ref, s=(1, -1, -1), t=((-1, 0, 1), (-1, 1), (0, 1)), D=((2, 3, 2), (2, 4), (3, 2)))
    B = mirror(b, our)
    close(yastn.tensordot(F, B.fuse_legs(axes=((0, 1), 2), mode="hard"), axes=(0, 0)),
          yastn.tensordot(f, b.fuse_legs(axes=((0, 1), 2), mode="hard"), axes=(0, 0)))


@pytest.mark.parametrize("policy", POLICIES)
def test_ncon_and_einsum(device, policy):
    ref, our = cfgs("U1", policy, device)
    ref.backend.random_seed(11)
    a, b = u1_operands(ref, "float64")
    A, B = mirror(a, our), mirror(b, our)
    r = yastn.ncon([a, b, b], [[1, 2, -0, -1], [1, 2, 3], [-2, -3, 3]], conjs=(0, 0, 1))
    c = yastn.ncon([A, B, B], [[1, 2, -0, -1], [1, 2, 3], [-2, -3, 3]], conjs=(0, 0, 1))
    close(c, r)
    r = yastn.einsum("abcd,abe->ecd", a, b)
    c = yastn.einsum("abcd,abe->ecd", A, B)
    close(c, r)
    close(B.transpose((1, 2, 0)) @ A, b.transpose((1, 2, 0)) @ a)     # __matmul__ = tensordot(axes=(ndim-1, 0))


@pytest.mark.parametrize("policy", POLICIES)
def test_autograd_through_tensordot(device, policy):
    """Gradients of |tensordot|^2 against the stock torch backend (reference: tests/tensor/test_tensordot_ad.py:32-60)."""
    stock = yastn.make_config(sym="U1", backend="torch", tensordot_policy=policy, default_dtype="float64")
    _, our = cfgs("U1", policy, device)
    stock.backend.random_seed(13)
    a, b = u1_operands(stock, "complex128")
    A, B = mirror(a, our), mirror(b, our)
    grads = []
    for x, y in ((a, b), (A, B)):
        x.requires_grad_(True); y.requires_grad_(True)
        c = yastn.tensordot(x, y, axes=((0, 1), (0, 1)))
        loss = c.norm() ** 2
        loss.backward()
        grads.append((x._data.grad.detach().cpu().numpy(), y._data.grad.detach().cpu().numpy()))
    for g_ref, g in zip(grads[0], grads[1]):
        assert np.linalg.norm(g - g_ref) <= 1e-11 * np.linalg.norm(g_ref)


def test_backend_module_contract(device):
    import yastn.backend.backend_torch as stock
    mod = yastn_backend.module()
    assert mod is yastn_backend.module()                     # stable identity: _config is an lru key
    assert mod.BACKEND_ID == "torch" and mod.DTYPE is stock.DTYPE
    for name in stock.__all__:
        assert hasattr(mod, name), name
    hash(mod)
    # tensors of the two backends are combinable (same BACKEND_ID), like the reference requires (_tests.py:38-39)
    ref, our = cfgs("U1", "fuse_to_matrix", device)
    if device == "cpu":
        stock_cfg = yastn.make_config(sym="U1", backend="torch", tensordot_policy="fuse_to_matrix")
        a, b = u1_operands(stock_cfg, "float64")
        A = mirror(a, our)
        close(yastn.tensordot(A, b, axes=((0, 1), (0, 1))), yastn.tensordot(a, b, axes=((0, 1), (0, 1))))


def test_activate_rebinds_stock_backend(device):
    import yastn.backend.backend_torch as stock
    orig = stock.dot
    try:
        yastn_backend.activate()
        assert stock.dot is not orig
        cfg = yastn.make_config(sym="U1", backend="torch", default_device=device)
        ref = yastn.make_config(sym="U1", backend="np")
        ref.backend.random_seed(17)
        a, b = u1_operands(ref, "float64")
        n0 = yastn_backend.call_counts()["native"]["dot"]
        close(yastn.tensordot(mirror(a, cfg), mirror(b, cfg), axes=((0, 1), (0, 1))), yastn.tensordot(a, b, axes=((0, 1), (0, 1))))
        assert yastn_backend.call_counts()["native"]["dot"] == n0 + 1
    finally:
        yastn_backend.deactivate()
    assert stock.dot is orig


def test_cpu_tensor_is_rejected_without_shim():
    cpu_shim.uninstall()
    cfg = yastn.make_config(sym="U1", backend=yastn_backend.module(), default_device="cpu")
    a, b = u1_operands(cfg, "float64")
    with pytest.raises(TypeError, match="no CPU path"):
        yastn.tensordot(a, b, axes=((0, 1), (0, 1)))


def test_activate_cpu_passthrough_is_opt_in():
    """After activate() the patched stock module rejects CPU tensors unless cpu_passthrough=True, in which case they go to
    the functions that were bound before activate() (the reference's code) and are counted as delegated."""
    cpu_shim.uninstall()
    try:
        yastn_backend.activate(cpu_passthrough=True)
        cfg = yastn.make_config(sym="U1", backend="torch", default_device="cpu")
        ref = yastn.make_config(sym="U1", backend="np")
        ref.backend.random_seed(5)
        a, b = u1_operands(ref, "float64")
        d0 = yastn_backend.call_counts()["delegated"]["dot"]
        close(yastn.tensordot(mirror(a, cfg), mirror(b, cfg), axes=((0, 1), (0, 1))), yastn.tensordot(a, b, axes=((0, 1), (0, 1))))
        assert yastn_backend.call_counts()["delegated"]["dot"] == d0 + 1
        yastn_backend.deactivate()
        yastn_backend.activate()
        with pytest.raises(TypeError, match="no CPU path"):
            yastn.tensordot(mirror(a, cfg), mirror(b, cfg), axes=((0, 1), (0, 1)))
    finally:
        yastn_backend.deactivate()


def test_transpose_plan_cache_hits(device):
    """consume_transpose rebuilds its meta on every call (yastn/tensor/_single.py:343 is not lru-cached): the transpose plan is
    keyed by content, so repeated transposes of equally structured tensors build ONE plan."""
    ref, our = cfgs("U1", "fuse_to_matrix", device)
    ref.backend.random_seed(11)
    a, _ = u1_operands(ref, "float64")
    A = mirror(a, our)
    our.backend.clear_plan_cache()
    s0 = our.backend.plan_cache_stats()
    for _ in range(20):
        close(A.transpose((2, 0, 3, 1)).consume_transpose(), a.transpose((2, 0, 3, 1)).consume_transpose())
    s1 = our.backend.plan_cache_stats()
    assert s1["misses"] - s0["misses"] == 1 and s1["hits"] - s0["hits"] == 19 and s1["size"] == 1


def test_clear_cache_also_drops_plans(device):
    """yastn.clear_cache() (yastn/tensor/_control_lru.py:44-63) invalidates every meta the plans are keyed on."""
    ref, our = cfgs("U1", "fuse_to_matrix", device)
    ref.backend.random_seed(12)
    a, b = u1_operands(ref, "float64")
    A, B = mirror(a, our), mirror(b, our)
    yastn.tensordot(A, B, axes=((0, 1), (0, 1)))
    assert our.backend.plan_cache_stats()["size"] > 0
    yastn.clear_cache()
    assert our.backend.plan_cache_stats()["size"] == 0
    close(yastn.tensordot(A, B, axes=((0, 1), (0, 1))), yastn.tensordot(a, b, axes=((0, 1), (0, 1))))
    yastn.set_cache_maxsize(maxsize=1024)
    assert our.backend.plan_cache_stats()["size"] == 0


@pytest.mark.parametrize("dtype", ["float64", "complex128"])
def test_vector_ops_run_on_the_elementwise_engine(device, dtype):
    """SURVEY 8f rows 2-3 through the real YASTN: add / sub / linear combinations, broadcast (dot_diag), apply_mask, trace,
    svd_with_truncation (apply_mask on U, S, V) against the numpy backend; the calls are counted as native."""
    ref, our = cfgs("U1", "fuse_to_matrix", device)
    ref.backend.random_seed(21)
    leg = yastn.Leg(ref, s=1, t=(-1, 0, 1), D=(4, 5, 6))
    l2 = yastn.Leg(ref, s=1, t=(-1, 0, 2), D=(2, 3, 4))
    a = yastn.rand(ref, legs=[leg.conj(), l2, leg, l2.conj()], dtype=dtype)
    b = yastn.rand(ref, legs=[leg.conj(), l2, leg, l2.conj()], dtype=dtype)
    c = yastn.Tensor(config=ref, s=a.struct.s, dtype=dtype)
    c.set_block(ts=(0, 0, 0, 0), Ds=(5, 3, 5, 3), val='rand')
    c.set_block(ts=(1, -1, 1, -1), Ds=(6, 2, 6, 2), val='rand')
    d = yastn.rand(ref, legs=[leg.conj(), leg], isdiag=True, dtype=dtype)
    m = yastn.rand(ref, legs=[leg.conj(), leg], isdiag=True, dtype="float64") > 0.
    A, B, C, D, M = (mirror(x, our) for x in (a, b, c, d, m))
    n0 = dict(yastn_backend.call_counts()["native"])
    close(A + B, a + b)
    close(A - C, a - c)
    close(yastn.add(A, B, C, A, B, amplitudes=(1, -2, 0.5, None, 3)), yastn.add(a, b, c, a, b, amplitudes=(1, -2, 0.5, None, 3)), tol=1e-13)
    close(yastn.broadcast(D, A, axes=2), yastn.broadcast(d, a, axes=2), tol=1e-13)
    close(yastn.apply_mask(M, A, axes=0), yastn.apply_mask(m, a, axes=0))
    close(yastn.trace(A, axes=(0, 2)), yastn.trace(a, axes=(0, 2)), tol=1e-13)
    close(yastn.trace(A, axes=((0, 1), (2, 3))), yastn.trace(a, axes=((0, 1), (2, 3))), tol=1e-13)
    n1 = yastn_backend.call_counts()["native"]
    for name in ("add", "sub", "dot_diag", "apply_mask", "trace"):
        assert n1[name] > n0[name], name
    U, S, V = yastn.svd_with_truncation(A, axes=((0, 1), (2, 3)), D_total=9)
    u, s, v = yastn.svd_with_truncation(a, axes=((0, 1), (2, 3)), D_total=9)
    close(S, s, tol=1e-12)
    close(U @ S @ V, u @ s @ v, tol=1e-11)


def test_flip_charges_mask_given_as_slice(device):
    """yastn.flip_charges permutes whole blocks through embed_mask with ``mask = {0: slice(None)}`` (yastn/tensor/_single.py:224):
    a mask may be a slice, not only an index vector (found by the reference's test_fuse_hard.py::test_initialize_eye)."""
    ref, our = cfgs("U1", "fuse_to_matrix", device)
    ref.backend.random_seed(5)
    a, _ = u1_operands(ref, "float64")
    A = mirror(a, our)
    n0 = yastn_backend.call_counts()["native"]["embed_mask"]
    for axes in ((1,), (0, 2), (1, 2, 3)):
        close(A.flip_charges(axes=axes), a.flip_charges(axes=axes))
    assert yastn_backend.call_counts()["native"]["embed_mask"] - n0 == 3


def test_swap_gate_negates_blocks_natively(device):
    """Fermionic swap gates (negate_blocks; every CTM move of a fermionic PEPS): one launch, bit-exact."""
    ref = yastn.make_config(sym="Z2", backend="np", fermionic=True)
    our = yastn.make_config(sym="Z2", backend=yastn_backend.module(), fermionic=True, default_device=device)
    ref.backend.random_seed(4)
    l1 = yastn.Leg(ref, s=1, t=(0, 1), D=(3, 4))
    l2 = yastn.Leg(ref, s=1, t=(0, 1), D=(2, 5))
    a = yastn.rand(ref, legs=[l1.conj(), l2, l1, l2.conj()], n=1, dtype="float64")
    A = mirror(a, our)
    n0 = yastn_backend.call_counts()["native"]["negate_blocks"]
    for axes in ((0, 1), ((0, 1), (2, 3)), (0, 2, 1, 3)):
        x, y = A.swap_gate(axes=axes), a.swap_gate(axes=axes)
        same_structure(x, y)
        assert np.array_equal(x.to_numpy(), y.to_numpy())
    assert yastn_backend.call_counts()["native"]["negate_blocks"] == n0 + 3


def test_dmrg_heisenberg_small(device):
    """Config 1 (reduced): U(1) Heisenberg chain 2-site DMRG runs unmodified on our backend and reproduces the
    energy of the reference numpy backend (reference: tests/mps/test_dmrg.py, yastn/tn/mps/_dmrg.py:42-249)."""
    import yastn.tn.mps as mps
    energies = []
    N, D = (8, 16)
    for which in ("ref", "our"):
        ref, our = cfgs("U1", "fuse_to_matrix", device)
        cfg = ref if which == "ref" else our
        ops = yastn.operators.Spin12(sym="U1", backend=cfg.backend, default_device=cfg.default_device, tensordot_policy="fuse_to_matrix")
        ops.random_seed(seed=0)
        I = mps.product_mpo(ops.I(), N)
        terms = []
        for n in range(N - 1):
            terms.append(mps.Hterm(1.0, [n, n + 1], [ops.sz(), ops.sz()]))
            terms.append(mps.Hterm(0.5, [n, n + 1], [ops.sp(), ops.sm()]))
            terms.append(mps.Hterm(0.5, [n, n + 1], [ops.sm(), ops.sp()]))
        H = mps.generate_mpo(I, terms)
        psi = mps.random_mps(I, n=0, D_total=8)
        out = mps.dmrg_(psi, H, method="2site", max_sweeps=6, opts_svd={"tol": 1e-10, "D_total": D}, energy_tol=1e-10)
        energies.append(float(out.energy))
    # N=8 open Heisenberg chain ground state energy (exact diagonalisation): -3.374932598687897
    assert abs(energies[0] - (-3.374932598687897)) < 1e-7
    assert abs(energies[1] - energies[0]) < 1e-8


@pytest.mark.parametrize("dtype", ["float64", "complex128"])
def test_fused_tensordot_path(device, dtype):
    """enable_fused_tensordot(): fuse_to_matrix tensordot = merge, merge, ONE dot+unmerge launch; same tensors as the reference,
    autograd included; backends without dot_unmerge (numpy) are untouched."""
    ref, our = cfgs("U1", "fuse_to_matrix", device)
    ref.backend.random_seed(23)
    a, b = u1_operands(ref, dtype)
    A, B = mirror(a, our), mirror(b, our)
    yastn_backend.enable_fused_tensordot()
    try:
        n0 = yastn_backend.call_counts()["native"]
        for axes in (((0, 1), (0, 1)), (0, 0), ((), ())):
            close(yastn.tensordot(A, B, axes=axes), yastn.tensordot(a, b, axes=axes))
        n1 = yastn_backend.call_counts()["native"]
        assert n1["dot_unmerge"] - n0["dot_unmerge"] >= 2 and n1["unmerge"] == n0["unmerge"]
        A.requires_grad_(True)
        loss = yastn.tensordot(A, B, axes=(0, 0)).norm() ** 2
        loss.backward()
        g_fused = A._data.grad.detach().cpu().numpy().copy()
    finally:
        yastn_backend.disable_fused_tensordot()
    A2 = mirror(a, our)
    A2.requires_grad_(True)
    (yastn.tensordot(A2, B, axes=(0, 0)).norm() ** 2).backward()
    g_plain = A2._data.grad.detach().cpu().numpy()
    assert np.linalg.norm(g_fused - g_plain) <= 1e-12 * np.linalg.norm(g_plain)


@pytest.mark.parametrize("sym", ["dense", "U1", "Z2", "U1xU1"])
def test_kernel_tensordot_bs_boundary(device, sym):
    """The single-call boundary: YASTN's no_fusion tensordot hands raw block tables to ``backend.kernel_tensordot_bs``
    when BACKEND_ID == 'torch_cpp' (yastn/tensor/_contractions.py:199-242; reference implementation
    backend_torch_cpp.py:173-228 = cuTENSOR).  Structure bit-exact, data 1e-12 against the numpy backend; autograd too."""
    ref = yastn.make_config(sym=sym, backend="np", tensordot_policy="no_fusion")
    our = yastn.make_config(sym=sym, backend=yastn_backend.module(bs_boundary=True), tensordot_policy="no_fusion", default_device=device)
    assert our.backend.BACKEND_ID == "torch_cpp" and yastn_backend.module().BACKEND_ID == "torch"
    ref.backend.random_seed(29)
    if sym == "dense":
        a = yastn.rand(config=ref, s=(-1, 1, 1, -1), D=(2, 3, 4, 5))
        b = yastn.rand(config=ref, s=(1, -1, 1), D=(2, 3, 5), dtype="complex128")
        cases = [((0, 3), (0, 2)), ((1,), (1,))]
    elif sym == "U1":
        a, b = u1_operands(ref, "complex128")
        cases = [((0, 1), (0, 1)), ((0,), (0,)), ((1, 0), (1, 0))]
    elif sym == "Z2":
        L = yastn.Leg(ref, s=1, t=(0, 1), D=(7, 9)); p = yastn.Leg(ref, s=1, t=(0, 1), D=(1, 1))
        a = yastn.rand(ref, legs=[L.conj(), p, p, L], n=0)
        b = yastn.rand(ref, legs=[L.conj(), p, L], n=1)
        cases = [((3,), (0,))]
    else:
        L = yastn.gaussian_leg(ref, s=1, n=(0, 0), sigma=1.0, D_total=24, method="round")
        p = yastn.Leg(ref, s=1, t=((0, 0), (1, 0), (0, 1), (1, 1)), D=(1, 1, 1, 1))
        a = yastn.rand(ref, legs=[L.conj(), p, L], n=(0, 0), dtype="complex128")
        b = yastn.rand(ref, legs=[L.conj(), p.conj(), p, L], n=(0, 0), dtype="complex128")
        cases = [((2,), (0,)), ((1, 2), (1, 0))]
    A, B = mirror(a, our), mirror(b, our)
    n0 = yastn_backend.call_counts()["native"]["kernel_tensordot_bs"]
    for axes in cases:
        close(yastn.tensordot(A, B, axes=axes), yastn.tensordot(a, b, axes=axes))
    # lazy transpose + conj of one operand reach the boundary as permuted native axes / a conj bit on the data
    close(yastn.tensordot(A.transpose(tuple(range(A.ndim))[::-1]), B, axes=((A.ndim - 1 - cases[0][0][0],), (cases[0][1][0],))),
          yastn.tensordot(a.transpose(tuple(range(a.ndim))[::-1]), b, axes=((a.ndim - 1 - cases[0][0][0],), (cases[0][1][0],))))
    close(yastn.tensordot(A, A, axes=((0, 1), (0, 1)), conj=(0, 1)), yastn.tensordot(a, a, axes=((0, 1), (0, 1)), conj=(0, 1)))
    assert yastn_backend.call_counts()["native"]["kernel_tensordot_bs"] - n0 >= len(cases) + 2
    # gradients through the boundary against the stock torch backend
    stock = yastn.make_config(sym=sym, backend="torch", tensordot_policy="no_fusion")
    grads = []
    if sym == "dense":      # the reference's own no_fusion backward does not promote mixed dtypes: same dtype on both sides
        a = a.to(dtype="complex128")
    for cfg in (stock, our):
        x, y = mirror(a, cfg), mirror(b, cfg)
        x.requires_grad_(True); y.requires_grad_(True)
        (yastn.tensordot(x, y, axes=cases[0]).norm() ** 2).backward()
        grads.append((x._data.grad.detach().cpu().numpy(), y._data.grad.detach().cpu().numpy()))
    for g_ref, g in zip(grads[0], grads[1]):
        assert np.linalg.norm(g - g_ref) <= 1e-11 * np.linalg.norm(g_ref)


@pytest.mark.parametrize("sym", ["U1", "U1xU1"])
def test_bs_to_tds_matches_reference_meta(sym):
    """plans.bs_to_tds (the meta pass behind kernel_tensordot_bs) against the reference's own _meta_tensordot_nf
    (yastn/tensor/_contractions.py:349-450) on the block tables YASTN would hand over: same result-block slices and
    shapes, same reshape records, same set of block pairs per result block."""
    from yastn.tensor._contractions import _meta_tensordot_nf, _common_inds
    from yastn_b200 import plans
    cfg = yastn.make_config(sym=sym, backend="np")
    cfg.backend.random_seed(31)
    if sym == "U1":
        a, b = u1_operands(cfg, "float64")
        nout_a, nin_a, nin_b, nout_b = (2, 3), (0, 1), (0, 1), (2,)
    else:
        L = yastn.gaussian_leg(cfg, s=1, n=(0, 0), sigma=1.5, D_total=40, method="round")
        p = yastn.Leg(cfg, s=1, t=((0, 0), (1, 0), (0, 1), (1, 1)), D=(1, 2, 1, 1))
        a = yastn.rand(cfg, legs=[L.conj(), p, L], n=(0, 0))
        b = yastn.rand(cfg, legs=[L.conj(), p.conj(), p, L], n=(1, 0))
        nout_a, nin_a, nin_b, nout_b = (1, 0), (2,), (0,), (3, 1, 2)
    nsym = cfg.sym.NSYM
    ind_a, ind_b = _common_inds(a.struct.t, b.struct.t, nin_a, nin_b, a.ndim_n, b.ndim_n, nsym)
    meta_dot, reshape_a, reshape_b, struct_c, slices_c, _, _ = _meta_tensordot_nf(a.struct, a.slices, b.struct, b.slices, ind_a, ind_b,
                                                                                  nout_a, nin_a, nin_b, nout_b)
    at = a.struct.t if ind_a is None else tuple(a.struct.t[i] for i in ind_a)
    asl = a.slices if ind_a is None else tuple(a.slices[i] for i in ind_a)
    bt = b.struct.t if ind_b is None else tuple(b.struct.t[i] for i in ind_b)
    bsl = b.slices if ind_b is None else tuple(b.slices[i] for i in ind_b)
    md, ra, rb = plans.bs_to_tds(at, asl, nout_a, nin_a, bt, bsl, nout_b, nin_b, struct_c.t, slices_c)
    assert ra == tuple((sl, tuple(D), int(l), int(r)) for sl, D, l, r in reshape_a)
    assert rb == tuple((sl, tuple(D), int(l), int(r)) for sl, D, l, r in reshape_b)
    assert len(md) == len(meta_dot)
    for (sl, Dlr, pairs), (sl_r, Dlr_r, pairs_r) in zip(md, meta_dot):
        assert sl == tuple(sl_r) and Dlr == tuple(int(x) for x in Dlr_r)
        assert sorted(pairs) == sorted(tuple(p) for p in pairs_r)


@pytest.mark.parametrize("dtype", ["float64", "complex128"])
def test_vdot_is_one_grouped_gemm(device, dtype):
    """yastn.vdot on tensors with different block structures (meta path, yastn/tensor/_contractions.py:590-630): all common
    blocks in one launch, every conj combination, against the numpy backend."""
    ref, our = cfgs("U1", "fuse_to_matrix", device)
    ref.backend.random_seed(11)
    a = yastn.rand(config=ref, s=(-1, 1, 1), t=((-1, 0, 1, 2), (-1, 0, 1), (-2, -1, 0, 1, 2)), D=((2, 3, 4, 5), (3, 2, 4), (1, 5, 6, 2, 3)), dtype=dtype)
    b = yastn.rand(config=ref, s=(-1, 1, 1), t=((-1, 0, 1), (-1, 0, 1), (-1, 0, 1, 2)), D=((2, 3, 4), (3, 2, 4), (5, 6, 2, 3)), dtype=dtype)
    A, B = mirror(a, our), mirror(b, our)
    before = yastn_backend.call_counts()["native"]["vdot"]
    for conj in ((1, 0), (0, 1), (0, 0), (1, 1)):
        x, y = (a, b) if conj[0] != conj[1] else (a, b.conj())      # signatures must match after the conjugations
        X, Y = (A, B) if conj[0] != conj[1] else (A, B.conj())
        r = yastn.vdot(x, y, conj=conj)
        c = yastn.vdot(X, Y, conj=conj)
        assert abs(complex(c) - complex(r)) <= TOL * max(abs(complex(r)), 1.0)
    assert yastn_backend.call_counts()["native"]["vdot"] >= before + 4
    # same structure: the reference takes its single-dot path, no meta
    assert abs(complex(yastn.vdot(A, A)) - complex(yastn.vdot(a, a))) <= TOL * abs(complex(yastn.vdot(a, a)))
