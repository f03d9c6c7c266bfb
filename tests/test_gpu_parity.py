"""GPU parity: the CUDA path, called through the C ABI, against golden vectors recorded from the reference
and against the CPU oracle.  Bit-exact for data movement; rel. Frobenius error <= 1e-12 for GEMMs
(the tolerance BASELINE.json's north_star states for float64 / complex128)."""
import numpy as np
import pytest
import torch

from oracle import backend_oracle as orc
from golden_io import small_calls, bench_structs

pytestmark = pytest.mark.gpu
TOL = 1e-12

CALLS = small_calls()


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _relerr(out, ref):
    return np.linalg.norm(out - ref) / max(np.linalg.norm(ref), 1e-300)


def _ids(calls):
    return [f"{c['case']}-{c['policy'][:6]}-{c['dtype']}-{k}" for k, c in enumerate(calls)]


@pytest.fixture(scope="module")
def bk():
    from yastn_b200 import backend_b200
    return backend_b200


COPY_CALLS = [c for c in CALLS if c["fn"] in ("transpose_and_merge", "unmerge", "transpose")]
GEMM_CALLS = [c for c in CALLS if c["fn"] in ("dot", "transpose_dot_sum")]


@pytest.mark.parametrize("call", COPY_CALLS, ids=_ids(COPY_CALLS))
def test_copy_functions_bit_exact(bk, call):
    a = dict(call["args"])
    data = _dev(a.pop("data"))
    out = getattr(bk, call["fn"])(data, *a.values())
    torch.cuda.synchronize()
    assert out.dtype == data.dtype and out.is_contiguous() and out.data_ptr() != data.data_ptr()
    assert np.array_equal(out.cpu().numpy(), call["out"])


@pytest.mark.parametrize("call", GEMM_CALLS, ids=_ids(GEMM_CALLS))
def test_gemm_functions(bk, call):
    a = dict(call["args"])
    A, B = _dev(a.pop("Adata")), _dev(a.pop("Bdata"))
    out = getattr(bk, call["fn"])(A, B, *a.values())
    torch.cuda.synchronize()
    assert out.cpu().numpy().dtype == call["out"].dtype
    assert _relerr(out.cpu().numpy(), call["out"]) <= TOL


def test_lazy_conj_inputs(bk):
    """torch's conj bit (backend_torch.conj is lazy, backend_torch.py:264) must be honoured by raw-pointer kernels."""
    n = 0
    for call in CALLS:
        if call["dtype"] != "complex128":
            continue
        a = dict(call["args"])
        if call["fn"] in ("transpose_and_merge", "unmerge", "transpose"):
            data = a.pop("data")
            out = getattr(bk, call["fn"])(_dev(data.conj()).conj(), *a.values())   # lazily conjugated twice-stored data
            ref = getattr(orc, call["fn"])(data, *a.values())
            assert np.array_equal(out.cpu().numpy(), ref)
            n += 1
        elif call["fn"] == "dot":
            A, B = a.pop("Adata"), a.pop("Bdata")
            out = bk.dot(_dev(A.conj()).conj(), _dev(B).conj(), *a.values())
            ref = orc.dot(A, B.conj(), *a.values())
            assert _relerr(out.cpu().numpy(), ref) <= TOL
            n += 1
    assert n > 10


def test_backward_matches_oracle_adjoints(bk):
    rng = np.random.default_rng(7)
    seen = set()
    for call in CALLS:
        fn = call["fn"]
        key = (fn, call["dtype"], call["policy"])
        if key in seen:
            continue
        seen.add(key)
        a = dict(call["args"])
        cplx = call["dtype"] == "complex128"

        def rnd(n):
            g = rng.standard_normal(n)
            return g + 1j * rng.standard_normal(n) if cplx else g
        if fn in ("transpose_and_merge", "unmerge", "transpose"):
            data = _dev(a.pop("data")).requires_grad_(True)
            out = getattr(bk, fn)(data, *a.values())
            G = rnd(out.numel())
            out.backward(_dev(G))
            if fn == "transpose_and_merge":
                ref = orc.transpose_and_merge_backward(G, a["order"], a["meta_new"], a["meta_mrg"], data.numel())
            elif fn == "unmerge":
                ref = orc.unmerge_backward(G, a["meta"])
            else:
                ref = orc.transpose_backward(G, a["axes"], a["meta_transpose"])
            assert np.array_equal(data.grad.cpu().numpy(), ref)
        elif fn == "dot":
            A0, B0 = a.pop("Adata"), a.pop("Bdata")
            A, B = _dev(A0).requires_grad_(True), _dev(B0).requires_grad_(True)
            out = bk.dot(A, B, *a.values())
            G = rnd(out.numel())
            out.backward(_dev(G))
            gA, gB = orc.dot_backward(G, A0, B0, a["meta_dot"])
            assert _relerr(A.grad.cpu().numpy(), gA) <= TOL
            assert _relerr(B.grad.cpu().numpy(), gB) <= TOL
        else:
            # transpose_dot_sum: compare with torch autograd through the oracle-equivalent dense formula
            A0, B0 = a.pop("Adata"), a.pop("Bdata")
            A, B = _dev(A0).requires_grad_(True), _dev(B0).requires_grad_(True)
            out = bk.transpose_dot_sum(A, B, *a.values())
            G = rnd(out.numel())
            out.backward(_dev(G))
            At, Bt = torch.from_numpy(A0).requires_grad_(True), torch.from_numpy(B0).requires_grad_(True)
            Am = [At[sl[0]:sl[1]].view(Di).permute(a["Aorder"]).reshape(Dl, Dr) for sl, Di, Dl, Dr in a["Areshape"]]
            Bm = [Bt[sl[0]:sl[1]].view(Di).permute(a["Border"]).reshape(Dl, Dr) for sl, Di, Dl, Dr in a["Breshape"]]
            loss = 0
            Gt = torch.from_numpy(G)
            for sl, Dslc, pairs in a["meta_dot"]:
                blk = sum(Am[ia] @ Bm[ib] for ia, ib in pairs)
                g = Gt[sl[0]:sl[1]].view(Dslc)
                loss = loss + (blk * g.conj()).sum().real if cplx else loss + (blk * g).sum()
            loss.backward()
            assert _relerr(A.grad.cpu().numpy(), At.grad.numpy()) <= TOL
            assert _relerr(B.grad.cpu().numpy(), Bt.grad.numpy()) <= TOL
    assert len(seen) >= 10


def _run_f2m(bk, A, B, case):
    st = case["f2m"]
    ma, mb = st["merge_a"], st["merge_b"]
    Am = A if ma is None else bk.transpose_and_merge(A, ma["order"], ma["meta_new"], ma["meta_mrg"], ma["Dsize"])
    Bm = B if mb is None else bk.transpose_and_merge(B, mb["order"], mb["meta_new"], mb["meta_mrg"], mb["Dsize"])
    C = bk.dot(Am, Bm, st["dot"]["meta_dot"], st["dot"]["Dsize"])
    if st["unmerge"] is not None:
        C = bk.unmerge(C, st["unmerge"]["meta"])
    return C


@pytest.mark.parametrize("name", ["U1_D64_P1", "U1_D64_P2", "U1_D64_P3", "U1_D1024_P1", "U1_D1024_P2", "U1_D1024_P3",
                                  "Z2_D512_P1", "Z2_D512_P2", "U1_D2048_P2", "U1xU1_D4096_P1", "U1xU1_D4096_P2", "U1_D4096_T1"])
@pytest.mark.parametrize("dtype", ["float64", "complex128"])
def test_tensordot_pipeline_vs_oracle(bk, name, dtype):
    """merge -> dot -> unmerge on benchmark-shaped structures (reference metas) against the CPU oracle."""
    case = bench_structs()[name]
    rng = np.random.default_rng(2)
    A = rng.uniform(-1, 1, case["a"]["size"]); B = rng.uniform(-1, 1, case["b"]["size"])
    if dtype == "complex128":
        A = A + 1j * rng.uniform(-1, 1, A.size); B = B + 1j * rng.uniform(-1, 1, B.size)
    ref = orc.tensordot_f2m(A, B, case)
    out = _run_f2m(bk, _dev(A), _dev(B), case)
    torch.cuda.synchronize()
    assert out.numel() == case["f2m"]["struct_c"]["size"]
    assert _relerr(out.cpu().numpy(), ref) <= TOL


@pytest.mark.parametrize("name", ["U1_D1024_P1", "U1_D1024_P2", "Z2_D512_P1", "U1_D64_P3"])
def test_policies_agree_on_gpu(bk, name):
    """fuse_contracted and no_fusion metas recorded from the reference give the same tensor as fuse_to_matrix."""
    case = bench_structs()[name]
    rng = np.random.default_rng(4)
    A = _dev(rng.uniform(-1, 1, case["a"]["size"])); B = _dev(rng.uniform(-1, 1, case["b"]["size"]))
    ref = _run_f2m(bk, A, B, case).cpu().numpy()
    st = case["fc"]
    ma, mb = st["merge_a"], st["merge_b"]
    Am = A if ma is None else bk.transpose_and_merge(A, ma["order"], ma["meta_new"], ma["meta_mrg"], ma["Dsize"])
    Bm = B if mb is None else bk.transpose_and_merge(B, mb["order"], mb["meta_new"], mb["meta_mrg"], mb["Dsize"])
    out = bk.dot(Am, Bm, st["dot"]["meta_dot"], st["dot"]["Dsize"]).cpu().numpy()
    assert _relerr(out, ref) <= TOL
    t = case["nf"]["tds"]
    out = bk.transpose_dot_sum(A, B, t["meta_dot"], t["Areshape"], t["Breshape"], t["Aorder"], t["Border"], t["Dsize"]).cpu().numpy()
    assert _relerr(out, ref) <= TOL


@pytest.mark.parametrize("dtype", ["float64", "complex128"])
@pytest.mark.parametrize("name", ["U1_D4096_T1", "U1_D4096_P3", "U1xU1_D4096_P1"])
def test_large_merges_bit_exact(bk, name, dtype):
    """Both merges of benchmark-sized contractions against the oracle, bit for bit: T1 transposes every (Dl x Dr) block
    (tiled path with ragged edge slabs), P3 interleaves two source blocks element-wise in the destination, U1xU1 is made
    of thousands of small blocks (pack path)."""
    case = bench_structs()[name]
    rng = np.random.default_rng(4)
    for key, n in (("merge_a", case["a"]["size"]), ("merge_b", case["b"]["size"])):
        m = case["f2m"][key]
        if m is None:
            continue
        X = rng.uniform(-1, 1, n)
        if dtype == "complex128":
            X = X + 1j * rng.uniform(-1, 1, n)
        ref = orc.transpose_and_merge(X, m["order"], m["meta_new"], m["meta_mrg"], m["Dsize"])
        got = bk.transpose_and_merge(_dev(X), m["order"], m["meta_new"], m["meta_mrg"], m["Dsize"])
        assert np.array_equal(got.cpu().numpy(), ref), (name, key)


@pytest.mark.parametrize("name", ["U1_D8192_P1", "U1_D16384_P2", "U1_D16384_T1"])
def test_full_size_linearity_and_roundtrip(bk, name):
    """Full benchmark sizes (oracle too slow): size-independent properties.
    (1) merge followed by its adjoint restores every block that took part; (2) the contraction is linear in A."""
    case = bench_structs()[name]
    gen = torch.Generator(device="cuda").manual_seed(0)
    A = torch.rand(case["a"]["size"], dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    A2 = torch.rand(case["a"]["size"], dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    B = torch.rand(case["b"]["size"], dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    ma = case["f2m"]["merge_a"]
    if ma is not None:
        x = A.clone().requires_grad_(True)
        y = bk.transpose_and_merge(x, ma["order"], ma["meta_new"], ma["meta_mrg"], ma["Dsize"])
        y.backward(y.detach())
        assert torch.equal(x.grad, A)          # all blocks take part here: adjoint(merge(A)) == A exactly
        assert torch.equal(y.sum(), y.sum()) and abs(float(y.detach().abs().sum() - A.abs().sum())) <= 1e-6 * float(A.abs().sum())
    C1 = _run_f2m(bk, A, B, case)
    C2 = _run_f2m(bk, A2, B, case)
    C12 = _run_f2m(bk, A + 0.5 * A2, B, case)
    err = float(torch.linalg.norm(C12 - (C1 + 0.5 * C2)) / torch.linalg.norm(C12))
    assert err <= 1e-12
    # spot-check a few sectors against torch.matmul on the merged operands
    st = case["f2m"]
    Am = A if st["merge_a"] is None else bk.transpose_and_merge(A, st["merge_a"]["order"], st["merge_a"]["meta_new"], st["merge_a"]["meta_mrg"], st["merge_a"]["Dsize"])
    mb = st["merge_b"]
    Bm = B if mb is None else bk.transpose_and_merge(B, mb["order"], mb["meta_new"], mb["meta_mrg"], mb["Dsize"])
    Cm = bk.dot(Am, Bm, st["dot"]["meta_dot"], st["dot"]["Dsize"])
    for rec in st["dot"]["meta_dot"][:: max(1, len(st["dot"]["meta_dot"]) // 5)]:
        slc, Dc, sla, Da, slb, Db = rec
        ref = Am[sla[0]:sla[1]].view(Da) @ Bm[slb[0]:slb[1]].view(Db)
        got = Cm[slc[0]:slc[1]].view(Dc)
        assert float(torch.linalg.norm(got - ref) / torch.linalg.norm(ref)) <= 1e-12


@pytest.mark.parametrize("dtype", ["float64", "complex128"])
def test_merge_zero_fills_only_the_holes(bk, dtype):
    """Merged blocks with missing source blocks (1-, 2-, 3-d targets, ragged grids, single uncovered columns): the kernel
    writes zeros into exactly the uncovered cells (zero-fill runs) — the destination is NOT memset — and every other element
    of the output buffer, here pre-poisoned with NaN through the caching allocator, belongs to some record."""
    rng = np.random.default_rng(0)
    from yastn_b200 import plans as _plans
    for g, Dn in ((1, (11,)), (2, (7, 9)), (3, (4, 5, 6)), (2, (5, 1)), (3, (3, 1, 4)), (2, (300, 170)), (2, (2000, 3))):
        cuts = [sorted({0, d} | set(rng.integers(0, d + 1, 3).tolist())) for d in Dn]
        cells = [tuple(c) for c in np.ndindex(*[len(c) - 1 for c in cuts])]
        keep = [c for c in cells if rng.random() < 0.55] or cells[:1]
        vol = int(np.prod(Dn))
        meta_new = (((0,), Dn, (0, vol)),)
        meta_mrg, lo = [], 0
        for c in keep:
            box = tuple((cuts[d][c[d]], cuts[d][c[d] + 1]) for d in range(g))
            ext = tuple(b - a for a, b in box)
            meta_mrg.append(((0,), (lo, lo + int(np.prod(ext))), ext, box, ext))
            lo += int(np.prod(ext))
        meta_mrg = tuple(meta_mrg)
        x = rng.standard_normal(lo) + (1j * rng.standard_normal(lo) if dtype == "complex128" else 0)
        ref = orc.transpose_and_merge(x, tuple(range(g)), meta_new, meta_mrg, vol)
        recs, rank, covered = _plans.merge_records(tuple(range(g)), meta_new, meta_mrg)
        assert covered == vol
        poison = torch.full((vol,), float("nan"), dtype=torch.complex128 if dtype == "complex128" else torch.float64, device="cuda")
        del poison                                     # the next allocation of this size gets the poisoned block back
        out = bk.transpose_and_merge(_dev(x), tuple(range(g)), meta_new, meta_mrg, vol)
        assert np.array_equal(out.cpu().numpy(), ref), (g, Dn)


def test_rejects_cpu_and_unsupported_inputs(bk):
    case = bench_structs()["U1_D64_P1"]
    st = case["f2m"]["dot"]
    with pytest.raises(TypeError):
        bk.dot(torch.zeros(10, dtype=torch.float64), torch.zeros(10, dtype=torch.float64), st["meta_dot"], st["Dsize"])
    with pytest.raises(TypeError):
        bk.dot(torch.zeros(10, dtype=torch.float32, device="cuda"), torch.zeros(10, dtype=torch.float32, device="cuda"), st["meta_dot"], st["Dsize"])


def test_empty_and_degenerate(bk):
    # empty result (tests/tensor/test_tensordot.py:161-163 in the reference): no records, Dsize 0
    out = bk.dot(torch.zeros(6, dtype=torch.float64, device="cuda"), torch.zeros(20, dtype=torch.float64, device="cuda"), (), 0)
    assert out.numel() == 0
    out = bk.transpose_and_merge(torch.zeros(6, dtype=torch.float64, device="cuda"), (0, 1), (), (), 0)
    assert out.numel() == 0
    # 1x1x1 problem and K = 1 outer product
    A = _dev(np.array([3.0])); B = _dev(np.array([-2.0]))
    out = bk.dot(A, B, (((0, 1), (1, 1), (0, 1), (1, 1), (0, 1), (1, 1)),), 1)
    assert out.cpu().numpy().tolist() == [-6.0]
    a = np.arange(1, 6, dtype=np.float64); b = np.arange(1, 8, dtype=np.float64)
    out = bk.dot(_dev(a), _dev(b), (((0, 35), (5, 7), (0, 5), (5, 1), (0, 7), (1, 7)),), 35)
    assert np.array_equal(out.cpu().numpy(), np.outer(a, b).reshape(-1))


@pytest.mark.parametrize("name", ["U1_D64_P1", "U1_D1024_P1", "U1_D1024_P3", "Z2_D512_P1", "U1xU1_D4096_P1", "U1_D4096_P1"])
@pytest.mark.parametrize("dtype", ["float64", "complex128"])
def test_fused_dot_unmerge_matches_two_calls(bk, name, dtype):
    """The scatter epilogue (dot + unmerge in one launch) gives bit-identical data to dot followed by unmerge."""
    case = bench_structs()[name]
    st = case["f2m"]
    rng = np.random.default_rng(12)
    md, um = st["dot"]["meta_dot"], st["unmerge"]["meta"]
    na = max(r[2][1] for r in md); nb = max(r[4][1] for r in md)
    A = rng.uniform(-1, 1, na); B = rng.uniform(-1, 1, nb)
    if dtype == "complex128":
        A = A + 1j * rng.uniform(-1, 1, na); B = B + 1j * rng.uniform(-1, 1, nb)
    A, B = _dev(A), _dev(B)
    two = bk.unmerge(bk.dot(A, B, md, st["dot"]["Dsize"]), um)
    one = bk.dot_unmerge(A, B, md, st["dot"]["Dsize"], um)
    torch.cuda.synchronize()
    assert torch.equal(one, two)
    # backward of the fused call = backward of the two calls
    A1, B1 = A.clone().requires_grad_(True), B.clone().requires_grad_(True)
    A2, B2 = A.clone().requires_grad_(True), B.clone().requires_grad_(True)
    G = torch.randn_like(two)
    bk.dot_unmerge(A1, B1, md, st["dot"]["Dsize"], um).backward(G)
    bk.unmerge(bk.dot(A2, B2, md, st["dot"]["Dsize"]), um).backward(G)
    assert torch.equal(A1.grad, A2.grad) and torch.equal(B1.grad, B2.grad)


@pytest.mark.parametrize("name", ["U1_D8192_P1", "U1_D16384_P1", "U1_D16384_P2"])
@pytest.mark.parametrize("dtype", ["float64", "complex128"])
def test_fused_dot_unmerge_at_bench_sizes(bk, name, dtype):
    """The launch bench.py times (grouped GEMM with the fused unmerge scatter, D = 8192 / 16384): bit-identical to
    unmerge(dot(...)), and >= 8 sectors — the largest, the smallest, every kind of edge tile — equal to a float64 /
    complex128 NUMPY product of the merged operands (host BLAS, not cuBLAS) to 1e-12."""
    case = bench_structs()[name]
    st = case["f2m"]
    cplx = dtype == "complex128"
    gen = torch.Generator(device="cuda").manual_seed(11)

    def rnd(n):
        x = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
        return torch.complex(x, torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1) if cplx else x
    A, B = rnd(case["a"]["size"]), rnd(case["b"]["size"])
    ma, mb = st["merge_a"], st["merge_b"]
    Am = A if ma is None else bk.transpose_and_merge(A, ma["order"], ma["meta_new"], ma["meta_mrg"], ma["Dsize"])
    Bm = B if mb is None else bk.transpose_and_merge(B, mb["order"], mb["meta_new"], mb["meta_mrg"], mb["Dsize"])
    md, Dsize = st["dot"]["meta_dot"], st["dot"]["Dsize"]
    Cm = bk.dot(Am, Bm, md, Dsize)
    if st["unmerge"] is not None:
        two = bk.unmerge(Cm, st["unmerge"]["meta"])
        one = bk.dot_unmerge(Am, Bm, md, Dsize, st["unmerge"]["meta"])
        assert torch.equal(one, two)
        del one, two
    flops = [r[3][0] * r[3][1] * r[5][1] for r in md]
    order = sorted(range(len(md)), key=lambda i: flops[i])
    pick = {order[-1], order[-2], order[0], order[len(order) // 2]}
    bn = 64 if cplx else 128
    edge = {i for i in order if md[i][3][0] % 64 in range(1, 9) or md[i][5][1] % bn in range(1, 9)}    # a nearly empty last tile row / column
    pick |= set(sorted(edge, key=lambda i: -flops[i])[:3])
    pick |= set(order[:: max(1, len(order) // 5)])
    assert len(pick) >= 8
    for i in sorted(pick):
        slc, Dc, sla, Da, slb, Db = md[i]
        ref = Am[sla[0]:sla[1]].view(Da).cpu().numpy() @ Bm[slb[0]:slb[1]].view(Db).cpu().numpy()
        got = Cm[slc[0]:slc[1]].view(Dc).cpu().numpy()
        assert _relerr(got, ref) <= TOL, (name, dtype, i, Da, Db)


def test_p3_runs_on_the_skinny_path_and_is_deterministic(bk):
    """Huge-K / tiny-output contraction (pattern P3): the contraction index is shared out over thousands of warps of the
    skinny kernel (no DMMA tiles at all); partial sums are added in fixed run order, so repeated runs are bit-identical, and the
    result matches a per-sector torch.matmul.  (Stream-K on tiles: tests/test_gpu_skinny.py::test_stream_k_still_splits...)"""
    from yastn_b200 import plans as _plans
    case = bench_structs()["U1_D2048_P3"]
    st = case["f2m"]
    md = st["dot"]["meta_dot"]
    problems, segments = _plans.dot_tables(md)
    info = _plans.GemmPlan(problems, segments, 0, torch.cuda.current_device()).info()
    assert info["tiles"] == 0 and info["skinny_warps"] > 100 and info["skinny_runs"] >= info["skinny_warps"]   # 3 tiny blocks, whole GPU
    gen = torch.Generator(device="cuda").manual_seed(5)
    na = max(r[2][1] for r in md); nb = max(r[4][1] for r in md)
    A = torch.rand(na, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    B = torch.rand(nb, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    C1 = bk.dot(A, B, md, st["dot"]["Dsize"])
    for _ in range(3):
        assert torch.equal(bk.dot(A, B, md, st["dot"]["Dsize"]), C1)
    for slc, Dc, sla, Da, slb, Db in md:
        ref = A[sla[0]:sla[1]].view(Da) @ B[slb[0]:slb[1]].view(Db)
        got = C1[slc[0]:slc[1]].view(Dc)
        assert float(torch.linalg.norm(got - ref) / torch.linalg.norm(ref)) <= 1e-12


def test_gemm_unaligned_operands_and_ragged_shapes(bk):
    """Odd offsets / odd leading dimensions (8-byte aligned only), shapes that are not multiples of any tile, K = 1 .. 50."""
    rng = np.random.default_rng(21)
    recs, oa, ob, oc = [], 1, 3, 5
    for (M, K, N) in [(1, 1, 1), (3, 5, 7), (65, 17, 129), (64, 16, 128), (63, 15, 127), (130, 33, 67), (7, 50, 300), (257, 1, 9), (33, 47, 1)]:
        recs.append(((oc, oc + M * N), (M, N), (oa, oa + M * K), (M, K), (ob, ob + K * N), (K, N)))
        oa += M * K + 1; ob += K * N + 1; oc += M * N
    for dtype in ("float64", "complex128"):
        A = rng.standard_normal(oa); B = rng.standard_normal(ob)
        if dtype == "complex128":
            A = A + 1j * rng.standard_normal(oa); B = B + 1j * rng.standard_normal(ob)
        meta = tuple(recs)
        ref = orc.dot(A, B, meta, oc)
        out = bk.dot(_dev(A), _dev(B), meta, oc).cpu().numpy()
        for slc, *_ in recs:
            assert _relerr(out[slc[0]:slc[1]], ref[slc[0]:slc[1]]) <= TOL


def test_gemm_tall_and_skinny_problems(bk):
    """An MPO tensor applied to a large two-site tensor: 10^4..10^5 rows, K and N of a few elements (thousands of one-iteration
    tiles per problem), mixed with ordinary problems at odd offsets; forward, determinism and both adjoint GEMMs."""
    rng = np.random.default_rng(23)
    recs, oa, ob, oc = [], 1, 3, 5
    for (M, K, N) in [(20011, 7, 5), (300, 40, 200), (70001, 3, 1), (9000, 64, 4), (4099, 1, 33), (513, 6, 6)]:
        recs.append(((oc, oc + M * N), (M, N), (oa, oa + M * K), (M, K), (ob, ob + K * N), (K, N)))
        oa += M * K + 1; ob += K * N + 1; oc += M * N
    meta = tuple(recs)
    for dtype in ("float64", "complex128"):
        A = rng.standard_normal(oa); B = rng.standard_normal(ob)
        if dtype == "complex128":
            A = A + 1j * rng.standard_normal(oa); B = B + 1j * rng.standard_normal(ob)
        ref = orc.dot(A, B, meta, oc)
        dA, dB = _dev(A), _dev(B)
        out = bk.dot(dA, dB, meta, oc)
        again = bk.dot(dA, dB, meta, oc)
        assert torch.equal(out[recs[0][0][0]:], again[recs[0][0][0]:])      # elements below the first block are never written
        out = out.cpu().numpy()
        for slc, *_ in recs:
            assert _relerr(out[slc[0]:slc[1]], ref[slc[0]:slc[1]]) <= TOL
        # adjoint GEMMs of the same table (A_b = C_b B^H is again tall and skinny, B_b = A^H C_b has K = M huge: stream-K)
        gA = dA.clone().requires_grad_(True); gB = dB.clone().requires_grad_(True)
        G = rng.standard_normal(oc) + (1j * rng.standard_normal(oc) if dtype == "complex128" else 0)
        bk.dot(gA, gB, meta, oc).backward(_dev(G))
        ra, rb = orc.dot_backward(G, A, B, meta)
        assert _relerr(gA.grad.cpu().numpy(), ra) <= TOL and _relerr(gB.grad.cpu().numpy(), rb) <= TOL


@pytest.mark.parametrize("name", ["U1_D1024_P1", "U1_D2048_P3", "U1_D4096_T1"])
def test_cuda_graph_capture_and_replay(bk, name):
    """The run calls allocate nothing and never synchronise (include/yastn_b200.h): a whole tensordot — merges (copy and tiled
    kernels, the zero-fill memset), grouped GEMM with stream-K partials (P3 splits tiles over CTAs; their flags must rest at 0
    between launches) and unmerge — is captured into a CUDA graph once the plans exist and replayed on new operand values."""
    case = bench_structs()[name]
    gen = torch.Generator(device="cuda").manual_seed(3)
    A = torch.rand(case["a"]["size"], dtype=torch.float64, device="cuda", generator=gen)
    B = torch.rand(case["b"]["size"], dtype=torch.float64, device="cuda", generator=gen)
    A2, B2 = torch.rand_like(A), torch.rand_like(B)
    _run_f2m(bk, A, B, case)                       # builds the plans (plan creation uploads tables: not capturable)
    expect2 = _run_f2m(bk, A2, B2, case).clone()
    expect1 = _run_f2m(bk, A, B, case).clone()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        _run_f2m(bk, A, B, case)                   # warm the capture stream's allocator pool
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(graph):
        out = _run_f2m(bk, A, B, case)
    for _ in range(2):
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, expect1)
    A.copy_(A2); B.copy_(B2)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, expect2)
