"""GPU parity of the skinny-output path of the grouped GEMM (yb_skinny.cu): result blocks of at most 8 x 8 with a long
contraction index — backend.vdot (yastn/backend/backend_torch.py:537-546), SURVEY.md 8d pattern P3, and the adjoint
B_b = A^H C_b of tall-and-skinny products (yastn/backend/_backend_torch_backwards.py:136-137).  Checked through the C ABI
against the numpy table interpreter (tests/table_exec.py) and the oracle; rel. Frobenius error <= 1e-12."""
import numpy as np
import pytest
import torch

from oracle import backend_oracle as orc
from golden_io import bench_structs
from table_exec import exec_gemm

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _relerr(out, ref):
    return np.linalg.norm(out - ref) / max(np.linalg.norm(ref), 1e-300)


@pytest.fixture(scope="module")
def bk():
    from yastn_b200 import backend_b200
    return backend_b200


def _rand(rng, n, cplx):
    x = rng.uniform(-1, 1, n)
    return x + 1j * rng.uniform(-1, 1, n) if cplx else x


def _run_plan(problems, segments, A, B, csize, cplx, conj_a=False, conj_b=False, scatter=None, prefill=None):
    from yastn_b200 import plans, _lib
    plan = plans.GemmPlan(problems, segments, _lib.YB_C128 if cplx else _lib.YB_F64, torch.cuda.current_device(), scatter)
    dA, dB = _dev(A), _dev(B)
    C = torch.full((csize,), float("nan"), dtype=dA.dtype, device="cuda") if prefill is None else _dev(prefill)
    flags = (_lib.YB_GEMM_CONJ_A if conj_a else 0) | (_lib.YB_GEMM_CONJ_B if conj_b else 0)
    import ctypes
    plan.run(dA.data_ptr(), dB.data_ptr(), C.data_ptr(), flags, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return C.cpu().numpy(), plan.info()


@pytest.mark.parametrize("cplx", [False, True], ids=["f64", "c128"])
@pytest.mark.parametrize("layout", ["KC_XC", "KC_KC", "XC_XC", "XC_KC"])
def test_skinny_layouts_shapes_and_conj(layout, cplx):
    """Every (M, N) in 1..8 (complex128: 1..4) with K from 0 to 10^5 at odd offsets, all four operand layouts, two segments
    per problem, both conj flags."""
    rng = np.random.default_rng(5)
    lim = 4 if cplx else 8
    shapes = [(M, N) for M in (1, 2, 3, 4, 5, 7, 8) for N in (1, 2, 3, 4, 6, 8) if M <= lim and N <= lim]
    Ks = [0, 1, 31, 33, 1000, 100003]
    problems, segments = [], []
    oa, ob, oc = 1, 3, 5
    for i, (M, N) in enumerate(shapes):
        s0 = len(segments)
        for K in (Ks[i % len(Ks)], Ks[(i + 3) % len(Ks)]):
            al, bl = layout.split("_")
            sAm, sAk = (K + (i % 2), 1) if al == "KC" else (1, M + (i % 3))        # padded leading dimensions too
            sBk, sBn = (N + (i % 2), 1) if bl == "XC" else (1, K + (i % 3))
            segments.append((K, oa, sAm, sAk, ob, sBk, sBn))
            oa += max(1, M * max(sAm, 1) + K * max(sAk, 1)) + 1
            ob += max(1, K * max(sBk, 1) + N * max(sBn, 1)) + 1
        problems.append((M, N, oc, N, s0, len(segments)))
        oc += M * N + (i % 2)
    problems, segments = np.array(problems, dtype=np.int64), np.array(segments, dtype=np.int64)
    A, B = _rand(rng, oa + 8, cplx), _rand(rng, ob + 8, cplx)
    for conj_a, conj_b in ((False, False), (True, False), (False, True)) if cplx else ((False, False),):
        got, info = _run_plan(problems, segments, A, B, oc, cplx, conj_a, conj_b)
        assert info["tiles"] == 0 and info["skinny_warps"] > 0
        ref = exec_gemm(problems, segments, A, B, np.full(oc, np.nan, dtype=A.dtype), conj_a, conj_b)
        live = ~np.isnan(ref.real)
        assert np.array_equal(live, ~np.isnan(got.real))          # writes exactly the result blocks, nothing else
        assert _relerr(got[live], ref[live]) <= TOL
        for (M, N, offC, ldc, s0, s1) in problems:                 # and block by block (small blocks are not hidden by big ones)
            blk = slice(offC, offC + M * N)
            assert _relerr(got[blk], ref[blk]) <= TOL


@pytest.mark.parametrize("name", ["U1_D64_P3", "U1_D4096_P3", "U1_D16384_P3", "U1xU1_D4096_P3", "Z2_D512_P3"])
@pytest.mark.parametrize("dtype", ["float64", "complex128"])
def test_pattern_p3_pipeline_vs_oracle(bk, name, dtype):
    """SURVEY 8d pattern P3 (both big legs contracted: K up to 1.2e7, result blocks 1x1 .. 4x4) at the benchmark sizes,
    merge -> dot -> unmerge against the CPU oracle; repeated launches are bit-identical; fused dot_unmerge is bit-identical."""
    case = bench_structs()[name]
    rng = np.random.default_rng(2)
    cplx = dtype == "complex128"
    A, B = _rand(rng, case["a"]["size"], cplx), _rand(rng, case["b"]["size"], cplx)
    ref = orc.tensordot_f2m(A, B, case)
    st = case["f2m"]
    ma, mb = st["merge_a"], st["merge_b"]
    dA, dB = _dev(A), _dev(B)
    Am = dA if ma is None else bk.transpose_and_merge(dA, ma["order"], ma["meta_new"], ma["meta_mrg"], ma["Dsize"])
    Bm = dB if mb is None else bk.transpose_and_merge(dB, mb["order"], mb["meta_new"], mb["meta_mrg"], mb["Dsize"])
    C = bk.dot(Am, Bm, st["dot"]["meta_dot"], st["dot"]["Dsize"])
    for _ in range(3):
        assert torch.equal(bk.dot(Am, Bm, st["dot"]["meta_dot"], st["dot"]["Dsize"]), C)
    out = bk.unmerge(C, st["unmerge"]["meta"]) if st["unmerge"] is not None else C
    assert _relerr(out.cpu().numpy(), ref) <= TOL
    if st["unmerge"] is not None:
        assert torch.equal(bk.dot_unmerge(Am, Bm, st["dot"]["meta_dot"], st["dot"]["Dsize"], st["unmerge"]["meta"]), out)


@pytest.mark.parametrize("cplx", [False, True], ids=["f64", "c128"])
def test_vdot_large_and_many_blocks(bk, cplx):
    """backend.vdot: a few 10^7-element blocks, and 3000 small blocks of 1..400 elements, against numpy; deterministic."""
    rng = np.random.default_rng(9)
    for sizes in ([12_000_001, 7, 9_999_999, 1], list(rng.integers(1, 400, 3000))):
        lo, meta = 0, []
        for n in sizes:
            meta.append(((lo, lo + int(n)), (lo, lo + int(n))))
            lo += int(n)
        meta = tuple(meta)
        A, B = _rand(rng, lo, cplx), _rand(rng, lo, cplx)
        ref = sum(np.dot(A[a0:a1], B[b0:b1]) for (a0, a1), (b0, b1) in meta)
        dA, dB = _dev(A), _dev(B)
        got = bk.vdot(dA, dB, meta)
        assert torch.equal(bk.vdot(dA, dB, meta), got)
        assert abs(complex(got.item()) - ref) <= 1e-12 * np.sqrt(lo) * max(1.0, abs(ref))
        if cplx:   # <a|b>: torch's lazy conj bit on the first operand (yastn.vdot, yastn/tensor/_contractions.py:590-630)
            got = bk.vdot(dA.conj(), dB, meta)
            ref = sum(np.vdot(A[a0:a1], B[b0:b1]) for (a0, a1), (b0, b1) in meta)
            assert abs(complex(got.item()) - ref) <= 1e-12 * np.sqrt(lo) * max(1.0, abs(ref))


@pytest.mark.parametrize("cplx", [False, True], ids=["f64", "c128"])
def test_mixed_plan_skinny_and_tiles_with_scatter(cplx):
    """One plan holding tile problems and long skinny problems (two kernels in one yb_gemm_run), with the fused-unmerge
    scatter epilogue on both kinds, against the table interpreter."""
    rng = np.random.default_rng(13)
    shapes = [(4, 4, 60001), (130, 70, 33), (1, 1, 250000), (2, 3, 5000), (65, 129, 17), (3, 2, 7)]   # last: short skinny -> stays a tile
    problems, segments, oa, ob, oc = [], [], 0, 0, 0
    row_ptr, row_cuts, col_ptr, col_cuts, dst_ptr, dst = [0], [], [0], [], [0], []
    for i, (M, N, K) in enumerate(shapes):
        segments.append((K, oa, K, 1, ob, N, 1))
        problems.append((M, N, oc, N, i, i + 1))
        oa += M * K; ob += K * N
        rc, cc = sorted({0, M // 2, M}), sorted({0, N // 3, N})
        row_cuts += rc; col_cuts += cc
        row_ptr.append(len(row_cuts)); col_ptr.append(len(col_cuts))
        # sub-blocks are stored back to back in reverse order inside the problem's range of C
        blocks = [(r, c) for r in range(len(rc) - 1) for c in range(len(cc) - 1)]
        off, place = oc, {}
        for (r, c) in reversed(blocks):
            place[(r, c)] = off
            off += (rc[r + 1] - rc[r]) * (cc[c + 1] - cc[c])
        dst += [place[b] for b in blocks]
        dst_ptr.append(len(dst))
        oc += M * N
    arr = lambda x: np.array(x, dtype=np.int64)
    problems, segments = arr(problems), arr(segments)
    scatter = (arr(range(len(shapes))), arr(row_ptr), arr(row_cuts), arr(col_ptr), arr(col_cuts), arr(dst_ptr), arr(dst))
    A, B = _rand(rng, oa, cplx), _rand(rng, ob, cplx)
    got, info = _run_plan(problems, segments, A, B, oc, cplx, scatter=scatter)
    assert info["tiles"] > 0 and info["skinny_warps"] > 0
    ref = exec_gemm(problems, segments, A, B, np.zeros(oc, dtype=A.dtype), scatter=scatter)
    o = 0
    for (M, N, K) in shapes:
        assert _relerr(got[o:o + M * N], ref[o:o + M * N]) <= TOL
        o += M * N


def test_stream_k_still_splits_and_is_deterministic(bk):
    """A 16 x 16 result with K = 4e6 is too large for the skinny path: one tile whose k-range is shared by every CTA
    (stream-K, cooperative launch); repeated runs are bit-identical and match numpy."""
    from yastn_b200 import plans
    M, N, K = 16, 16, 4_000_037
    meta = (((0, M * N), (M, N), (0, M * K), (M, K), (0, K * N), (K, N)),)
    problems, segments = plans.dot_tables(meta)
    info = plans.GemmPlan(problems, segments, 0, torch.cuda.current_device()).info()
    assert info["split_ctas"] > 100 and info["grid"] > 100 and info["skinny_warps"] == 0
    rng = np.random.default_rng(3)
    A, B = rng.uniform(-1, 1, M * K), rng.uniform(-1, 1, K * N)
    dA, dB = _dev(A), _dev(B)
    C1 = bk.dot(dA, dB, meta, M * N)
    for _ in range(3):
        assert torch.equal(bk.dot(dA, dB, meta, M * N), C1)
    assert _relerr(C1.cpu().numpy(), (A.reshape(M, K) @ B.reshape(K, N)).reshape(-1)) <= TOL


def test_concurrent_streams_do_not_share_reduction_scratch(bk):
    """Two stream-K GEMMs and two skinny reductions in flight on two streams at once (the cross-CTA scratch is per
    (device, stream)), plus the same calls while an NCCL-free busy kernel occupies SMs on a third stream: every result equals
    the serial one bit for bit."""
    rng = np.random.default_rng(17)
    M, N, K = 24, 24, 1_500_003
    meta_t = (((0, M * N), (M, N), (0, M * K), (M, K), (0, K * N), (K, N)),)
    Ks = 3_000_001
    meta_s = (((0, 4), (2, 2), (0, 2 * Ks), (2, Ks), (0, Ks * 2), (Ks, 2)),)
    ops = []
    for _ in range(2):
        ops.append((meta_t, M * N, _dev(rng.uniform(-1, 1, M * K)), _dev(rng.uniform(-1, 1, K * N))))
        ops.append((meta_s, 4, _dev(rng.uniform(-1, 1, 2 * Ks)), _dev(rng.uniform(-1, 1, Ks * 2))))
    serial = [bk.dot(A, B, meta, n).clone() for meta, n, A, B in ops]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    busy = torch.cuda.Stream()
    X = torch.rand(4096, 4096, device="cuda")
    for rep in range(5):
        outs = [None] * len(ops)
        with torch.cuda.stream(busy):
            for _ in range(4):
                X = torch.sin(X) @ X.t() * 1e-4
        for i, (meta, n, A, B) in enumerate(ops):
            with torch.cuda.stream(streams[i % 2]):
                outs[i] = bk.dot(A, B, meta, n)
        torch.cuda.synchronize()
        for o, s in zip(outs, serial):
            assert torch.equal(o, s)


def test_skinny_adjoint_of_tall_products(bk):
    """B_b = A^H C_b for a tall-and-skinny product has a tiny result and contracts over the 10^5 rows: it runs on the skinny
    kernel (both operands stored as rows of a few contiguous elements) and matches the oracle adjoint."""
    rng = np.random.default_rng(23)
    recs, oa, ob, oc = [], 1, 3, 5
    for (M, K, N) in [(200011, 7, 5), (70001, 3, 1), (90000, 8, 4), (123457, 1, 8)]:
        recs.append(((oc, oc + M * N), (M, N), (oa, oa + M * K), (M, K), (ob, ob + K * N), (K, N)))
        oa += M * K + 1; ob += K * N + 1; oc += M * N
    meta = tuple(recs)
    for cplx in (False, True):
        A, B = _rand(rng, oa, cplx), _rand(rng, ob, cplx)
        gA, gB = _dev(A).requires_grad_(True), _dev(B).requires_grad_(True)
        G = _rand(rng, oc, cplx)
        bk.dot(gA, gB, meta, oc).backward(_dev(G))
        ra, rb = orc.dot_backward(G, A, B, meta)
        assert _relerr(gA.grad.cpu().numpy(), ra) <= TOL and _relerr(gB.grad.cpu().numpy(), rb) <= TOL


@pytest.mark.parametrize("cplx", [False, True], ids=["f64", "c128"])
def test_panel_products_all_layouts(cplx):
    """A tiny matrix times a very long one (K <= 8, one of M, N <= 8: an MPO block applied to an environment), both
    orientations, every operand layout (row-major and transposed views of either operand, padded leading dimensions), conj
    flags, two segments per problem, next to tile and skinny problems in the same plan; against the table interpreter."""
    rng = np.random.default_rng(31)
    shapes = [(4, 4, 300_003), (1, 3, 50_001), (300_007, 4, 2), (100_001, 1, 8), (8, 8, 70_001), (77_777, 8, 8), (2, 1, 1_000_000),
              (3, 5, 257), (200, 3, 3), (130, 40, 70), (3, 100_000, 2)]         # the last three: too short / a tile / a skinny problem
    for variant in range(3):
        problems, segments = [], []
        oa, ob, oc = 1, 3, 5
        for i, (M, K, N) in enumerate(shapes):
            s0 = len(segments)
            for Ks in ((K,) if i % 3 else (K, max(1, K // 2))):
                if variant == 0:      # row-major A (M x K), row-major B (K x N)
                    sAm, sAk, sBk, sBn = Ks, 1, N, 1
                elif variant == 1:    # A^T and B^T views with padded leading dimensions
                    sAm, sAk, sBk, sBn = 1, M + 1, 1, Ks + 2
                else:                 # padded row-major
                    sAm, sAk, sBk, sBn = Ks + 3, 1, N + 1, 1
                segments.append((Ks, oa, sAm, sAk, ob, sBk, sBn))
                oa += (M - 1) * sAm + (Ks - 1) * sAk + 2
                ob += (Ks - 1) * sBk + (N - 1) * sBn + 2
            problems.append((M, N, oc, N + (i % 2), s0, len(segments)))
            oc += M * (N + (i % 2)) + 1
        problems, segments = np.array(problems, dtype=np.int64), np.array(segments, dtype=np.int64)
        A, B = _rand(rng, oa + 4, cplx), _rand(rng, ob + 4, cplx)
        for conj_a, conj_b in (((False, False), (True, False), (False, True)) if cplx else ((False, False),)):
            got, info = _run_plan(problems, segments, A, B, oc, cplx, conj_a, conj_b)
            assert info["panel_units"] > 0 and info["tiles"] > 0 and info["skinny_warps"] > 0
            ref = exec_gemm(problems, segments, A, B, np.full(oc, np.nan, dtype=A.dtype), conj_a, conj_b)
            live = ~np.isnan(ref.real)
            assert np.array_equal(live, ~np.isnan(got.real))
            for (M, N, offC, ldc, _, _) in problems:
                blk = (offC + np.arange(M)[:, None] * ldc + np.arange(N)[None, :]).reshape(-1)
                assert _relerr(got[blk], ref[blk]) <= TOL, (variant, M, N)


def test_panel_products_through_dot_and_adjoint(bk):
    """backend.dot on MPO-application shapes (plain row-major blocks) and its two adjoint GEMMs: A_b = C_b B^H is again a
    panel product, B_b = A^H C_b a skinny reduction over 10^5..10^6 rows."""
    rng = np.random.default_rng(37)
    recs, oa, ob, oc = [], 0, 0, 0
    for (M, K, N) in [(4, 4, 1_000_003), (2, 3, 250_000), (400_001, 4, 4), (1, 1, 70_000), (90_001, 2, 1)]:
        recs.append(((oc, oc + M * N), (M, N), (oa, oa + M * K), (M, K), (ob, ob + K * N), (K, N)))
        oa += M * K; ob += K * N; oc += M * N
    meta = tuple(recs)
    for cplx in (False, True):
        A, B = _rand(rng, oa, cplx), _rand(rng, ob, cplx)
        ref = orc.dot(A, B, meta, oc)
        gA, gB = _dev(A).requires_grad_(True), _dev(B).requires_grad_(True)
        out = bk.dot(gA, gB, meta, oc)
        assert _relerr(out.detach().cpu().numpy(), ref) <= TOL
        G = _rand(rng, oc, cplx)
        out.backward(_dev(G))
        ra, rb = orc.dot_backward(G, A, B, meta)
        assert _relerr(gA.grad.cpu().numpy(), ra) <= TOL and _relerr(gB.grad.cpu().numpy(), rb) <= TOL
