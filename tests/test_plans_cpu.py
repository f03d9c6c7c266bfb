"""Host logic (no GPU): the int64 tables built by yastn_b200.plans reproduce the oracle on every recorded call."""
import os
import re

import numpy as np
import pytest

from oracle import backend_oracle as orc
from golden_io import small_calls, bench_structs
from table_exec import exec_copy, exec_gemm
from yastn_b200 import plans, _lib

CALLS = small_calls()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sel(fn):
    return [c for c in CALLS if c["fn"] == fn]


@pytest.mark.parametrize("call", _sel("transpose_and_merge"), ids=lambda c: f"{c['case']}-{c['policy'][:6]}-{c['dtype']}")
def test_merge_records(call):
    a = call["args"]
    recs, rank, covered = plans.merge_records(a["order"], a["meta_new"], a["meta_mrg"])
    out = exec_copy(recs, rank, a["data"], np.zeros(a["Dsize"], dtype=a["data"].dtype))
    assert np.array_equal(out, call["out"])
    assert covered == sum(int(np.prod(m[2])) for m in a["meta_mrg"])
    # adjoint records gather the same elements back
    back = exec_copy(plans.reverse_records(recs, rank), rank, out, np.zeros_like(a["data"]))
    assert np.array_equal(back, orc.transpose_and_merge_backward(out, a["order"], a["meta_new"], a["meta_mrg"], a["data"].size))


@pytest.mark.parametrize("call", _sel("unmerge"), ids=lambda c: f"{c['case']}-{c['dtype']}")
def test_unmerge_records(call):
    a = call["args"]
    recs, rank = plans.unmerge_records(a["meta"])
    out = exec_copy(recs, rank, a["data"], np.full_like(a["data"], np.nan))
    assert np.array_equal(out, call["out"])


@pytest.mark.parametrize("call", _sel("transpose"), ids=lambda c: f"{c['case']}-{c['dtype']}")
def test_transpose_records(call):
    a = call["args"]
    recs, rank = plans.transpose_records(a["axes"], a["meta_transpose"])
    out = exec_copy(recs, rank, a["data"], np.full_like(a["data"], np.nan))
    assert np.array_equal(out, call["out"])


@pytest.mark.parametrize("call", _sel("dot"), ids=lambda c: f"{c['case']}-{c['policy'][:6]}-{c['dtype']}")
def test_dot_tables(call):
    a = call["args"]
    problems, segments = plans.dot_tables(a["meta_dot"])
    A, B = a["Adata"], a["Bdata"]
    dt = np.promote_types(A.dtype, B.dtype)
    out = exec_gemm(problems, segments, A.astype(dt), B.astype(dt), np.zeros(a["Dsize"], dtype=dt))
    assert np.linalg.norm(out - call["out"]) <= 1e-13 * max(1.0, np.linalg.norm(call["out"]))
    # backward tables against the oracle adjoint
    rng = np.random.default_rng(3)
    G = rng.standard_normal(a["Dsize"]).astype(dt)
    if dt.kind == "c":
        G = G + 1j * rng.standard_normal(a["Dsize"])
    gA_ref, gB_ref = orc.dot_backward(G, A.astype(dt), B.astype(dt), a["meta_dot"])
    pa, sa, pb, sb = plans.dot_backward_tables(a["meta_dot"])
    gA = exec_gemm(pa, sa, G, B.astype(dt), np.zeros(A.size, dtype=dt), conj_b=True)
    gB = exec_gemm(pb, sb, A.astype(dt), G, np.zeros(B.size, dtype=dt), conj_a=True)
    assert np.linalg.norm(gA - gA_ref) <= 1e-12 * max(1.0, np.linalg.norm(gA_ref))
    assert np.linalg.norm(gB - gB_ref) <= 1e-12 * max(1.0, np.linalg.norm(gB_ref))


@pytest.mark.parametrize("call", _sel("transpose_dot_sum"), ids=lambda c: f"{c['case']}-{c['dtype']}")
def test_tds_tables(call):
    a = call["args"]
    problems, segments, pack_a, pack_b = plans.tds_tables(a["meta_dot"], a["Areshape"], a["Breshape"], a["Aorder"], a["Border"])
    A, B = a["Adata"], a["Bdata"]
    dt = np.promote_types(A.dtype, B.dtype)
    A, B = A.astype(dt), B.astype(dt)
    if pack_a:
        recs, rank = plans.pack_records(a["Areshape"], a["Aorder"])
        A = exec_copy(recs, rank, A, np.zeros_like(A))
    if pack_b:
        recs, rank = plans.pack_records(a["Breshape"], a["Border"])
        B = exec_copy(recs, rank, B, np.zeros_like(B))
    out = exec_gemm(problems, segments, A, B, np.zeros(a["Dsize"], dtype=dt))
    assert np.linalg.norm(out - call["out"]) <= 1e-13 * max(1.0, np.linalg.norm(call["out"]))


def test_bench_struct_tables_small():
    """Structure fixtures at benchmark shapes: tables execute to the oracle result (small D only on CPU)."""
    for name in ("U1_D64_P1", "U1_D64_P2", "U1_D64_P3"):
        case = bench_structs()[name]
        rng = np.random.default_rng(5)
        A = rng.standard_normal(case["a"]["size"]); B = rng.standard_normal(case["b"]["size"])
        ref = orc.tensordot_f2m(A, B, case)
        st = case["f2m"]
        Am, Bm = A, B
        for side, key in (("a", "merge_a"), ("b", "merge_b")):
            m = st[key]
            if m is not None:
                recs, rank, _ = plans.merge_records(m["order"], m["meta_new"], m["meta_mrg"])
                res = exec_copy(recs, rank, A if side == "a" else B, np.zeros(m["Dsize"]))
                if side == "a":
                    Am = res
                else:
                    Bm = res
        problems, segments = plans.dot_tables(st["dot"]["meta_dot"])
        C = exec_gemm(problems, segments, Am, Bm, np.zeros(st["dot"]["Dsize"]))
        if st["unmerge"] is not None:
            recs, rank = plans.unmerge_records(st["unmerge"]["meta"])
            C = exec_copy(recs, rank, C, np.zeros_like(C))
        assert np.linalg.norm(C - ref) <= 1e-12 * np.linalg.norm(ref)


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads on a CPU-only box and exports every function include/yastn_b200.h declares."""
    header = open(os.path.join(ROOT, "include", "yastn_b200.h")).read()
    declared = set(re.findall(r"\b(yb_[a-z0-9_]+)\s*\(", header))
    lib = _lib.load()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"libyastn_b200.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert lib.yb_abi_version() == _lib.ABI_VERSION


def test_unmerge_scatter_tables():
    """Fused unmerge epilogue tables: dot with scatter == unmerge(dot) on the recorded fuse_to_matrix pipelines."""
    for name in ("U1_D64_P1", "U1_D64_P3", "Z2_D512_P1", "U1xU1_D4096_P1"):
        case = bench_structs()[name]
        st = case["f2m"]
        if st["unmerge"] is None:
            continue
        md, um = st["dot"]["meta_dot"], st["unmerge"]["meta"]
        scatter = plans.unmerge_scatter_tables(md, um)
        problems, segments = plans.dot_tables(md)
        if name == "U1xU1_D4096_P1":      # structure-only check at full size (the interpreter is too slow for the data)
            assert (scatter[0] >= 0).sum() == len({m[2] for m in um})
            assert scatter[6].size == len(um)
            continue
        rng = np.random.default_rng(8)
        na = max(r[2][1] for r in md); nb = max(r[4][1] for r in md)
        A = rng.standard_normal(na); B = rng.standard_normal(nb)
        ref = orc.unmerge(orc.dot(A, B, md, st["dot"]["Dsize"]), um)
        out = exec_gemm(problems, segments, A, B, np.full(st["dot"]["Dsize"], np.nan), scatter=scatter)
        assert np.linalg.norm(out - ref) <= 1e-13 * np.linalg.norm(ref)


def test_fused_dot_unmerge_autograd_matches_two_calls_shim():
    """Host wiring of the fused call under autograd (CPU table interpreter): forward and both gradients equal dot + unmerge."""
    import torch
    import cpu_shim
    from yastn_b200 import backend_b200 as bk
    cpu_shim.install()
    try:
        st = bench_structs()["U1_D64_P1"]["f2m"]
        md, um = st["dot"]["meta_dot"], st["unmerge"]["meta"]
        na = max(r[2][1] for r in md); nb = max(r[4][1] for r in md)
        g = torch.Generator().manual_seed(5)
        A = torch.rand(na, dtype=torch.float64, generator=g); B = torch.rand(nb, dtype=torch.float64, generator=g)
        A1, B1 = A.clone().requires_grad_(True), B.clone().requires_grad_(True)
        A2, B2 = A.clone().requires_grad_(True), B.clone().requires_grad_(True)
        one = bk.dot_unmerge(A1, B1, md, st["dot"]["Dsize"], um)
        two = bk.unmerge(bk.dot(A2, B2, md, st["dot"]["Dsize"]), um)
        assert torch.equal(one, two)
        G = torch.randn(two.shape, dtype=torch.float64, generator=g)
        one.backward(G); two.backward(G)
        assert torch.equal(A1.grad, A2.grad) and torch.equal(B1.grad, B2.grad)
        out = torch.empty(st["dot"]["Dsize"], dtype=torch.float64)
        assert bk.dot_unmerge(A, B, md, st["dot"]["Dsize"], um, out=out) is out and torch.equal(out, two.detach())
    finally:
        cpu_shim.uninstall()


def test_fused_dot_unmerge_falls_back_when_lookup_tables_would_be_huge(monkeypatch):
    """Tall-and-skinny products (10^6+ rows per merged block) would need row/column lookup tables larger than the operands:
    above the limit dot_unmerge runs as dot + unmerge, same result, also into a caller-provided ``out``."""
    import torch
    import cpu_shim
    from yastn_b200 import backend_b200 as bk
    cpu_shim.install()
    try:
        st = bench_structs()["U1_D64_P1"]["f2m"]
        md, um = st["dot"]["meta_dot"], st["unmerge"]["meta"]
        na = max(r[2][1] for r in md); nb = max(r[4][1] for r in md)
        g = torch.Generator().manual_seed(6)
        A = torch.rand(na, dtype=torch.float64, generator=g); B = torch.rand(nb, dtype=torch.float64, generator=g)
        fused = bk.dot_unmerge(A, B, md, st["dot"]["Dsize"], um)
        monkeypatch.setattr(bk, "_SCATTER_TABLE_LIMIT", 1)
        bk.clear_plan_cache()
        two = bk.dot_unmerge(A, B, md, st["dot"]["Dsize"], um)
        out = torch.empty_like(two)
        assert bk.dot_unmerge(A, B, md, st["dot"]["Dsize"], um, out=out) is out
        assert torch.equal(fused, two) and torch.equal(fused, out)
    finally:
        cpu_shim.uninstall()


def test_vdot_tables_pass_plan_validation_for_huge_blocks():
    """A joined vdot slice can hold 10^8 elements: the C ABI must accept the 1 x 1 problems (no stride limit hit).  Without a
    GPU plan creation stops at cudaSetDevice, i.e. AFTER the table validation — which is what is checked here."""
    import ctypes
    meta = (((0, 150_000_000), (7, 150_000_007)), ((150_000_000, 150_000_003), (150_000_007, 150_000_010)))
    problems, segments = plans.vdot_tables(meta)
    A = np.arange(12, dtype=np.float64); B = np.arange(20, dtype=np.float64)
    small = (((0, 5), (3, 8)), ((5, 12), (10, 17)))
    p2, s2 = plans.vdot_tables(small)
    out = exec_gemm(p2, s2, A, B, np.zeros(2))
    assert np.allclose(out, [A[0:5] @ B[3:8], A[5:12] @ B[10:17]])
    lib = _lib.load()
    for dtype_code in (_lib.YB_F64, _lib.YB_C128):
        h = ctypes.c_void_p()
        rc = lib.yb_gemm_plan_create(problems.ctypes.data_as(ctypes.c_void_p), 2, segments.ctypes.data_as(ctypes.c_void_p), 2,
                                     dtype_code, 0, ctypes.byref(h))
        msg = lib.yb_last_error().decode()
        if rc == 0:
            lib.yb_gemm_plan_destroy(h)
        else:
            assert "cudaSetDevice" in msg or "CUDA" in msg or "cuda" in msg, msg
