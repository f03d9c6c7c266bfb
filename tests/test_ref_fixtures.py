"""The reference's own captured merge fixtures (experimental/main_d7chi98_1d.py:18-61): a D=7 chi=98 CTMRG merge
(rank 6, 281 blocks, order (0,2,4,5,1,3), zero padding, skinny targets), a small rank-5 merge and a rank-5 -> 3-D fuse.
Expected outputs are SHA-256 hashes of what the reference's own numpy implementation produced (tests/golden/
make_ref_fixtures.py): the op is a permutation, parity is bit-exact.  CPU: oracle and the C-ABI copy tables (numpy
interpreter); GPU: the CUDA kernel through the C ABI, forward and adjoint."""
import hashlib

import numpy as np
import pytest

from golden_io import ref_merge_fixtures, closed_form_input
from oracle import backend_oracle as orc
from table_exec import exec_copy
from yastn_b200 import plans


def _sha(x):
    return hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()


@pytest.mark.parametrize("dtype", ["float64", "complex128"])
@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_oracle_and_tables_reproduce_reference_hash(name, dtype):
    fx = ref_merge_fixtures()[name]
    x = closed_form_input(fx["data_size"], dtype)
    y = orc.transpose_and_merge(x, fx["order"], fx["meta_new"], fx["meta_mrg"], fx["Dsize"])
    assert _sha(y) == fx["sha256_" + dtype]
    assert int(np.count_nonzero(y)) == fx["nonzero_" + dtype]
    # without zero-fill records the destination must be cleared first ...
    recs, rank, covered = plans.merge_records(fx["order"], fx["meta_new"], fx["meta_mrg"], zero_records=False)
    assert covered <= fx["Dsize"]
    z = exec_copy(recs, rank, x, np.zeros(fx["Dsize"], dtype=x.dtype))
    assert _sha(z) == fx["sha256_" + dtype]
    # ... with them (the product path) every destination element is written exactly once: a NaN-filled buffer comes out right
    recs, rank, covered_z = plans.merge_records(fx["order"], fx["meta_new"], fx["meta_mrg"])
    assert covered_z == fx["Dsize"] and (covered < fx["Dsize"]) == bool((recs[:, 0] == plans.SRC_ZERO).any())
    z = exec_copy(recs, rank, x, np.full(fx["Dsize"], np.nan, dtype=x.dtype))
    assert _sha(z) == fx["sha256_" + dtype]
    count = np.zeros(fx["Dsize"], dtype=np.int64)
    r = rank
    for rec in recs:
        idx = np.indices(tuple(int(e) for e in rec[2:2 + r])).reshape(r, -1)
        np.add.at(count, rec[1] + (idx * rec[2 + 2 * r:, None]).sum(axis=0), 1)
    assert (count == 1).all()


def test_zero_records_fill_holes_of_nd_targets():
    """fuse_legs targets (N-d merged blocks, yastn/tensor/_merging.py:304-377) with missing source blocks: the zero-fill
    records are exactly the complement of the source boxes, for 1-, 2- and 3-d targets, including single uncovered columns."""
    rng = np.random.default_rng(0)
    for g, Dn in ((1, (11,)), (2, (7, 9)), (3, (4, 5, 6)), (2, (5, 1)), (3, (3, 1, 4))):
        cuts = [sorted({0, d} | set(rng.integers(0, d + 1, 2).tolist())) for d in Dn]
        cells = [tuple(c) for c in np.ndindex(*[len(c) - 1 for c in cuts])]
        keep = [c for c in cells if rng.random() < 0.55] or cells[:1]
        meta_new = (((0,), Dn, (3, 3 + int(np.prod(Dn)))),)
        meta_mrg, lo = [], 0
        for c in keep:
            box = tuple((cuts[d][c[d]], cuts[d][c[d] + 1]) for d in range(g))
            ext = tuple(b - a for a, b in box)
            meta_mrg.append(((0,), (lo, lo + int(np.prod(ext))), ext, box, ext))
            lo += int(np.prod(ext))
        x = rng.standard_normal(lo)
        Dsize = 3 + int(np.prod(Dn)) + 2
        ref = orc.transpose_and_merge(x, tuple(range(g)), meta_new, tuple(meta_mrg), Dsize)
        recs, rank, covered = plans.merge_records(tuple(range(g)), meta_new, tuple(meta_mrg))
        assert covered == int(np.prod(Dn))
        out = exec_copy(recs, rank, x, np.full(Dsize, np.nan))
        inside = slice(3, 3 + int(np.prod(Dn)))
        assert np.array_equal(out[inside], ref[inside]) and np.isnan(out[:3]).all() and np.isnan(out[inside.stop:]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "complex128"])
@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_cuda_merge_reproduces_reference_hash(name, dtype):
    import torch
    from yastn_b200 import backend_b200 as bk
    fx = ref_merge_fixtures()[name]
    x = closed_form_input(fx["data_size"], dtype)
    X = torch.from_numpy(x).cuda()
    Y = bk.transpose_and_merge(X, fx["order"], fx["meta_new"], fx["meta_mrg"], fx["Dsize"])
    assert _sha(Y.cpu().numpy()) == fx["sha256_" + dtype]
    # lazy-conjugated input resolves inside the kernel
    if dtype == "complex128":
        Yc = bk.transpose_and_merge(torch.from_numpy(x.conj()).cuda().conj(), fx["order"], fx["meta_new"], fx["meta_mrg"], fx["Dsize"])
        assert _sha(Yc.cpu().numpy()) == fx["sha256_" + dtype]
    # adjoint: scatter the merged data back; padding is dropped, every source element returns to its place
    Xg = X.clone().requires_grad_(True)
    out = bk.transpose_and_merge(Xg, fx["order"], fx["meta_new"], fx["meta_mrg"], fx["Dsize"])
    out.backward(out.detach().conj() if dtype == "complex128" else out.detach())
    covered = np.zeros(fx["data_size"], dtype=bool)
    for rec in fx["meta_mrg"]:
        covered[rec[1][0]:rec[1][1]] = True
    g = Xg.grad.cpu().numpy()
    expect = np.where(covered, x.conj() if dtype == "complex128" else x, 0)
    assert np.array_equal(g, expect)
