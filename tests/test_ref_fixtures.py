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
    recs, rank, covered = plans.merge_records(fx["order"], fx["meta_new"], fx["meta_mrg"])
    assert covered <= fx["Dsize"]
    z = exec_copy(recs, rank, x, np.zeros(fx["Dsize"], dtype=x.dtype))
    assert _sha(z) == fx["sha256_" + dtype]


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
