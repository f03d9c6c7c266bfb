"""Peer-memory block exchange (yastn_b200/sharding.py PeerExchange, backend_b200.dot_unmerge(dst_shift=)): host tables on CPU.

The device side writes into other ranks' HBM through mapped peer addresses; here the ranks' buffers are laid side by side in
ONE numpy array (a "virtual address space": rank r's buffer starts at r * N, so the shift from rank r to rank p is
(p - r) * N elements) and the tables are executed by the numpy interpreters of the C ABI (tests/table_exec.py).  Checked
on a recorded chain of two contractions that group the shared tensor by different legs (tests/golden/structs_chain.json.gz):
every block reaches exactly the rank that multiplies it next, by the separate exchange launch and by the fused GEMM epilogue,
and the sharded chain equals the unsharded oracle.  The real NVLink path is tools/multigpu_chain.py (2+ GPUs).
"""
import numpy as np
import pytest

from golden_io import chain_structs
from oracle import backend_oracle as orc
from table_exec import exec_copy, exec_gemm
from yastn_b200 import plans, sharding


def _operands(case, seed=0):
    rng = np.random.default_rng(seed)
    s1, s2 = case["step1"], case["step2"]
    A = rng.standard_normal(s1["a"]["size"]); F = rng.standard_normal(s1["b"]["size"]); G = rng.standard_normal(s2["a"]["size"])
    C = orc.tensordot_f2m(A, F, s1)
    E = orc.tensordot_f2m(G, C, s2)
    return A, F, G, C, E


def _run_shard(stage, A, B, out):
    """Numpy execution of one rank's share of a fuse_to_matrix contraction; writes only the blocks it owns into ``out``."""
    ma, mb = stage["merge_a"], stage["merge_b"]
    Am = A if ma is None else orc.transpose_and_merge(A, ma["order"], ma["meta_new"], ma["meta_mrg"], ma["Dsize"])
    Bm = B if mb is None else orc.transpose_and_merge(B, mb["order"], mb["meta_new"], mb["meta_mrg"], mb["Dsize"])
    Cm = np.full(stage["dot"]["Dsize"], np.nan)
    for slc, Dc, sla, Da, slb, Db in stage["dot"]["meta_dot"]:
        Cm[slc[0]:slc[1]] = (Am[sla[0]:sla[1]].reshape(Da) @ Bm[slb[0]:slb[1]].reshape(Db)).reshape(-1)
    if stage["unmerge"] is None:
        for rec in stage["dot"]["meta_dot"]:
            out[rec[0][0]:rec[0][1]] = Cm[rec[0][0]:rec[0][1]]
    else:
        for sln, Dn, slo, Do, sub in stage["unmerge"]["meta"]:
            out[sln[0]:sln[1]] = Cm[slo[0]:slo[1]].reshape(Do)[tuple(slice(*x) for x in sub)].reshape(-1)
    return Am, Bm


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("fused", [False, True])
def test_chain_blocks_reach_their_next_owner(world, fused):
    case = chain_structs()["U1_D1024_chain"]
    A, F, G, C_full, E_full = _operands(case)
    st1, st2 = case["step1"]["f2m"], case["step2"]["f2m"]
    owner1, owner2, slices, produced_by, needed_by = sharding.chain_ownership(st1, st2, world)
    assert set(produced_by) <= set(range(world)) and any(p != q for p, q in zip(produced_by, needed_by))
    N = C_full.size
    V = np.full(world * N, np.nan)                        # the ranks' C buffers side by side
    for r in range(world):
        sh1, _ = sharding.shard_f2m(st1, r, world, panels=False)
        shifts = [(p - r) * N for p in range(world)]
        local = V[r * N:(r + 1) * N]
        if not fused:
            _run_shard(sh1, A, F, local)                   # own blocks land locally ...
            recs = sharding.push_records(slices, produced_by, needed_by, r, shifts)
            recs[:, 0] += r * N; recs[:, 1] += r * N       # ... and one copy launch stores them into the peers' buffers
            exec_copy(recs, 1, V, V)
        else:
            ma, mb = sh1["merge_a"], sh1["merge_b"]
            Am = A if ma is None else orc.transpose_and_merge(A, ma["order"], ma["meta_new"], ma["meta_mrg"], ma["Dsize"])
            Bm = F if mb is None else orc.transpose_and_merge(F, mb["order"], mb["meta_new"], mb["meta_mrg"], mb["Dsize"])
            md = sh1["dot"]["meta_dot"]
            shift = sharding.unmerge_dst_shift(sh1, slices, needed_by, r, shifts)
            problems, segments = plans.dot_tables(md)
            scatter = plans.unmerge_scatter_tables(md, sh1["unmerge"]["meta"], shift + r * N)
            exec_gemm(problems, segments, Am, Bm, V, scatter=scatter)
    # every rank now holds exactly the blocks it multiplies next (bit-exact copies of the oracle's C), the fused epilogue
    # leaves nothing behind on the producer, the separate exchange leaves the producer's copy in place
    for r in range(world):
        local = V[r * N:(r + 1) * N]
        for sl, p, q in zip(slices, produced_by, needed_by):
            have = not np.isnan(local[sl[0]:sl[1]]).any()
            want = (q == r) or (q is None and p == r) or (not fused and p == r)
            assert have == want, (r, sl, p, q)
            if have:
                assert np.array_equal(local[sl[0]:sl[1]], C_full[sl[0]:sl[1]])
    # contraction 2 on the redistributed blocks: the union over ranks is the unsharded result
    E = np.full(E_full.size, np.nan)
    for r in range(world):
        sh2, _ = sharding.shard_f2m(st2, r, world, panels=False)
        part = np.full(E_full.size, np.nan)
        _run_shard(sh2, G, V[r * N:(r + 1) * N], part)
        mine = ~np.isnan(part)
        assert np.isnan(E[mine]).all()
        E[mine] = part[mine]
    assert not np.isnan(E).any()
    assert np.linalg.norm(E - E_full) <= 1e-12 * np.linalg.norm(E_full)


def test_push_records_gather_to_all():
    """owner_new as a list of ranks: all-gather-v of whole blocks (every rank pushes what it owns to every other rank)."""
    slices = [(0, 3), (3, 10), (10, 11), (11, 11), (11, 20)]
    owner = [0, 1, 2, 0, 1]
    world, N = 3, 20
    V = np.full(world * N, np.nan)
    data = np.arange(N, dtype=np.float64)
    for r in range(world):
        for sl, o in zip(slices, owner):
            if o == r:
                V[r * N + sl[0]:r * N + sl[1]] = data[sl[0]:sl[1]]
    for r in range(world):
        recs = sharding.push_records(slices, owner, [list(range(world))] * len(slices), r, [(p - r) * N for p in range(world)])
        recs[:, 0] += r * N; recs[:, 1] += r * N
        exec_copy(recs, 1, V.copy(), V)
    assert np.array_equal(V, np.tile(data, world))
