"""Multi-rank path on CPU (gloo, world_size 2): FLOP-balanced sector sharding of a fuse_to_matrix contraction, block
redistribution between contractions and result gathering.  The per-rank numerics run through the numpy oracle on the sharded
metas (the kernels themselves are covered by the gpu tests); what is under test is that shards are disjoint, complete, derive
identically on every rank, and that the exchange plumbing moves exactly the blocks that change owner."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import backend_oracle as orc
from golden_io import bench_structs
from yastn_b200 import sharding

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, port, name, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        case = bench_structs()[name]
        st = case["f2m"]
        rng = np.random.default_rng(3)                      # same operands on every rank
        A = rng.standard_normal(case["a"]["size"]); B = rng.standard_normal(case["b"]["size"])
        full = orc.tensordot_f2m(A, B, case)
        shard, flops = sharding.shard_f2m(st, rank, WORLD)
        # 1. run only this rank's sectors; everything else stays NaN
        Am = A if shard["merge_a"] is None else orc.transpose_and_merge(A, *[shard["merge_a"][k] for k in ("order", "meta_new", "meta_mrg", "Dsize")])
        Bm = B if shard["merge_b"] is None else orc.transpose_and_merge(B, *[shard["merge_b"][k] for k in ("order", "meta_new", "meta_mrg", "Dsize")])
        Cm = np.full(st["dot"]["Dsize"], np.nan)
        for slc, Dc, sla, Da, slb, Db in shard["dot"]["meta_dot"]:
            Cm[slc[0]:slc[1]] = (Am[sla[0]:sla[1]].reshape(Da) @ Bm[slb[0]:slb[1]].reshape(Db)).reshape(-1)
        C = np.full(st["dot"]["Dsize"], np.nan)
        if shard["unmerge"] is not None:
            for sln, Dn, slo, Do, sub in shard["unmerge"]["meta"]:
                C[sln[0]:sln[1]] = Cm[slo[0]:slo[1]].reshape(Do)[tuple(slice(*x) for x in sub)].reshape(-1)
        else:
            C = Cm
        mine = ~np.isnan(C)
        assert np.allclose(C[mine], full[mine], rtol=1e-12, atol=1e-12)
        # 2. shards are disjoint and complete: element-wise ownership sums to one over ranks
        cover = torch.from_numpy(mine.astype(np.int64))
        dist.all_reduce(cover)
        assert bool((cover == 1).all())
        fl = torch.tensor([float(flops)], dtype=torch.float64)
        gathered = [torch.zeros(1, dtype=torch.float64) for _ in range(WORLD)]
        dist.all_gather(gathered, fl)
        total = sum(float(g) for g in gathered)
        assert abs(total - sum(2 * Da[0] * Da[1] * Db[1] for (_, _, _, Da, _, Db) in st["dot"]["meta_dot"])) < 1
        imbalance = max(float(g) for g in gathered) / (total / WORLD)
        # 3. gather the result blocks on every rank (ownership of every output block derived from the shards of all ranks)
        slices, blk_owner = [], []
        for r in range(WORLD):
            sh_r, _ = sharding.shard_f2m(st, r, WORLD)
            recs = sh_r["unmerge"]["meta"] if sh_r["unmerge"] is not None else sh_r["dot"]["meta_dot"]
            slices += [m[0] for m in recs]
            blk_owner += [r] * len(recs)
        Ct = torch.from_numpy(C.copy())
        sharding.gather_blocks(Ct, slices, blk_owner)
        assert np.allclose(Ct.numpy(), full, rtol=1e-12, atol=1e-12)
        # 4. redistribute: swap the owner of every block, every rank must end with exactly its new blocks valid
        Cr = torch.from_numpy(C.copy())
        new_owner = [(o + 1) % WORLD for o in blk_owner]
        sharding.redistribute_blocks(Cr, slices, blk_owner, new_owner)
        for sl, o in zip(slices, new_owner):
            if o == rank:
                assert np.allclose(Cr.numpy()[sl[0]:sl[1]], full[sl[0]:sl[1]], rtol=1e-12, atol=1e-12)
        ret[rank] = imbalance
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["U1_D64_P1", "U1_D64_P2", "U1_D1024_P1"])
def test_sector_sharding_world2(name):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, name, ret)) for r in range(WORLD)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert len(ret) == WORLD
    assert max(ret.values()) < 1.05        # row panels keep even a 5-sector D=64 contraction balanced over 2 ranks


@pytest.mark.parametrize("name", ["U1xU1_D4096_P1", "U1_D16384_P1", "U1_D16384_P2", "Z2_D512_P1", "Z2_D512_P2", "U1_D1024_P1"])
def test_assignment_is_deterministic_balanced_and_complete(name):
    """Host-only: over 2/4/8 ranks the shards are identical when recomputed, cover every FLOP and every output element exactly
    once, and are FLOP-balanced within 10 % (Z2 has two sectors: only row panels make 8 ranks usable)."""
    st = bench_structs()[name]["f2m"]
    md = st["dot"]["meta_dot"]
    total = sum(2 * Da[0] * Da[1] * Db[1] for (_, _, _, Da, _, Db) in md)
    for world in (2, 4, 8):
        shards = [sharding.shard_f2m(st, r, world) for r in range(world)]
        assert [s[1] for s in shards] == [sharding.shard_f2m(st, r, world)[1] for r in range(world)]
        assert sum(s[1] for s in shards) == total
        assert max(s[1] for s in shards) / (total / world) < 1.02, (name, world, [s[1] for s in shards])
        cover = np.zeros(st["dot"]["Dsize"], dtype=np.int8)
        for sh, _ in shards:
            for rec in (sh["unmerge"]["meta"] if sh["unmerge"] is not None else sh["dot"]["meta_dot"]):
                cover[rec[0][0]:rec[0][1]] += 1
        assert (cover == 1).all()


@pytest.mark.parametrize("name", ["U1_D1024_P1", "U1_D16384_P2", "U1xU1_D4096_P1"])
def test_bench_shard_ranges_cover_exactly_what_a_rank_touches(name):
    """bench.py e2e leg at N > 1: the coalesced storage ranges a rank copies in / out are exactly the operand blocks its
    sharded metas read and the result blocks they write, and the result ranges of all ranks tile the result."""
    import bench
    st = bench_structs()[name]["f2m"]
    for world in (2, 8):
        cover = np.zeros(st["dot"]["Dsize"], dtype=np.int8)
        for r in range(world):
            sh, _ = sharding.shard_f2m(st, r, world)
            a, b, c = bench.shard_ranges(sh)
            for rng, key in ((a, "merge_a"), (b, "merge_b")):
                assert all(x[1] <= y[0] for x, y in zip(rng, rng[1:]))          # sorted, disjoint
                need = np.zeros(max(hi for _, hi in rng), dtype=bool)
                got = np.zeros_like(need)
                for rec in sh[key]["meta_mrg"]:
                    need[rec[1][0]:rec[1][1]] = True
                for lo, hi in rng:
                    got[lo:hi] = True
                assert np.array_equal(need, got)
            for lo, hi in c:
                cover[lo:hi] += 1
        assert (cover == 1).all()
