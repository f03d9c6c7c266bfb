"""Charge-sector sharding of one block-sparse contraction over the GPUs of a box.

Under fuse_to_matrix every matched charge sector is an independent GEMM writing a disjoint C block
(yastn/tensor/_contractions.py:287-295), and merge / unmerge are per-block permutations, so a contraction
shards by sector with no data-path collective: every rank derives the same FLOP-balanced (LPT) sector ->
rank assignment from ``meta_dot`` and keeps only the merge records, GEMM problems and unmerge records of its
sectors.  Redistribution of blocks between *consecutive* contractions (different grouping leg) is the only
exchange step and is an all-to-all-v of whole blocks (``redistribute_blocks``).
"""
import heapq


def assign_sectors(meta_dot, world):
    """LPT greedy: heaviest sector first onto the least-loaded rank. Deterministic on every rank."""
    work = [(Da[0] * Da[1] * Db[1], i) for i, (slc, Dc, sla, Da, slb, Db) in enumerate(meta_dot)]
    work.sort(key=lambda x: (-x[0], x[1]))
    heap = [(0, r) for r in range(world)]
    owner = [0] * len(meta_dot)
    for w, i in work:
        load, r = heapq.heappop(heap)
        owner[i] = r
        heapq.heappush(heap, (load + w, r))
    return owner


def shard_f2m(stage, rank, world):
    """Restrict the recorded metas of one fuse_to_matrix tensordot to the sectors owned by ``rank``.

    ``stage`` = dict(merge_a, merge_b, dot, unmerge) in the reference's meta formats (None = stage skipped).
    Offsets are kept global, so shards write disjoint parts of full-size buffers and the union over ranks is
    the unsharded result.  Returns (sharded_stage, owned_flops).
    """
    meta_dot = stage["dot"]["meta_dot"]
    owner = assign_sectors(meta_dot, world)
    mine = [rec for rec, o in zip(meta_dot, owner) if o == rank]
    a_slices = {rec[2] for rec in mine}
    b_slices = {rec[4] for rec in mine}
    c_slices = {rec[0] for rec in mine}
    out = {"dot": {"meta_dot": tuple(mine), "Dsize": stage["dot"]["Dsize"]}}
    for key, keep in (("merge_a", a_slices), ("merge_b", b_slices)):
        m = stage[key]
        if m is None:
            out[key] = None
            continue
        new = tuple(x for x in m["meta_new"] if x[2] in keep)
        tns = {x[0] for x in new}
        out[key] = {"order": m["order"], "meta_new": new, "meta_mrg": tuple(x for x in m["meta_mrg"] if x[0] in tns), "Dsize": m["Dsize"]}
    u = stage["unmerge"]
    out["unmerge"] = None if u is None else {"meta": tuple(x for x in u["meta"] if x[2] in c_slices)}
    flops = sum(2 * Da[0] * Da[1] * Db[1] for (_, _, _, Da, _, Db) in mine)
    return out, flops
