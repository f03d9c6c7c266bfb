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


def partition_rows(meta_dot, world):
    """FLOP-balanced assignment with row panels: lay the rows of all sectors on one line weighted by their cost (K*N per
    row), cut it into ``world`` equal shares.  Every rank gets a contiguous run of whole sectors plus at most two partial
    ones, balance is exact up to one row, and the result is identical on every rank (pure function of the metadata).
    A Z2 tensor has two sectors: without row panels it cannot use more than two GPUs.
    Returns per rank a list of (sector index, r0, r1)."""
    cost = [Da[0] * Da[1] * Db[1] for (_, _, _, Da, _, Db) in meta_dot]
    total = sum(cost)
    out = [[] for _ in range(world)]
    if total == 0:
        out[0] = [(p, 0, rec[3][0]) for p, rec in enumerate(meta_dot)]
        return out
    bounds = [total * r // world for r in range(world + 1)]
    prefix, r = 0, 0
    for p, (slc, Dc, sla, Da, slb, Db) in enumerate(meta_dot):
        M, per_row = Da[0], Da[1] * Db[1]
        if cost[p] == 0:
            out[min(r, world - 1)].append((p, 0, M))
            continue
        row = 0
        while row < M:
            # rows of this sector that still fit into rank r's share
            room = bounds[r + 1] - (prefix + row * per_row)
            take = M - row if r == world - 1 else min(M - row, max(0, -(-room // per_row)))
            if take == 0:
                r += 1
                continue
            out[r].append((p, row, row + take))
            row += take
            if r < world - 1 and prefix + row * per_row >= bounds[r + 1]:
                r += 1
        prefix += cost[p]
    return out


def shard_f2m(stage, rank, world, panels=True):
    """Restrict the recorded metas of one fuse_to_matrix tensordot to the work units owned by ``rank``.

    ``stage`` = dict(merge_a, merge_b, dot, unmerge) in the reference's meta formats (None = stage skipped).
    Offsets are kept global, so shards write disjoint parts of full-size buffers and the union over ranks is
    the unsharded result.  With ``panels`` sectors may be split into row panels (see ``partition_rows``): the ranks sharing a
    sector each merge its B operand and the source blocks of A that fill their own rows, and multiply only those rows;
    without, whole sectors are dealt by LPT.
    Returns (sharded_stage, owned_flops).
    """
    meta_dot = stage["dot"]["meta_dot"]
    if panels:
        units = partition_rows(meta_dot, world)[rank]
    else:
        units = [(p, 0, meta_dot[p][3][0]) for p, o in enumerate(assign_sectors(meta_dot, world)) if o == rank]
    mine, a_slices, b_slices, panel_of = [], set(), set(), {}
    for (p, r0, r1) in units:
        slc, Dc, sla, Da, slb, Db = meta_dot[p]
        K, N = Da[1], Db[1]
        a_slices.add(sla)
        b_slices.add(slb)
        pslc = (slc[0] + r0 * N, slc[0] + r1 * N)
        mine.append((pslc, (r1 - r0, N), (sla[0] + r0 * K, sla[0] + r1 * K), (r1 - r0, K), slb, Db))
        panel_of.setdefault(slc, []).append((r0, r1, pslc))
    out = {"dot": {"meta_dot": tuple(mine), "Dsize": stage["dot"]["Dsize"]}}
    # rows of every merged A block that this rank multiplies (several panels of one sector may land on one rank)
    a_rows = {}
    for (p, r0, r1) in units:
        a_rows.setdefault(meta_dot[p][2], []).append((r0, r1))
    for key, keep in (("merge_a", a_slices), ("merge_b", b_slices)):
        m = stage[key]
        if m is None:
            out[key] = None
            continue
        new = tuple(x for x in m["meta_new"] if x[2] in keep)
        tns = {x[0] for x in new}
        recs = tuple(x for x in m["meta_mrg"] if x[0] in tns)
        if key == "merge_a" and panels:
            # a source block fills the rows Dslc[0] of its merged block: only blocks that intersect this rank's row panels are
            # merged (and, in an end-to-end run, copied to the device); the other rows of the merged block are never read
            rows_of = {x[0]: a_rows[x[2]] for x in new}
            recs = tuple(x for x in recs if any(x[3][0][0] < r1 and r0 < x[3][0][1] for (r0, r1) in rows_of[x[0]]))
        out[key] = {"order": m["order"], "meta_new": new, "meta_mrg": recs, "Dsize": m["Dsize"]}
    u = stage["unmerge"]
    if u is None:
        out["unmerge"] = None
    else:
        # an output block is a contiguous (rows x width) rectangle: a panel owns the rows of it that fall inside [r0, r1)
        recs = []
        for sln, Dn, slo, Do, sub in u["meta"]:
            (a0, a1), (c0, c1) = sub
            w = c1 - c0
            for r0, r1, pslc in panel_of.get(slo, ()):
                x0, x1 = max(a0, r0), min(a1, r1)
                if x0 >= x1:
                    continue
                part = (sln, Dn) if (x0, x1) == (a0, a1) else ((sln[0] + (x0 - a0) * w, sln[0] + (x1 - a0) * w), (x1 - x0, w))
                recs.append((*part, pslc, (r1 - r0, Do[1]), ((x0 - r0, x1 - r0), sub[1])))
        out["unmerge"] = {"meta": tuple(sorted(recs))}
    flops = sum(2 * Da[0] * Da[1] * Db[1] for (_, _, _, Da, _, Db) in mine)
    return out, flops


def block_owner_from_sectors(meta_dot, owner, which):
    """Owner rank of every operand / result block of a sharded fuse_to_matrix contraction.

    which = 'c' (result blocks, keyed by slc), 'a' (merged A blocks, sla) or 'b' (slb).  Returns {slice: rank}.
    """
    pos = {"c": 0, "a": 2, "b": 4}[which]
    return {rec[pos]: o for rec, o in zip(meta_dot, owner)}


def exchange_plan(slices, owner_old, owner_new, rank, world):
    """Who sends which block to whom when ownership changes between two contractions.

    slices: block slices (lo, hi) of a tensor in storage order; owner_old / owner_new: rank per block (None = replicated /
    not needed).  Returns (sends, recvs): per peer rank the list of slices to send / receive, in storage order on both sides,
    so every rank derives matching message layouts from metadata alone (no size exchange).
    """
    sends = [[] for _ in range(world)]
    recvs = [[] for _ in range(world)]
    for sl, o_old, o_new in zip(slices, owner_old, owner_new):
        if o_old is None or o_new is None or o_old == o_new:
            continue
        if o_old == rank:
            sends[o_new].append(sl)
        if o_new == rank:
            recvs[o_old].append(sl)
    return sends, recvs


def redistribute_blocks(data, slices, owner_old, owner_new, group=None):
    """All-to-all-v of whole blocks: after the call ``data`` (a full-size 1-D buffer on every rank) holds valid
    blocks according to ``owner_new``.  One packed message per peer through grouped send/recv (ncclSend/ncclRecv under the
    NCCL backend, i.e. NVLink/NVSwitch P2P on one box); message sizes follow from the block metadata on every rank.
    """
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sends, recvs = exchange_plan(slices, owner_old, owner_new, rank, world)
    ops, staging = [], []
    for peer in range(world):
        if sends[peer]:
            buf = torch.cat([data[lo:hi] for lo, hi in sends[peer]])
            ops.append(dist.P2POp(dist.isend, buf, peer if group is None else dist.get_global_rank(group, peer), group))
        if recvs[peer]:
            n = sum(hi - lo for lo, hi in recvs[peer])
            buf = torch.empty(n, dtype=data.dtype, device=data.device)
            staging.append((buf, recvs[peer]))
            ops.append(dist.P2POp(dist.irecv, buf, peer if group is None else dist.get_global_rank(group, peer), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for buf, sls in staging:
        pos = 0
        for lo, hi in sls:
            data[lo:hi] = buf[pos:pos + hi - lo]
            pos += hi - lo
    return data


def gather_blocks(data, slices, owner, group=None):
    """Make every block valid on every rank (all-gather-v of whole blocks): used to hand a sharded result to a consumer
    that is not sharded (e.g. the SVD of a DMRG step)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ops, staging = [], []
    mine = [sl for sl, o in zip(slices, owner) if o == rank]
    sendbuf = torch.cat([data[lo:hi] for lo, hi in mine]) if mine else None
    for peer in range(world):
        if peer == rank:
            continue
        gp = peer if group is None else dist.get_global_rank(group, peer)
        if sendbuf is not None:
            ops.append(dist.P2POp(dist.isend, sendbuf, gp, group))
        theirs = [sl for sl, o in zip(slices, owner) if o == peer]
        if theirs:
            buf = torch.empty(sum(hi - lo for lo, hi in theirs), dtype=data.dtype, device=data.device)
            staging.append((buf, theirs))
            ops.append(dist.P2POp(dist.irecv, buf, gp, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for buf, sls in staging:
        pos = 0
        for lo, hi in sls:
            data[lo:hi] = buf[pos:pos + hi - lo]
            pos += hi - lo
    return data


# -------------------------------------------------------------------------------------------------
# block ownership of operands / results of a sector-sharded contraction, and the peer-memory exchange
# -------------------------------------------------------------------------------------------------

def operand_block_owner(stage, owner, which):
    """Owner rank of every block of operand ``which`` ('a' or 'b') of a fuse_to_matrix contraction whose sectors are owned
    according to ``owner`` (one rank per record of meta_dot).  Keys are the block slices (lo, hi) in the operand's own
    storage: through ``meta_mrg`` every source block feeds exactly one merged block, and a merged block belongs to the
    sector that multiplies it.  Blocks that take part in no product are absent."""
    pos = 2 if which == "a" else 4
    merged = {rec[pos]: o for rec, o in zip(stage["dot"]["meta_dot"], owner)}
    m = stage["merge_" + which]
    if m is None:
        return merged
    of_charge = {tn: merged[sln] for (tn, Dn, sln) in m["meta_new"] if sln in merged}
    return {rec[1]: of_charge[rec[0]] for rec in m["meta_mrg"] if rec[0] in of_charge}


def result_block_owner(stage, owner):
    """Owner rank of every block of the result ({slice in the result's storage: rank})."""
    merged = {rec[0]: o for rec, o in zip(stage["dot"]["meta_dot"], owner)}
    if stage["unmerge"] is None:
        return merged
    return {rec[0]: merged[rec[2]] for rec in stage["unmerge"]["meta"]}


def push_records(slices, owner_old, owner_new, rank, shifts):
    """Copy records ([src_base, dst_base, extent, 1, 1], yb_copy_plan_create with rank 1) that move every block this rank
    owns to its new owner: ``shifts[r]`` is the element offset turning a local address into rank r's (PeerArena.shift; 0 for
    the rank itself).  ``owner_new`` entries may be a rank, a list of ranks (block needed by several) or None."""
    import numpy as np
    rows = []
    for sl, o_old, o_new in zip(slices, owner_old, owner_new):
        if o_old != rank or o_new is None:
            continue
        for dest in (o_new if isinstance(o_new, (list, tuple)) else (o_new,)):
            if dest != rank and sl[1] > sl[0]:
                rows.append((sl[0], sl[0] + shifts[dest], sl[1] - sl[0], 1, 1))
    return np.array(rows, dtype=np.int64).reshape(len(rows), 5)


class PeerExchange:
    """All-to-all-v of whole blocks as ONE launch of the block-copy kernel over NVLink peer memory.

    ``data`` must be carved out of a ``peer.PeerArena`` (same offset on every rank).  Every rank stores the blocks it owns
    and another rank needs straight into that rank's copy of the buffer; ``arena.publish()`` afterwards makes them visible.
    Replaces the per-block ``torch.cat`` / send / recv / copy-back of :func:`redistribute_blocks` (which remains the
    NCCL / gloo path for buffers outside an arena)."""

    def __init__(self, arena, data, slices, owner_old, owner_new):
        from . import plans
        if not arena.contains(data):
            raise ValueError("PeerExchange: data does not live in the peer arena")
        isz = data.element_size()
        shifts = [0 if r == arena.rank else arena.shift(r, isz) for r in range(arena.world)]
        self.recs = push_records(slices, owner_old, owner_new, arena.rank, shifts)
        self.elements = int(self.recs[:, 2].sum()) if len(self.recs) else 0
        self.arena = arena
        self.plan = plans.CopyPlan(self.recs, 1, isz, data.device.index) if len(self.recs) else None

    def run(self, data, publish=True):
        import ctypes
        import torch
        if self.plan is not None:
            st = ctypes.c_void_p(torch.cuda.current_stream(data.device).cuda_stream)
            self.plan.run(data.data_ptr(), data.data_ptr(), data.numel(), 0, st)
        if publish:
            self.arena.publish()
        return data


def chain_ownership(stage1, stage2, world, operand="b"):
    """Ownership of the blocks of the tensor that contraction 1 produces and contraction 2 consumes as ``operand``.

    Both contractions are sharded by whole charge sectors (LPT).  Returns (owner1, owner2, slices, produced_by, needed_by):
    the sector owners of both contractions, the block slices of the shared tensor in storage order, the rank that produces
    each block and the rank that multiplies it next (None: the block takes part in no product of contraction 2)."""
    owner1 = assign_sectors(stage1["dot"]["meta_dot"], world)
    owner2 = assign_sectors(stage2["dot"]["meta_dot"], world)
    produced = result_block_owner(stage1, owner1)
    needed = operand_block_owner(stage2, owner2, operand)
    slices = sorted(produced)
    return owner1, owner2, slices, [produced[s] for s in slices], [needed.get(s) for s in slices]


def unmerge_dst_shift(shard_stage, slices, needed_by, rank, shifts):
    """``dst_shift`` of backend_b200.dot_unmerge for a sharded contraction: one element offset per record of the shard's
    unmerge meta, sending every produced block to the rank that needs it next (0 = stays here)."""
    import numpy as np
    nxt = dict(zip(slices, needed_by))
    out = np.zeros(len(shard_stage["unmerge"]["meta"]), dtype=np.int64)
    for k, rec in enumerate(shard_stage["unmerge"]["meta"]):
        dest = nxt.get(rec[0])
        if dest is not None and dest != rank:
            out[k] = shifts[dest]
    return out
