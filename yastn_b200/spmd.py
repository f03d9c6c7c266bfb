"""Sector-sharded execution of an unmodified YASTN program on several GPUs (BASELINE config 3 "sectors sharded
over 1/2/4/8 B200").

One process per GPU, every process runs the SAME YASTN program on replicated tensors (SPMD) — DMRG, CTMRG, a user's script —
and the two kinds of work a sweep is made of are split between the ranks from metadata alone:

* **contractions**: a ``fuse_to_matrix`` tensordot (yastn/tensor/_contractions.py:139-156) above ``min_flops`` is cut into
  FLOP-balanced row panels of its charge sectors (``sharding.shard_f2m`` — a Z2 tensor has two sectors and still uses 8 GPUs).
  A rank merges only the source blocks that fill its rows of A and the B blocks of its sectors and multiplies its panels into
  the merged result.  The panels of a rank are ONE contiguous range of that 1-D result, so completing it on every rank is one
  broadcast per rank over NVLink/NVSwitch (NCCL, on the caller's stream; pure copies — an all-reduce of a zero-padded result
  was measured 4x slower because NCCL's float64 sum runs at 160 GB/s) followed by the local unmerge.  Every element is
  produced by exactly one rank, so all ranks hold the same bits and take the same decisions afterwards (truncation,
  convergence tests).
* **decompositions**: the charge sectors of ``svd`` / ``eigh`` / ``qr`` are dealt to the ranks (LPT on the sector cost),
  factorised into zeroed outputs and all-reduced the same way (yastn_b200.decomp).

Small contractions and everything else (vector operations, metadata) run replicated: they are launch-bound, an exchange would
cost more than it saves.  The reference has no multi-GPU path for a single tensor network contraction; its only multi-process
code farms whole CTM environments out to workers (yastn/tn/fpeps/envs/_env_ctm_dist_mp.py:117-168).

    torchrun --nproc-per-node 4 my_dmrg.py        # my_dmrg.py: dist.init_process_group("nccl"); yastn_b200.spmd.enable()
"""
import torch

from . import backend_b200 as _bk
from . import sharding as _sh

_state = {"group": None, "rank": 0, "world": 1, "min_flops": 4.0e9, "saved_f2m": None, "profile": False,
          "stats": {"sharded": 0, "replicated": 0, "allreduce_bytes": 0, "sharded_flops": 0.0, "sharded_s": 0.0, "allreduce_s": 0.0}}
_plans = {}
_arena = {"arena": None, "slot_bytes": 0, "next": 0, "disabled": False}      # two result slots in peer-mapped memory (double buffer)


def active():
    return _state["world"] > 1 and _state["saved_f2m"] is not None


def rank_world():
    return _state["rank"], _state["world"]


def stats():
    return dict(_state["stats"])


def set_profile(on):
    """Device-synchronised timing of the sharded contractions and of the all-reduces into ``stats()`` (perturbs the totals)."""
    _state["profile"] = bool(on)


def all_reduce_(t):
    """Sum ``t`` over the ranks in place, on the caller's stream (complex tensors through their real view)."""
    import time
    import torch.distributed as dist
    buf = torch.view_as_real(t) if t.is_complex() else t
    if _state["profile"] and t.is_cuda:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dist.all_reduce(buf, group=_state["group"])
        torch.cuda.synchronize()
        _state["stats"]["allreduce_s"] += time.perf_counter() - t0
    else:
        dist.all_reduce(buf, group=_state["group"])
    _state["stats"]["allreduce_bytes"] += t.numel() * t.element_size()
    return t


def exchange_(t, runs):
    """Complete ``t`` on every rank: ``runs[r]`` are the contiguous element ranges rank ``r`` has computed; each is broadcast from
    its owner (pure copies over NVLink on the caller's stream — no reduction, so every rank ends with the same bits)."""
    import time
    import torch.distributed as dist
    prof = _state["profile"] and t.is_cuda
    if prof:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
    flat = torch.view_as_real(t).reshape(-1) if t.is_complex() else t
    k = 2 if t.is_complex() else 1
    group = _state["group"]
    for r, rr in enumerate(runs):
        src = r if group is None else dist.get_global_rank(group, r)
        for lo, hi in rr:
            dist.broadcast(flat[k * lo:k * hi], src=src, group=group)
            _state["stats"]["allreduce_bytes"] += (hi - lo) * t.element_size()
    if prof:
        torch.cuda.synchronize()
        _state["stats"]["allreduce_s"] += time.perf_counter() - t0
    return t


def _arena_slot(nbytes, device):
    """A result slot of at least ``nbytes`` in this rank's peer arena (every rank makes the same calls in the same order, so the
    slots sit at the same offsets everywhere).  Two slots alternate: a peer may already push the panels of the next contraction
    while this rank still unmerges the previous one.  Growing the arena is collective."""
    from . import peer
    a = _arena
    if a["arena"] is None or a["slot_bytes"] < nbytes:
        if a["arena"] is not None:
            a["arena"].close()
        slot = max(int(nbytes * 1.25), 1 << 30)          # growing is collective (IPC handles are re-opened on every peer): start large
        slot = (slot + 4095) // 4096 * 4096
        a["arena"] = peer.PeerArena(2 * slot, device=device, group=_state["group"])
        a["slot_bytes"], a["next"] = slot, 0
        _plans_push.clear()
    k = a["next"]
    a["next"] ^= 1
    return a["arena"], k, k * a["slot_bytes"]


_plans_push = {}


def _push_plan(arena, slot_index, slot_off, runs, itemsize, device_index):
    """Copy plan that stores this rank's contiguous panels into the same slot of every peer (one launch, SM stores over NVLink)."""
    import numpy as np
    from . import plans
    key = (id(runs), slot_index, itemsize)
    ent = _plans_push.get(key)
    if ent is None or ent[0] is not runs:
        recs = []
        for peer_rank in range(arena.world):
            if peer_rank == arena.rank:
                continue
            sh = arena.shift(peer_rank, itemsize)
            for lo, hi in runs[arena.rank]:
                pos = lo
                while pos < hi:                     # records stay below 2^31 elements
                    n = min(hi - pos, (1 << 30))
                    recs.append([pos, pos + sh, n, 1, 1])
                    pos += n
        plan = plans.CopyPlan(np.array(recs, dtype=np.int64).reshape(len(recs), 5), 1, itemsize, device_index) if recs else None
        if len(_plans_push) > 4096:
            _plans_push.clear()
        ent = (runs, plan)
        _plans_push[key] = ent
    return ent[1]


def _exchange_arena(merged, arena, slot_index, slot_off, runs):
    """``merged`` is this rank's slot: push the panels computed here into the slot of every peer, then publish."""
    import ctypes
    import time
    prof = _state["profile"]
    if prof:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
    plan = _push_plan(arena, slot_index, slot_off, runs, merged.element_size(), merged.device.index)
    if plan is not None:
        st = ctypes.c_void_p(torch.cuda.current_stream(merged.device).cuda_stream)
        plan.run(merged.data_ptr(), merged.data_ptr(), merged.numel(), 0, st)
    arena.publish()
    _state["stats"]["allreduce_bytes"] += sum(hi - lo for lo, hi in runs[arena.rank]) * merged.element_size() * (arena.world - 1)
    if prof:
        torch.cuda.synchronize()
        _state["stats"]["allreduce_s"] += time.perf_counter() - t0


def _stage_for(rank, world, order_a, mm_a, mn_a, size_a, order_b, mm_b, mn_b, size_b, meta_dot, size_m, meta_unmerge):
    stage = {"merge_a": None if mm_a is None else {"order": order_a, "meta_new": mn_a, "meta_mrg": mm_a, "Dsize": size_a},
             "merge_b": None if mm_b is None else {"order": order_b, "meta_new": mn_b, "meta_mrg": mm_b, "Dsize": size_b},
             "dot": {"meta_dot": meta_dot, "Dsize": size_m},
             "unmerge": None if meta_unmerge is None else {"meta": meta_unmerge}}
    return _sh.shard_f2m(stage, rank, world, panels=True)


def _merge_part(data, m):
    """transpose_and_merge restricted to the blocks a rank needs: the destination has full size, blocks outside the rank's share
    are neither written nor read afterwards."""
    return _bk.transpose_and_merge_partial(data, m["order"], m["meta_new"], m["meta_mrg"], m["Dsize"])


def enable(group=None, min_flops=None, peer_arena=True):
    """Shard the large contractions and the decompositions of every YASTN call made from now on over the ranks of ``group``
    (default: the world group of an initialised torch.distributed).  Every rank must run the same program."""
    import torch.distributed as dist
    import yastn.tensor._contractions as C
    import yastn.tensor._merging as M
    from . import decomp
    if not dist.is_initialized():
        raise RuntimeError("yastn_b200.spmd.enable: torch.distributed is not initialised")
    _state.update(group=group, rank=dist.get_rank(group), world=dist.get_world_size(group))
    if min_flops is not None:
        _state["min_flops"] = float(min_flops)
    _arena["disabled"] = not peer_arena          # False: NCCL broadcasts instead of peer-memory stores (also the gloo / CPU path)
    decomp.set_spmd(all_reduce_ if _state["world"] > 1 else None, _state["rank"], _state["world"])
    if _state["saved_f2m"] is not None or _state["world"] == 1:
        return
    _state["saved_f2m"] = C._tensordot_f2m
    st = _state["stats"]

    def tensordot_f2m(a, b, nout_a, nin_a, nin_b, nout_b, s_c):
        backend = a.config.backend
        da, db = a._data, b._data
        native = getattr(backend, "dot_unmerge", None) is not None and da.dtype in _bk._DTYPE_CODE and da.dtype == db.dtype \
            and _bk_usable(da) and _bk_usable(db) and not (torch.is_grad_enabled() and (da.requires_grad or db.requires_grad))
        if not native:
            return _state["saved_f2m"](a, b, nout_a, nin_a, nin_b, nout_b, s_c)
        ind_a, ind_b = C._common_inds(a.struct.t, b.struct.t, nin_a, nin_b, a.ndim_n, b.ndim_n, a.config.sym.NSYM)
        struct_a, slices_a, mm_a, ls_l, ls_ac = M._meta_merge_to_matrix(a.config, a.struct, a.slices, (nout_a, nin_a), ind_a)
        struct_b, slices_b, mm_b, ls_bc, ls_r = M._meta_merge_to_matrix(b.config, b.struct, b.slices, (nin_b, nout_b), ind_b)
        if ls_ac != ls_bc:
            raise C.YastnError('Bond dimensions do not match.')
        meta_dot, struct_m, slices_m = C._meta_tensordot_f2m(struct_a, slices_a, struct_b, slices_b)
        key = (id(mm_a), id(mm_b), id(meta_dot), nout_a, nin_a, nin_b, nout_b, s_c, _state["rank"], _state["world"])
        ent = _plans.get(key)
        if ent is None or ent[0] is not mm_a or ent[1] is not mm_b or ent[2] is not meta_dot:
            flops = sum(2.0 * Da[0] * Da[1] * Db[1] for (_, _, _, Da, _, Db) in meta_dot) * (4 if da.is_complex() else 1)
            sharded = None
            if flops >= _state["min_flops"]:
                order_a, order_b = nout_a + nin_a, nin_b + nout_b
                mn_a = tuple((x, y, z.slcs[0]) for x, y, z in zip(struct_a.t, struct_a.D, slices_a))
                mn_b = tuple((x, y, z.slcs[0]) for x, y, z in zip(struct_b.t, struct_b.D, slices_b))
                plain_a = ind_a is None and tuple(range(len(order_a))) == order_a and struct_a.size == len(da) \
                    and M._no_change_in_transpose_and_merge(mm_a, mn_a, struct_a.size)
                plain_b = ind_b is None and tuple(range(len(order_b))) == order_b and struct_b.size == len(db) \
                    and M._no_change_in_transpose_and_merge(mm_b, mn_b, struct_b.size)
                meta_unmerge, struct_c, slices_c = C._meta_unmerge_matrix(a.config, struct_m, slices_m, ls_l, ls_r, s_c)
                plain_c = M._no_change_in_unmerge(meta_unmerge)
                # the unmerge is NOT sharded: a rank's row panels are one contiguous range of the merged result (`merged`), which every
                # rank can receive with a plain broadcast; cutting the result into the output blocks is then a local copy
                stage, _ = _stage_for(_state["rank"], _state["world"], order_a, None if plain_a else mm_a, mn_a, struct_a.size,
                                      order_b, None if plain_b else mm_b, mn_b, struct_b.size, meta_dot, struct_m.size, None)
                runs = []
                for units in _sh.partition_rows(meta_dot, _state["world"]):
                    spans = sorted((meta_dot[p][0][0] + r0 * meta_dot[p][5][1], meta_dot[p][0][0] + r1 * meta_dot[p][5][1]) for p, r0, r1 in units)
                    merged = []
                    for lo, hi in spans:
                        if hi <= lo:
                            continue
                        if merged and merged[-1][1] == lo:
                            merged[-1][1] = hi
                        else:
                            merged.append([lo, hi])
                    runs.append(tuple((lo, hi) for lo, hi in merged))
                sharded = (stage, struct_c, slices_c, flops, None if plain_c else meta_unmerge, tuple(runs))
            if len(_plans) > 4096:
                _plans.clear()
            ent = (mm_a, mm_b, meta_dot, sharded)
            _plans[key] = ent
        if ent[3] is None:
            st["replicated"] += 1
            return _state["saved_f2m"](a, b, nout_a, nin_a, nin_b, nout_b, s_c)
        stage, struct_c, slices_c, flops, meta_unmerge, runs = ent[3]
        if _state["profile"] and da.is_cuda:
            import time
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        if _bk._recorder is not None:
            _bk._recorder.unsupported("collective")      # a chain must never replay this without its exchange
        data_a = da if stage["merge_a"] is None else _merge_part(da, stage["merge_a"])
        data_b = db if stage["merge_b"] is None else _merge_part(db, stage["merge_b"])
        if da.is_cuda and not _arena["disabled"]:
            # the merged result lives in a peer-mapped slot: every rank stores its panels into all peers' slots with one
            # launch of the copy kernel (350-560 GB/s per rank over NVLink, §3.4), a one-element all-reduce publishes them
            arena, k, off = _arena_slot(struct_m.size * da.element_size(), da.device)
            merged = arena.view(off, struct_m.size, da.dtype)
            _bk.dot_into(data_a, data_b, stage["dot"]["meta_dot"], merged)
            _exchange_arena(merged, arena, k, off, runs)
            out = merged.clone() if meta_unmerge is None else backend.unmerge(merged, meta_unmerge)     # leaves the slot
        else:
            merged = torch.empty(struct_m.size, dtype=da.dtype, device=da.device)
            _bk.dot_into(data_a, data_b, stage["dot"]["meta_dot"], merged)
            exchange_(merged, runs)
            out = merged if meta_unmerge is None else backend.unmerge(merged, meta_unmerge)
        if _state["profile"] and da.is_cuda:
            torch.cuda.synchronize()
            st["sharded_s"] += time.perf_counter() - t0
        st["sharded"] += 1
        st["sharded_flops"] += flops
        return out, struct_c, slices_c
    C._tensordot_f2m = tensordot_f2m


def _bk_usable(d):
    return d.is_cuda


def disable():
    from . import decomp
    decomp.set_spmd(None, 0, 1)
    if _state["saved_f2m"] is not None:
        import yastn.tensor._contractions as C
        C._tensordot_f2m = _state["saved_f2m"]
        _state["saved_f2m"] = None
    _state.update(group=None, rank=0, world=1)
    _plans.clear()
    _plans_push.clear()
    if _arena["arena"] is not None:
        _arena["arena"].close()
        _arena.update(arena=None, slot_bytes=0, next=0)
