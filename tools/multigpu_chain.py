#!/usr/bin/env python
"""Two sharded contractions with a block redistribution in between, on N GPUs of one box (measurement + check).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multigpu_chain.py [--case U1_D4096_chain] [--dtype f64|c128] [--iters 20]

Chain (recorded from the reference, tests/golden/structs_chain.json.gz):  C = tensordot(A, F, (3, 0)) is sharded by the charge
of A's right leg, E = tensordot(G, C, (2, 0)) by the charge of C's left leg, so the blocks of C change owner in between.
Three ways of moving them, all checked against the unsharded single-GPU chain (rel. Frobenius error <= 1e-12):

  nccl      grouped ncclSend/ncclRecv of packed blocks (sharding.redistribute_blocks)
  exchange  ONE launch of the block-copy kernel storing the blocks into the peers' HBM over NVLink (sharding.PeerExchange)
  fused     no exchange launch: the grouped GEMM of contraction 1 scatters every result block straight to its next owner
            (backend_b200.dot_unmerge(out=, dst_shift=))

Prints one JSON line per variant (rank 0): ms per chain (CUDA events, max over ranks), ms and GB/s of the exchange alone.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_io import chain_structs  # noqa: E402
from yastn_b200 import backend_b200 as bk  # noqa: E402
from yastn_b200 import peer, sharding  # noqa: E402


def merged(stage, A, B):
    ma, mb = stage["merge_a"], stage["merge_b"]
    Am = A if ma is None else bk.transpose_and_merge(A, ma["order"], ma["meta_new"], ma["meta_mrg"], ma["Dsize"])
    Bm = B if mb is None else bk.transpose_and_merge(B, mb["order"], mb["meta_new"], mb["meta_mrg"], mb["Dsize"])
    return Am, Bm


def contract(stage, A, B, out=None, dst_shift=None):
    Am, Bm = merged(stage, A, B)
    return bk.dot_unmerge(Am, Bm, stage["dot"]["meta_dot"], stage["dot"]["Dsize"], stage["unmerge"]["meta"], out=out, dst_shift=dst_shift)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="U1_D4096_chain")
    ap.add_argument("--dtype", default="f64", choices=["f64", "c128"])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cplx = args.dtype == "c128"
    tdt = torch.complex128 if cplx else torch.float64
    case = chain_structs()[args.case]
    st1, st2 = case["step1"]["f2m"], case["step2"]["f2m"]
    rng = np.random.default_rng(7)                 # identical operands on every rank (A, F, G are replicated inputs)

    def rnd(n):
        x = rng.uniform(-1, 1, n)
        if cplx:
            x = x + 1j * rng.uniform(-1, 1, n)
        return torch.from_numpy(x).to(dev)
    A, F, G = rnd(case["step1"]["a"]["size"]), rnd(case["step1"]["b"]["size"]), rnd(case["step2"]["a"]["size"])
    nC, nE = st1["dot"]["Dsize"], st2["dot"]["Dsize"]
    isz = 16 if cplx else 8

    # unsharded chain on this GPU: the check
    C_full = contract(st1, A, F)
    E_full = contract(st2, G, C_full)
    torch.cuda.synchronize()

    owner1, owner2, slices, produced_by, needed_by = sharding.chain_ownership(st1, st2, world)
    sh1, fl1 = sharding.shard_f2m(st1, rank, world, panels=False)
    sh2, fl2 = sharding.shard_f2m(st2, rank, world, panels=False)
    e_owner = sharding.result_block_owner(st2, owner2)
    e_slices = sorted(e_owner)
    e_owned = [e_owner[s] for s in e_slices]

    arena = peer.PeerArena((nC + nE) * isz + 4096, dev)
    C = arena.empty(nC, tdt)
    E = arena.empty(nE, tdt)
    shifts = [0 if r == rank else arena.shift(r, isz) for r in range(world)]
    xchg = sharding.PeerExchange(arena, C, slices, produced_by, needed_by)
    gather = sharding.PeerExchange(arena, E, e_slices, e_owned, [list(range(world))] * len(e_slices))
    dst_shift = sharding.unmerge_dst_shift(sh1, slices, needed_by, rank, shifts)
    moved_bytes = xchg.elements * isz

    def chain(variant):
        if variant == "fused":
            contract(sh1, A, F, out=C, dst_shift=dst_shift)
            arena.publish()
        else:
            contract(sh1, A, F, out=C)
            if variant == "exchange":
                xchg.run(C)
            else:
                sharding.redistribute_blocks(C, slices, produced_by, needed_by)
        contract(sh2, G, C, out=E)

    def check(variant):
        C.fill_(float("nan")); E.fill_(float("nan"))
        torch.cuda.synchronize(); dist.barrier()
        chain(variant)
        gather.run(E)
        torch.cuda.synchronize(); dist.barrier()
        err = float(torch.linalg.vector_norm(E - E_full) / torch.linalg.vector_norm(E_full))
        # every block this rank multiplies in contraction 2 arrived (NaN-filled before) and matches the unsharded C
        cerr = max([float(torch.linalg.vector_norm(C[lo:hi] - C_full[lo:hi]) / max(float(torch.linalg.vector_norm(C_full[lo:hi])), 1e-300))
                    for (lo, hi), q in zip(slices, needed_by) if q == rank and hi > lo] or [0.0])
        return err, cerr

    def timed(fn, iters):
        for _ in range(3):
            fn(); arena.publish()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
            arena.publish()            # next iteration's pushes must not overtake a peer still reading (also paid by every variant)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    lines = []
    flops = sum((8 if cplx else 2) * Da[0] * Da[1] * Db[1] for st in (st1, st2) for (_, _, _, Da, _, Db) in st["dot"]["meta_dot"])
    # the same chain unsharded on one GPU (every rank runs it; reported from rank 0)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        contract(st2, G, contract(st1, A, F))
    e0.record()
    for _ in range(args.iters):
        contract(st2, G, contract(st1, A, F))
    e1.record(); torch.cuda.synchronize()
    t_single = e0.elapsed_time(e1) / args.iters
    for variant in ("nccl", "exchange", "fused"):
        err, cerr = check(variant)
        worst = torch.tensor([err, cerr], dtype=torch.float64, device=dev)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        ms = timed(lambda: chain(variant), args.iters)
        line = {"tool": "multigpu_chain", "case": args.case, "dtype": args.dtype, "n_gpus": world, "variant": variant,
                "rel_err_E": float(worst[0]), "rel_err_C_blocks": float(worst[1]), "ms_per_chain": ms,
                "ms_single_gpu_chain": t_single, "gflops": flops / ms * 1e-6, "moved_bytes_rank0": moved_bytes}
        if variant == "exchange":
            x_ms = timed(lambda: xchg.run(C, publish=False), args.iters)
            line["exchange_ms"] = x_ms
            line["exchange_GBps_rank0"] = moved_bytes / x_ms * 1e-6
        if variant == "nccl":
            x_ms = timed(lambda: sharding.redistribute_blocks(C, slices, produced_by, needed_by), max(args.iters // 4, 2))
            line["exchange_ms"] = x_ms
        assert worst[0] <= 1e-12 and worst[1] <= 1e-12, (variant, worst)
        lines.append(line)
        if rank == 0:
            print(json.dumps(line), flush=True)
    # bandwidth of the block exchange when EVERY block changes owner (the recorded chain moves few blocks at 2 ranks: the LPT
    # owners of both contractions alternate with the charge and the charges of l and r differ by even numbers), and the cost
    # of a GEMM epilogue that scatters every result block into the neighbour's HBM instead of the local one
    nxt = [(p + 1) % world for p in produced_by]
    swap = sharding.PeerExchange(arena, C, slices, produced_by, nxt)
    contract(sh1, A, F, out=C)
    swap_ms = timed(lambda: swap.run(C, publish=False), args.iters)
    nccl_ms = timed(lambda: sharding.redistribute_blocks(C, slices, produced_by, nxt), max(args.iters // 4, 2))
    local_ms = timed(lambda: contract(sh1, A, F, out=C), args.iters)
    all_remote = sharding.unmerge_dst_shift(sh1, slices, nxt, rank, shifts)
    remote_ms = timed(lambda: contract(sh1, A, F, out=C, dst_shift=all_remote), args.iters)
    line = {"tool": "multigpu_chain", "case": args.case, "dtype": args.dtype, "n_gpus": world, "variant": "swap_all_blocks",
            "moved_bytes_rank0": swap.elements * isz, "peer_exchange_ms": swap_ms, "peer_exchange_GBps_rank0": swap.elements * isz / swap_ms * 1e-6,
            "nccl_sendrecv_ms": nccl_ms, "nccl_GBps_rank0": swap.elements * isz / nccl_ms * 1e-6,
            "contraction1_local_epilogue_ms": local_ms, "contraction1_remote_epilogue_ms": remote_ms}
    lines.append(line)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if rank == 0 and args.out:
        with open(args.out, "a") as f:
            for line in lines:
                f.write(json.dumps(line) + "\n")
    arena.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
