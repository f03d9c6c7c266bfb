#!/usr/bin/env python
"""The exchange step of the SPMD tensordot in isolation (torchrun, N GPUs): every rank owns 1/N of a result of S bytes and
all ranks need all of it.  Variants: the peer-arena push (one copy-kernel launch storing into every peer's slot + publish),
NCCL broadcasts of the panels, NCCL all-gather, NCCL all-reduce of the zero-padded result.  Device-timed (CUDA events, barrier
before every repetition, max over ranks); prints one JSON line per size on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    from yastn_b200 import spmd, peer
    spmd._state.update(rank=rank, world=world, group=None)
    sizes = [int(x) for x in sys.argv[1:]] or [64 << 20, 512 << 20, 2048 << 20]
    arena = peer.PeerArena(2 * max(sizes) + (1 << 20), device=dev)
    for nbytes in sizes:
        n = nbytes // 16
        runs = tuple(((n * r // world, n * (r + 1) // world),) for r in range(world))
        slot = arena.view(0, n, torch.complex128)
        plain = torch.zeros(n, dtype=torch.complex128, device=dev)
        lo, hi = runs[rank][0]
        plain[lo:hi] = rank + 1
        slot.copy_(plain)
        flat = torch.view_as_real(plain).reshape(-1)
        pieces = [torch.empty((runs[r][0][1] - runs[r][0][0]) * 2, dtype=torch.float64, device=dev) for r in range(world)]

        def push():
            spmd._exchange_arena(slot, arena, 0, 0, runs)

        def bcast():
            for r in range(world):
                a, b = runs[r][0]
                dist.broadcast(flat[2 * a:2 * b], src=r)

        def gather():
            a, b = runs[rank][0]
            dist.all_gather(pieces, flat[2 * a:2 * b].contiguous()) if len({p.numel() for p in pieces}) == 1 else bcast()

        def reduce():
            dist.all_reduce(flat)
        out = {"bytes": nbytes, "world": world}
        for name, fn in (("peer_push", push), ("nccl_broadcast", bcast), ("nccl_all_gather", gather), ("nccl_all_reduce", reduce)):
            for _ in range(2):
                fn()
            ms = []
            for _ in range(5):
                torch.cuda.synchronize()
                dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms.append(float(t))
            best = min(ms)
            out[name + "_ms"] = round(best, 3)
            out[name + "_GBps_per_rank_in"] = round(nbytes * (world - 1) / world / (best * 1e-3) * 1e-9, 1)
        # the pushed slot holds every rank's panel
        push()
        torch.cuda.synchronize()
        dist.barrier()
        ok = all(bool((slot[runs[r][0][0]:runs[r][0][1]] == r + 1).all()) for r in range(world))
        out["peer_push_correct"] = ok
        if rank == 0:
            print(json.dumps(out), flush=True)
    arena.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
