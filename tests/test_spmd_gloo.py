"""SPMD execution of an unmodified YASTN program on 2 ranks (gloo on CPU, device side = the numpy table interpreter of
tests/cpu_shim.py): yastn_b200.spmd shards every fuse_to_matrix contraction above ``min_flops`` by FLOP-balanced row panels
and the sectors of svd / eigh / qr by cost, completes the results with one all-reduce, and every rank must then hold the same
bits as a single-process run holds to rounding — tensordots directly, and a 2-site DMRG end to end (same energy, same bond
dimensions).  The kernels themselves and NCCL are covered by tests/test_multigpu_gpu.py on the GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yastn_loader import load_yastn

if load_yastn() is None:
    pytest.skip("yastn not importable (no baseline/_ref, no reference checkout)", allow_module_level=True)

WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup(spmd_on):
    import cpu_shim
    from yastn_b200 import yastn_backend, spmd, decomp
    yastn = load_yastn()
    cpu_shim.install()
    spmd._bk_usable = lambda d: True               # CPU tensors stand in for device tensors under the shim
    decomp._THREADS_WITHOUT_STREAMS = True         # test hook: the sector schedule (and its sharding) on CPU tensors
    decomp.set_jacobi_max(0)                       # the batched Jacobi kernel is device code
    yastn_backend.enable_fused_tensordot()
    if spmd_on:
        decomp._SPMD_MIN_COST = 0.0
        spmd.enable(min_flops=0.0)
    return yastn, yastn_backend, spmd


def _tensordots(yastn, cfg):
    cfg.backend.random_seed(7)
    out = []
    for dtype in ("float64", "complex128"):
        a = yastn.rand(config=cfg, s=(-1, 1, 1, -1), t=((-1, 1, 2), (-1, 1, 2), (-1, 1, 2), (-1, 1, 2)),
                       D=((1, 2, 3), (4, 5, 6), (7, 8, 9), (10, 11, 12)), dtype=dtype)
        b = yastn.rand(config=cfg, s=(1, -1, 1), t=((-1, 1, 2), (-1, 1, 2), (-1, 0, 1)), D=((1, 2, 3), (4, 5, 6), (10, 7, 11)), dtype=dtype)
        for axes in ((0, 0), ((0, 1), (0, 1)), (1, 1), ((), ())):
            out.append(yastn.tensordot(a, b, axes=axes))
        out.append(yastn.tensordot(a, a.conj(), axes=((0, 1, 2), (0, 1, 2))))
        u, s, v = yastn.svd(a, axes=((0, 1), (2, 3)), sU=1)
        out.append(s)
        out.append(u @ s @ v)
        q, r = yastn.qr(a, axes=((0, 1), (2, 3)))
        out.append(q @ r)
    return out


def _dmrg(yastn, backend):
    import yastn.tn.mps as mps
    N = 6
    ops = yastn.operators.Spin12(sym="U1", backend=backend, default_device="cpu", tensordot_policy="fuse_to_matrix")
    ops.random_seed(seed=0)
    I = mps.product_mpo(ops.I(), N)
    terms = []
    for n in range(N - 1):
        terms += [mps.Hterm(1.0, [n, n + 1], [ops.sz(), ops.sz()]), mps.Hterm(0.5, [n, n + 1], [ops.sp(), ops.sm()]),
                  mps.Hterm(0.5, [n, n + 1], [ops.sm(), ops.sp()])]
    H = mps.generate_mpo(I, terms)
    psi = mps.random_mps(I, n=0, D_total=6)
    out = mps.dmrg_(psi, H, method="2site", max_sweeps=3, opts_svd={"tol": 1e-10, "D_total": 12})
    return float(out.energy), psi.get_bond_dimensions()


def _worker(rank, port, ret, world=WORLD):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        yastn, yastn_backend, spmd = _setup(True)
        cfg = yastn.make_config(sym="U1", backend=yastn_backend.module(), tensordot_policy="fuse_to_matrix", default_device="cpu")
        res = _tensordots(yastn, cfg)
        st = spmd.stats()
        energy, bonds = _dmrg(yastn, yastn_backend.module())
        from yastn_b200 import decomp
        ret[rank] = {"data": [t._data.resolve_conj().numpy().copy() for t in res], "struct": [(t.struct, t.slices) for t in res],
                     "stats": st, "stats_end": spmd.stats(), "energy": energy, "bonds": bonds, "decomp": decomp.stats()}
        spmd.disable()
    finally:
        dist.destroy_process_group()


def test_spmd_two_ranks_match_single_process():
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(port, ret), nprocs=WORLD, join=True)
        r0, r1 = ret[0], ret[1]
    # single process, same program
    yastn, yastn_backend, spmd = _setup(False)
    try:
        cfg = yastn.make_config(sym="U1", backend=yastn_backend.module(), tensordot_policy="fuse_to_matrix", default_device="cpu")
        ref = _tensordots(yastn, cfg)
        energy, bonds = _dmrg(yastn, yastn_backend.module())
    finally:
        import cpu_shim
        from yastn_b200 import decomp
        decomp._THREADS_WITHOUT_STREAMS = False
        decomp.set_jacobi_max(64)
        yastn_backend.disable_fused_tensordot()
        cpu_shim.uninstall()
    assert r0["stats"]["sharded"] >= 10 and r0["stats"]["sharded"] == r1["stats"]["sharded"]
    assert r0["decomp"].get("spmd_calls", 0) >= 2
    for k, t in enumerate(ref):
        x = t._data.resolve_conj().numpy()
        assert (t.struct, t.slices) == r0["struct"][k] == r1["struct"][k]
        assert np.array_equal(r0["data"][k], r1["data"][k]), k            # the ranks hold the same bits
        if k % 9 in (5,):                       # singular values
            assert np.abs(r0["data"][k] - x).max() <= 1e-12 * max(np.abs(x).max(), 1)
        else:
            assert np.linalg.norm(r0["data"][k] - x) <= 1e-12 * max(np.linalg.norm(x), 1e-300), k
    assert r0["energy"] == r1["energy"] and r0["bonds"] == r1["bonds"] == bonds
    assert abs(r0["energy"] - energy) <= 1e-12 * abs(energy)
    assert r0["stats_end"]["sharded"] > r0["stats"]["sharded"]


def _fermion_worker(rank, port, ret, world):
    """Z2 spinless fermions (swap gates -> negate_blocks, odd-parity blocks) on `world` ranks."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        yastn, yastn_backend, spmd = _setup(True)
        ret[rank] = _fermion_dmrg(yastn, yastn_backend.module()) + (spmd.stats()["sharded"],)
        spmd.disable()
    finally:
        dist.destroy_process_group()


def _fermion_dmrg(yastn, backend):
    import yastn.tn.mps as mps
    N = 6
    ops = yastn.operators.SpinlessFermions(sym="Z2", backend=backend, default_device="cpu", tensordot_policy="fuse_to_matrix")
    ops.random_seed(seed=0)
    I = mps.product_mpo(ops.I(), N)
    terms = []
    for n in range(N - 1):
        terms += [mps.Hterm(-1.0, [n, n + 1], [ops.cp(), ops.c()]), mps.Hterm(-1.0, [n + 1, n], [ops.cp(), ops.c()])]
    for n in range(N):
        terms.append(mps.Hterm(0.2 * ((n % 3) - 1), [n], [ops.n()]))
    H = mps.generate_mpo(I, terms)
    psi = mps.random_mps(I, n=1, D_total=6)
    out = mps.dmrg_(psi, H, method="2site", max_sweeps=3, opts_svd={"tol": 1e-10, "D_total": 10})
    return float(out.energy), psi.get_bond_dimensions()


def test_spmd_three_ranks_fermions_match_single_process():
    """An odd number of ranks (uneven row panels, sectors dealt 3 ways) on a fermionic model: every rank ends with the same
    energy and bond dimensions as the single-process run."""
    world = 3
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_fermion_worker, args=(port, ret, world), nprocs=world, join=True)
        res = [ret[r] for r in range(world)]
    yastn, yastn_backend, spmd = _setup(False)
    try:
        energy, bonds = _fermion_dmrg(yastn, yastn_backend.module())
    finally:
        import cpu_shim
        from yastn_b200 import decomp
        decomp._THREADS_WITHOUT_STREAMS = False
        decomp.set_jacobi_max(64)
        yastn_backend.disable_fused_tensordot()
        cpu_shim.uninstall()
    assert all(r[0] == res[0][0] and r[1] == res[0][1] for r in res)          # identical on every rank
    assert res[0][2] > 0 and res[0][1] == bonds
    assert abs(res[0][0] - energy) <= 1e-12 * abs(energy)
