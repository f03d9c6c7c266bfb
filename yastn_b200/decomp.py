"""Sector-parallel matrix decompositions (the block loops of svd / svdvals / eigh / qr).

The reference decomposes a block-sparse matrix with a Python loop over charge sectors, one cuSOLVER call per sector on
the current stream and one host synchronisation per call (``info`` check):

    svd      yastn/backend/_backend_torch_backwards.py:26-39  (SVDGESDD.forward, backend/linalg/torch_svd_gesdd.py:15-21,
             ``torch.linalg.svd(A, full_matrices=..., driver='gesvd')`` on CUDA)
    svdvals  yastn/backend/backend_torch.py:322-327
    eigh     yastn/backend/backend_torch.py:365-380
    qr       yastn/backend/backend_torch.py:475-484

After the contraction kernels these loops are what a DMRG / CTMRG sweep waits for (72 % of a D=4096 Hubbard sweep,
profiles/e2e_targets_r01.jsonl): a sector of a few hundred rows keeps a handful of SMs busy and the host blocks on every
sector.  Sectors are independent, so here they are dealt — heaviest first — to a small pool of host threads, each
owning a CUDA stream and (through torch's per-thread handle pool) its own cuSOLVER handle: the factorisations of
different sectors overlap on the device and the per-sector host synchronisations overlap with each other.  Every sector
is still factorised by the same library routine with the same arguments as in the reference, so the results are
bit-identical to the reference loop on the same device; only the schedule changes.

Inputs that require grad keep the reference's autograd implementations (the pool is a forward-only schedule), CPU
tensors keep the reference's loop.
"""
import ctypes
import os
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

# Sector streams.  Unset: 4 for complex128, 8 for float64 — with gesvdp a D=4096 complex128 sweep takes 12.0-12.3 s at 4 streams,
# 12.6-14.0 at 8, 14.4 at 12, 15.6 at 16, 16.1 at 2; the same sweep in float64 9.8 s at 8 and 10.8 at 4; CTMRG chi=256 (float64)
# 0.74-0.86 s at 8, 0.83-0.93 at 4 (profiles/decomp_workers_r02.jsonl).  Round 1's gesvd wanted 8 everywhere: the slower
# routine left more of the GPU idle per stream.
_WORKERS = int(os.environ["YASTN_B200_DECOMP_WORKERS"]) if "YASTN_B200_DECOMP_WORKERS" in os.environ else None


def _workers(complex_data=False):
    if _WORKERS is not None:
        return _WORKERS
    w = 4 if complex_data else 8
    # SPMD (yastn_b200.spmd): every rank factorises only its share of the sectors, and every stream thread spins on the host
    # while cuSOLVER synchronises — 8 ranks x 4 threads starve the replicated Python programs of cores (the D=4096 sweep took
    # 16.8-22.8 s on 8 GPUs against 8.7 s on 4).  The node keeps about two ranks' worth of stream threads in total.
    world = _spmd["world"] if _spmd["all_reduce"] is not None else 1
    return w if world <= 2 else max(1, (2 * w) // world)


# Per-sector SVD routine.  "gesvd" is the reference's choice on CUDA (torch_svd_gesdd.py:17) and gives U, S, Vh bit-identical to
# the stock torch backend.  The default "gesvdp" sends sectors of at least _SVDP_MIN rows and columns to cuSOLVER's
# polar-decomposition SVD (yastn_b200/cusolver_svdp.py): 26 vs 87 ms for a 652 x 652 complex128 sector, 5.6 vs 15.4 ms at 163,
# singular values of a spectrum graded over ten decades accurate to 4e-16 * S_max (gesvd: 5e-13), profiles/svd_probe_r02.jsonl.
_SVD_DRIVER = os.environ.get("YASTN_B200_SVD_DRIVER", "gesvdp")
_SVDP_MIN = int(os.environ.get("YASTN_B200_SVDP_MIN", "48"))
_MIN_SECTORS = 2          # a single sector has nothing to overlap with
_pools = {}
_pools_lock = threading.Lock()
_stats = {"parallel_calls": 0, "sectors": 0, "serial_calls": 0}
_THREADS_WITHOUT_STREAMS = False   # test hook: exercise the thread pool on CPU tensors (tests/test_decomp.py)


class _Pool:
    """Host threads with one CUDA stream each, per device."""

    def __init__(self, device, workers):
        self.device = device
        self.workers = workers
        self.tls = threading.local()
        self.streams = []
        self.lock = threading.Lock()
        self.exe = ThreadPoolExecutor(max_workers=workers, thread_name_prefix=f"yb_decomp{device.index}")

    def stream(self):
        if self.device.type != "cuda":
            return None
        s = getattr(self.tls, "stream", None)
        if s is None:
            s = torch.cuda.Stream(self.device)
            self.tls.stream = s
            with self.lock:
                self.streams.append(s)
        return s


def _warm_linalg(device):
    """torch loads its CUDA linalg library lazily on the first linalg call and that first call must not race with another
    one ("lazy wrapper should be called at most once"): make it here, on the caller's thread, before any worker runs."""
    x = torch.eye(2, dtype=torch.float64, device=device)
    torch.linalg.svd(x, full_matrices=False, driver=_torch_driver())
    torch.linalg.svdvals(x)
    torch.linalg.eigh(x)
    torch.linalg.qr(x)
    torch.cuda.synchronize(device)


def _torch_driver():
    return "gesvd" if _SVD_DRIVER == "gesvdp" else _SVD_DRIVER


def set_svd_driver(name):
    """"gesvdp" (default: polar-decomposition SVD for large sectors), "gesvd" (the reference's routine, bit-identical results),
    or any other torch.linalg.svd driver."""
    global _SVD_DRIVER
    _SVD_DRIVER = name


def _sector_svd(A, fullrank_uv):
    if _SVD_DRIVER == "gesvdp" and not fullrank_uv and A.is_cuda and min(A.shape) >= _SVDP_MIN:
        from . import cusolver_svdp
        if cusolver_svdp.available():
            U, S, Vh, err = cusolver_svdp.svd(A)
            if err <= 1e-10:          # gesvdp reports a perturbation it had to apply to an ill-conditioned input
                with _pools_lock:
                    _stats["svdp_sectors"] = _stats.get("svdp_sectors", 0) + 1
                return U, S, Vh
    return torch.linalg.svd(A, full_matrices=fullrank_uv, driver=_torch_driver() if A.is_cuda else None)


def _pool(device, workers):
    with _pools_lock:
        key = (device.type, device.index, workers)
        p = _pools.get(key)
        if p is None:
            if device.type == "cuda" and not any(k[:2] == key[:2] for k in _pools):
                _warm_linalg(device)
            p = _Pool(device, workers)
            _pools[key] = p
        return p


def set_workers(n):
    """Number of sector streams (1 = the reference's serial schedule; None = the default: 4 for complex128, 8 for float64)."""
    global _WORKERS
    _WORKERS = None if n is None else max(1, int(n))


def stats():
    return dict(_stats)


def run_sectors(fn, recs, costs, device, complex_data=False):
    """Call ``fn(rec)`` for every record, concurrently over the sector pool of ``device`` (heaviest first, dynamic
    dealing).  Returns when every call has been *issued* and the caller's current stream has been made to wait for all
    of them, i.e. with torch's usual stream semantics for the caller."""
    n = len(recs)
    if n == 0:
        return
    cuda = device.type == "cuda"
    workers = _workers(complex_data)
    if not (cuda or _THREADS_WITHOUT_STREAMS) or workers <= 1 or n < _MIN_SECTORS:
        _stats["serial_calls"] += 1
        for rec in recs:
            fn(rec)
        return
    pool = _pool(device, workers)
    order = sorted(range(n), key=lambda i: -costs[i])
    grad = torch.is_grad_enabled()
    if cuda:
        main = torch.cuda.current_stream(device)
        start = torch.cuda.Event()
        start.record(main)

    def task(i):
        if not cuda:
            with torch.set_grad_enabled(grad):
                return fn(recs[i])
        s = pool.stream()
        with torch.cuda.device(device), torch.cuda.stream(s), torch.set_grad_enabled(grad):
            s.wait_event(start)
            fn(recs[i])

    futures = [pool.exe.submit(task, i) for i in order]
    err = None
    for f in futures:
        try:
            f.result()
        except Exception as e:   # keep draining so that no worker is left writing into the outputs
            err = err or e
    if cuda:
        with pool.lock:
            streams = list(pool.streams)
        for s in streams:
            main.wait_stream(s)
    _stats["parallel_calls"] += 1
    _stats["sectors"] += n
    if err is not None:
        raise err


# ---- small sectors: ONE launch of the batched one-sided Jacobi kernel (csrc/yb_svd.cu) -------------------------------------
_JACOBI_MAX = int(os.environ.get("YASTN_B200_JACOBI_MAX", "64"))     # sectors up to this many rows and columns; 0 disables
_JACOBI_SWEEPS = 30
_jacobi_plans = {}


def set_jacobi_max(n):
    global _JACOBI_MAX
    _JACOBI_MAX = int(n)


def _jacobi_svd(data, meta, small, Udata, Sdata, Vhdata, vectors=True):
    """Factorise the sectors ``meta[i], i in small`` in one launch, straight into the output tensors.  Returns the indices that
    must be redone by the library routine (not converged / exactly singular: the kernel reports, it never guesses)."""
    from . import plans
    dev = data.device
    key = (id(meta), tuple(small) if len(small) < len(meta) else None, data.dtype, dev.index, vectors)
    ent = _jacobi_plans.get(key)
    if ent is None or ent[0] is not meta:
        if vectors:
            recs = [[meta[i][0][0], meta[i][1][0], meta[i][1][1], meta[i][2][0], meta[i][4][0], meta[i][5][0]] for i in small]
        else:      # svdvals meta: (sl, D, slU, DU, slS, ...)
            recs = [[meta[i][0][0], meta[i][1][0], meta[i][1][1], 0, meta[i][4][0], 0] for i in small]
        if len(_jacobi_plans) > 4096:
            _jacobi_plans.clear()
        ent = (meta, plans.SvdPlan(np.array(recs, dtype=np.int64), 16 if data.is_complex() else 8, dev.index))
        _jacobi_plans[key] = ent
    if data.is_conj():
        data = data.resolve_conj()
    data = data if data.is_contiguous() else data.contiguous()
    status = torch.empty(len(small), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        ent[1].run(data.data_ptr(), Udata.data_ptr() if vectors else None, Sdata.data_ptr(), Vhdata.data_ptr() if vectors else None,
                   status.data_ptr(), _JACOBI_SWEEPS, vectors, st)
    with _pools_lock:
        _stats["jacobi_calls"] = _stats.get("jacobi_calls", 0) + 1
        _stats["jacobi_sectors"] = _stats.get("jacobi_sectors", 0) + len(small)
    bad = status.nonzero().reshape(-1).tolist()        # one host synchronisation per block matrix (the reference has one per sector)
    return [small[i] for i in bad]


# ---- several GPUs (yastn_b200.spmd): the sectors of one decomposition are dealt to the ranks, every rank factorises its share
# into zeroed outputs and one all-reduce per output completes them everywhere (each entry is written by exactly one rank)
_spmd = {"all_reduce": None, "rank": 0, "world": 1}
_SPMD_MIN_COST = float(os.environ.get("YASTN_B200_SPMD_DECOMP_MIN", "5e7"))


def set_spmd(all_reduce, rank, world):
    _spmd.update(all_reduce=all_reduce, rank=rank, world=world)


def _my_sectors(costs):
    """Indices of the sectors this rank factorises (LPT on the cost, identical on every rank), or None: not sharded."""
    if _spmd["all_reduce"] is None or _spmd["world"] < 2 or len(costs) < 2 or sum(costs) < _SPMD_MIN_COST:
        return None
    import heapq
    heap = [(0, r) for r in range(_spmd["world"])]
    mine = []
    for i in sorted(range(len(costs)), key=lambda i: (-costs[i], i)):
        load, r = heapq.heappop(heap)
        if r == _spmd["rank"]:
            mine.append(i)
        heapq.heappush(heap, (load + costs[i], r))
    with _pools_lock:
        _stats["spmd_calls"] = _stats.get("spmd_calls", 0) + 1
    return sorted(mine)


def _svd_cost(D):
    m, n = D
    return m * n * min(m, n)


def make(stock):
    """The decomposition functions of a backend module, given the reference's own module ``stock`` (used for inputs that
    require grad or live on the CPU)."""

    def _defer(*tensors):
        return any((not (t.is_cuda or _THREADS_WITHOUT_STREAMS)) or (torch.is_grad_enabled() and t.requires_grad) for t in tensors)

    def svd(data, meta, sizes, fullrank_uv=False, ad_decomp_reg=1.0e-12, diagnostics=None, **kwargs):
        if _defer(data):
            return stock.svd(data, meta, sizes, fullrank_uv=fullrank_uv, ad_decomp_reg=ad_decomp_reg, diagnostics=diagnostics, **kwargs)
        real_dtype = data.real.dtype if data.is_complex() else data.dtype
        mine = None if fullrank_uv else _my_sectors([_svd_cost(m[1]) for m in meta])
        alloc = torch.empty if mine is None else torch.zeros
        Udata = alloc(sizes[0], dtype=data.dtype, device=data.device)
        Sdata = alloc(sizes[1], dtype=real_dtype, device=data.device)
        Vhdata = alloc(sizes[2], dtype=data.dtype, device=data.device)

        def one(rec):
            sl, D, slU, DU, slS, slV, DV = rec
            U, S, Vh = _sector_svd(data[sl[0]:sl[1]].view(D), fullrank_uv)
            Udata[slU[0]:slU[1]].view(DU).copy_(U)
            Sdata[slS[0]:slS[1]].copy_(S)
            Vhdata[slV[0]:slV[1]].view(DV).copy_(Vh)
        rest = list(range(len(meta))) if mine is None else mine
        if _JACOBI_MAX > 0 and not fullrank_uv and data.dtype in (torch.float64, torch.complex128):
            small = [i for i in rest if max(meta[i][1]) <= min(_JACOBI_MAX, 64) and min(meta[i][1]) >= 1]
            if small:
                redo = _jacobi_svd(data, meta, small, Udata, Sdata, Vhdata)
                done = set(small) - set(redo)
                rest = [i for i in rest if i not in done]
        recs = meta if len(rest) == len(meta) else [meta[i] for i in rest]
        run_sectors(one, recs, [_svd_cost(m[1]) for m in recs], data.device, data.is_complex())
        if mine is not None:
            for t in (Udata, Sdata, Vhdata):
                _spmd["all_reduce"](t)
        return Udata, Sdata, Vhdata

    def svdvals(data, meta, sizeS, **kwargs):
        if _defer(data):
            return stock.svdvals(data, meta, sizeS, **kwargs)
        real_dtype = data.real.dtype if data.is_complex() else data.dtype
        Sdata = torch.zeros((sizeS,), dtype=real_dtype, device=data.device)

        def one(rec):
            sl, D, slS = rec[0], rec[1], rec[4]
            Sdata[slS[0]:slS[1]].copy_(torch.linalg.svdvals(data[sl[0]:sl[1]].view(D)))
        run_sectors(one, meta, [_svd_cost(m[1]) for m in meta], data.device, data.is_complex())
        return Sdata

    def eigh(data, meta=None, sizes=(1, 1), order_by_magnitude=False, ad_decomp_reg=1.0e-12):
        if meta is None or order_by_magnitude or _defer(data):
            return stock.eigh(data, meta, sizes, order_by_magnitude=order_by_magnitude, ad_decomp_reg=ad_decomp_reg)
        real_dtype = data.real.dtype if data.is_complex() else data.dtype
        Sdata = torch.zeros((sizes[0],), dtype=real_dtype, device=data.device)
        Udata = torch.zeros((sizes[1],), dtype=data.dtype, device=data.device)

        def one(rec):
            sl, D, slU, DU, slS = rec
            S, U = torch.linalg.eigh(data[sl[0]:sl[1]].view(D))
            Sdata[slS[0]:slS[1]].copy_(S)
            Udata[slU[0]:slU[1]].view(DU).copy_(U)
        mine = _my_sectors([m[1][0] ** 3 for m in meta])
        recs = meta if mine is None else [meta[i] for i in mine]
        run_sectors(one, recs, [m[1][0] ** 3 for m in recs], data.device, data.is_complex())
        if mine is not None:
            _spmd["all_reduce"](Sdata)
            _spmd["all_reduce"](Udata)
        return Sdata, Udata

    def qr(data, meta, sizes):
        if _defer(data):
            return stock.qr(data, meta, sizes)
        Qdata = torch.zeros((sizes[0],), dtype=data.dtype, device=data.device)
        Rdata = torch.zeros((sizes[1],), dtype=data.dtype, device=data.device)

        def one(rec):
            sl, D, slQ, DQ, slR, DR = rec
            Q, R = torch.linalg.qr(data[sl[0]:sl[1]].view(D))
            d = R.diagonal()
            sR = torch.sign(d.real if d.is_complex() else d)
            sR[sR == 0] = 1
            Qdata[slQ[0]:slQ[1]].view(DQ).copy_(Q * sR)       # positive diagonal of R
            Rdata[slR[0]:slR[1]].view(DR).copy_(sR.reshape([-1, 1]) * R)
        mine = _my_sectors([_svd_cost(m[1]) for m in meta])
        recs = meta if mine is None else [meta[i] for i in mine]
        run_sectors(one, recs, [_svd_cost(m[1]) for m in recs], data.device, data.is_complex())
        if mine is not None:
            _spmd["all_reduce"](Qdata)
            _spmd["all_reduce"](Rdata)
        return Qdata, Rdata

    return {"svd": svd, "svdvals": svdvals, "eigh": eigh, "qr": qr}
