"""Drop-in replacements for the hot functions of ``yastn.backend.backend_torch`` on B200.

Same names, argument meaning and return contract as the reference backend
(yastn/backend/backend_torch.py:549-593; loops in yastn/backend/_backend_torch_backwards.py):

    transpose_and_merge(data, order, meta_new, meta_mrg, Dsize) -> Tensor[Dsize]
    dot(Adata, Bdata, meta_dot, Dsize)                            -> Tensor[Dsize]
    unmerge(data, meta)                                           -> Tensor[len(data)]
    transpose(data, axes, meta_transpose)                         -> Tensor[len(data)]
    transpose_dot_sum(Adata, Bdata, meta_dot, Areshape, Breshape, Aorder, Border, Dsize) -> Tensor[Dsize]

Inputs are borrowed and never mutated, outputs are freshly allocated 1-D contiguous CUDA tensors, every
function is differentiable (explicit backward, like the reference's autograd.Functions) and each forward is
ONE kernel launch through the C ABI (include/yastn_b200.h) instead of a Python loop over blocks.
Only CUDA float64 / complex128 tensors are accepted: there is no CPU or eager-torch fallback.

Use with YASTN through ``yastn_b200.yastn_backend`` (``module()`` for ``yastn.make_config(backend=...)`` or
``activate()`` which rebinds the five functions on ``yastn.backend.backend_torch``).
"""
import ctypes

import torch

from . import _lib, plans

BACKEND_ID = "torch"   # serialised tensors stay combinable with the stock torch backend

_CACHE = plans.PlanCache()
_DTYPE_CODE = {torch.float64: _lib.YB_F64, torch.complex128: _lib.YB_C128}
_ITEMSIZE = {torch.float64: 8, torch.complex128: 16}


def clear_plan_cache():
    _CACHE.clear()
    _BS_CACHE.clear()
    import sys
    chain = sys.modules.get(__package__ + ".chain")
    if chain is not None:          # recorded chains keep their plans alive: dropped together with the plan cache
        chain.clear()


def plan_cache_stats():
    return {"hits": _CACHE.hits, "misses": _CACHE.misses, "size": len(_CACHE._d)}


def _check(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"yastn_b200.{name}: expected a CUDA tensor (no CPU fallback), got {type(t).__name__}"
                        f"{'' if not isinstance(t, torch.Tensor) else ' on ' + str(t.device)}")
    if t.dtype not in _DTYPE_CODE:
        raise TypeError(f"yastn_b200.{name}: dtype {t.dtype} not supported (float64 / complex128 only)")
    if t.dim() != 1:
        raise ValueError(f"yastn_b200.{name}: data must be 1-D")


def _raw(t):
    """(tensor to keep alive, conj flag): contiguous storage view of t; lazy conj is resolved inside the kernels."""
    conj = t.is_conj()
    if conj:
        t = t.conj()          # flips the bit back: a view of the physical (un-conjugated) storage
    if not t.is_contiguous():
        t = t.contiguous()
    return t, conj


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _on_device(dev, launch):
    """Run ``launch(stream)`` with ``dev`` current (plans and streams are per device)."""
    if torch.cuda.current_device() != dev.index:
        with torch.cuda.device(dev):
            launch(_stream(dev))
    else:
        launch(_stream(dev))


_recorder = None        # yastn_b200.chain._Recorder while a chain of launches is being recorded
_gemm_hook = None       # tools/dmrg_bench.py --gemm-roofline: called with None before a grouped-GEMM launch, with its token after


def _run_copy(plan, src, dst, zero):
    raw, conj = _raw(src)
    flags = (_lib.YB_COPY_ZERO_DST if zero else 0) | (_lib.YB_COPY_CONJ if conj else 0)
    if _recorder is not None:
        _recorder.copy(plan, raw, dst, flags)
    _on_device(src.device, lambda st: plan.run(raw.data_ptr(), dst.data_ptr(), dst.numel(), flags, st))


def _run_gemm(plan, A, B, C, conj_a=False, conj_b=False):
    ra, ca = _raw(A)
    rb, cb = _raw(B)
    flags = (_lib.YB_GEMM_CONJ_A if (ca != conj_a) else 0) | (_lib.YB_GEMM_CONJ_B if (cb != conj_b) else 0)
    if _recorder is not None:
        _recorder.gemm(plan, ra, rb, C, flags)
    if _gemm_hook is not None:        # measurement tools: CUDA events right around the launch (plan creation stays outside)
        token = _gemm_hook(None)
        _on_device(C.device, lambda st: plan.run(ra.data_ptr(), rb.data_ptr(), C.data_ptr(), flags, st))
        _gemm_hook(token)
        return
    _on_device(C.device, lambda st: plan.run(ra.data_ptr(), rb.data_ptr(), C.data_ptr(), flags, st))


# -------------------------------------------------------------------------------------------------
# plan lookup (cached on the identity of YASTN's lru-cached meta objects)
# -------------------------------------------------------------------------------------------------

def _merge_plans(data, order, meta_new, meta_mrg, Dsize):
    key = ("mrg", id(meta_mrg), tuple(order), Dsize, len(meta_new), data.dtype, data.device.index)

    def build():
        # `covered` counts the zero-fill records of cells no source block covers: the destination needs no memset then
        recs, rank, covered = plans.merge_records(order, meta_new, meta_mrg)
        fwd = plans.CopyPlan(recs, rank, _ITEMSIZE[data.dtype], data.device.index, covered)
        src_read = int(recs[recs[:, 0] != plans.SRC_ZERO][:, 2:2 + rank].prod(axis=1).sum()) if recs.shape[0] else 0
        return {"fwd": fwd, "recs": recs, "rank": rank, "bwd": None, "src_read": src_read}
    return _CACHE.get(key, meta_mrg, build)


def _unmerge_plans(data, meta):
    key = ("unm", id(meta), data.dtype, data.device.index)

    def build():
        recs, rank = plans.unmerge_records(meta)
        return {"fwd": plans.CopyPlan(recs, rank, _ITEMSIZE[data.dtype], data.device.index), "recs": recs, "rank": rank, "bwd": None}
    return _CACHE.get(key, meta, build)


def _transpose_plans(data, axes, meta):
    # consume_transpose (yastn/tensor/_single.py:316-346) is NOT lru-cached: it hands over a fresh, equal meta tuple on
    # every call, so this plan is keyed by the meta's CONTENT (hash + equality of the nested int tuples), not by identity
    key = ("trn", meta, tuple(axes), data.dtype, data.device.index)

    def build():
        recs, rank = plans.transpose_records(axes, meta)
        return {"fwd": plans.CopyPlan(recs, rank, _ITEMSIZE[data.dtype], data.device.index), "recs": recs, "rank": rank, "bwd": None}
    return _CACHE.get(key, None, build)


def _bwd_copy_plan(ent, dtype, device):
    if ent["bwd"] is None:
        ent["bwd"] = plans.CopyPlan(plans.reverse_records(ent["recs"], ent["rank"]), ent["rank"], _ITEMSIZE[dtype], device)
    return ent["bwd"]


def _dot_plans(meta_dot, dtype, device):
    key = ("dot", id(meta_dot), dtype, device)

    def build():
        problems, segments = plans.dot_tables(meta_dot)
        return {"fwd": plans.GemmPlan(problems, segments, _DTYPE_CODE[dtype], device), "bwd": None}
    return _CACHE.get(key, meta_dot, build)


# -------------------------------------------------------------------------------------------------
# autograd functions (forward = one launch; backward = adjoint plan)
# -------------------------------------------------------------------------------------------------

class _TransposeAndMerge(torch.autograd.Function):
    @staticmethod
    def forward(data, order, meta_new, meta_mrg, Dsize):
        ent = _merge_plans(data, order, meta_new, meta_mrg, Dsize)
        out = torch.empty(Dsize, dtype=data.dtype, device=data.device)
        _run_copy(ent["fwd"], data, out, zero=ent["fwd"].covered < Dsize)
        return out

    @staticmethod
    def setup_context(ctx, inputs, output):
        data, order, meta_new, meta_mrg, Dsize = inputs
        ctx.args = (order, meta_new, meta_mrg, Dsize)
        ctx.n_src = data.numel()

    @staticmethod
    def backward(ctx, grad):
        order, meta_new, meta_mrg, Dsize = ctx.args
        grad = grad.resolve_conj() if grad.is_conj() else grad
        ent = _merge_plans(grad, order, meta_new, meta_mrg, Dsize)
        plan = _bwd_copy_plan(ent, grad.dtype, grad.device.index)
        out = torch.empty(ctx.n_src, dtype=grad.dtype, device=grad.device)
        _run_copy(plan, grad, out, zero=ent["src_read"] < ctx.n_src)      # source blocks that took no part keep a zero gradient
        return out, None, None, None, None


class _Unmerge(torch.autograd.Function):
    @staticmethod
    def forward(data, meta):
        ent = _unmerge_plans(data, meta)
        out = torch.empty(data.numel(), dtype=data.dtype, device=data.device)
        _run_copy(ent["fwd"], data, out, zero=False)
        return out

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.meta = inputs[1]

    @staticmethod
    def backward(ctx, grad):
        ent = _unmerge_plans(grad, ctx.meta)
        plan = _bwd_copy_plan(ent, grad.dtype, grad.device.index)
        out = torch.empty(grad.numel(), dtype=grad.dtype, device=grad.device)
        _run_copy(plan, grad, out, zero=False)
        return out, None


class _Transpose(torch.autograd.Function):
    @staticmethod
    def forward(data, axes, meta_transpose):
        ent = _transpose_plans(data, axes, meta_transpose)
        out = torch.empty(data.numel(), dtype=data.dtype, device=data.device)
        _run_copy(ent["fwd"], data, out, zero=False)
        return out

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.args = inputs[1:]

    @staticmethod
    def backward(ctx, grad):
        axes, meta = ctx.args
        ent = _transpose_plans(grad, axes, meta)
        plan = _bwd_copy_plan(ent, grad.dtype, grad.device.index)
        out = torch.empty(grad.numel(), dtype=grad.dtype, device=grad.device)
        _run_copy(plan, grad, out, zero=False)
        return out, None, None


def _promote(Adata, Bdata):
    dtype = torch.promote_types(Adata.dtype, Bdata.dtype)
    if Adata.dtype != dtype:
        Adata = Adata.to(dtype)
    if Bdata.dtype != dtype:
        Bdata = Bdata.to(dtype)
    return Adata, Bdata, dtype


class _Dot(torch.autograd.Function):
    @staticmethod
    def forward(Adata, Bdata, meta_dot, Dsize):
        Adata, Bdata, dtype = _promote(Adata, Bdata)
        ent = _dot_plans(meta_dot, dtype, Adata.device.index)
        # every element of C is written by exactly one GEMM: no zero fill needed
        out = torch.empty(Dsize, dtype=dtype, device=Adata.device)
        _run_gemm(ent["fwd"], Adata, Bdata, out)
        return out

    @staticmethod
    def setup_context(ctx, inputs, output):
        Adata, Bdata, meta_dot, Dsize = inputs
        ctx.save_for_backward(Adata, Bdata)
        ctx.meta_dot = meta_dot

    @staticmethod
    def backward(ctx, grad):
        Adata, Bdata = ctx.saved_tensors
        in_dtypes = (Adata.dtype, Bdata.dtype)
        Adata, Bdata, dtype = _promote(Adata, Bdata)
        if grad.dtype != dtype:
            grad = grad.to(dtype)
        dev = Adata.device.index
        ent = _dot_plans(ctx.meta_dot, dtype, dev)
        if ent["bwd"] is None:
            pa, sa, pb, sb = plans.dot_backward_tables(ctx.meta_dot)
            ent["bwd"] = (plans.GemmPlan(pa, sa, _DTYPE_CODE[dtype], dev), plans.GemmPlan(pb, sb, _DTYPE_CODE[dtype], dev))
        plan_a, plan_b = ent["bwd"]
        # blocks of A / B that take part in no product keep a zero gradient
        gA = torch.zeros(Adata.numel(), dtype=dtype, device=Adata.device)
        gB = torch.zeros(Bdata.numel(), dtype=dtype, device=Bdata.device)
        _run_gemm(plan_a, grad, Bdata, gA, conj_b=True)     # A_b = C_b @ B^H
        _run_gemm(plan_b, Adata, grad, gB, conj_a=True)     # B_b = A^H @ C_b
        if in_dtypes[0] != dtype:
            gA = gA.real.to(in_dtypes[0]) if not in_dtypes[0].is_complex else gA.to(in_dtypes[0])
        if in_dtypes[1] != dtype:
            gB = gB.real.to(in_dtypes[1]) if not in_dtypes[1].is_complex else gB.to(in_dtypes[1])
        return gA, gB, None, None


_SCATTER_TABLE_LIMIT = 1 << 19    # rows + columns of all merged blocks whose lookup tables the fused epilogue may hold


def _dot_unmerge_forward(Adata, Bdata, meta_dot, Dsize, meta_unmerge, out=None, dst_shift=None):
    Adata, Bdata, dtype = _promote(Adata, Bdata)
    dev = Adata.device.index
    key = ("dotunm", id(meta_dot), id(meta_unmerge), id(dst_shift), dtype, dev)

    def build():
        # the scatter epilogue looks every row and column of a merged block up in a table: for tall-and-skinny products (an
        # environment update has blocks of 10^6..10^7 rows and a handful of columns) those tables would be larger than the
        # operands (measured: 48 ms of plan creation per call in a D=4096 DMRG sweep) -> two launches instead
        if dst_shift is None and sum(rec[1][0] + rec[1][1] for rec in meta_dot) > _SCATTER_TABLE_LIMIT:
            return {"fwd": None, "ref": (meta_unmerge, dst_shift)}
        problems, segments = plans.dot_tables(meta_dot)
        scatter = plans.unmerge_scatter_tables(meta_dot, meta_unmerge, dst_shift)
        return {"fwd": plans.GemmPlan(problems, segments, _DTYPE_CODE[dtype], dev, scatter), "ref": (meta_unmerge, dst_shift)}
    ent = _CACHE.get(key, meta_dot, build)
    if ent["fwd"] is None:
        if out is None:
            return _Unmerge.forward(_Dot.forward(Adata, Bdata, meta_dot, Dsize), meta_unmerge)
        # caller's buffer (a rank's share of a sharded contraction): only the records' destinations may be touched
        tmp = _Dot.forward(Adata, Bdata, meta_dot, Dsize)
        _run_copy(_unmerge_plans(tmp, meta_unmerge)["fwd"], tmp, out, zero=False)
        return out
    if out is None:
        out = torch.empty(Dsize, dtype=dtype, device=Adata.device)
    elif out.dtype != dtype or out.numel() != Dsize or out.device != Adata.device or not out.is_contiguous():
        raise ValueError("yastn_b200.dot_unmerge: out must be a contiguous 1-D tensor of the result's dtype, size and device")
    _run_gemm(ent["fwd"], Adata, Bdata, out)
    return out


class _DotUnmerge(torch.autograd.Function):
    """dot followed by unmerge in ONE launch: the GEMM epilogue scatters straight into the unmerged block layout
    (reference: the two calls at yastn/tensor/_contractions.py:152-155).  Backward = adjoint unmerge + dot backward."""

    @staticmethod
    def forward(Adata, Bdata, meta_dot, Dsize, meta_unmerge):
        return _dot_unmerge_forward(Adata, Bdata, meta_dot, Dsize, meta_unmerge)

    @staticmethod
    def setup_context(ctx, inputs, output):
        Adata, Bdata, meta_dot, Dsize, meta_unmerge = inputs
        ctx.save_for_backward(Adata, Bdata)
        ctx.metas = (meta_dot, Dsize, meta_unmerge)

    @staticmethod
    def backward(ctx, grad):
        meta_dot, Dsize, meta_unmerge = ctx.metas
        Adata, Bdata = ctx.saved_tensors
        with torch.enable_grad():
            A = Adata.detach().requires_grad_(True)
            B = Bdata.detach().requires_grad_(True)
            out = _Unmerge.apply(_Dot.apply(A, B, meta_dot, Dsize), meta_unmerge)
        gA, gB = torch.autograd.grad(out, (A, B), grad)
        return gA, gB, None, None, None


def _tds_plans(meta_dot, Areshape, Breshape, Aorder, Border, dtype, device):
    key = ("tds", id(meta_dot), id(Areshape), id(Breshape), tuple(Aorder), tuple(Border), dtype, device)

    def build():
        problems, segments, pack_a, pack_b = plans.tds_tables(meta_dot, Areshape, Breshape, Aorder, Border)
        ent = {"gemm": plans.GemmPlan(problems, segments, _DTYPE_CODE[dtype], device), "pack_a": None, "pack_b": None,
               "refs": (Areshape, Breshape)}   # keep the ids used in the key alive
        if pack_a:
            recs, rank = plans.pack_records(Areshape, Aorder)
            ent["pack_a"] = plans.CopyPlan(recs, rank, _ITEMSIZE[dtype], device)
        if pack_b:
            recs, rank = plans.pack_records(Breshape, Border)
            ent["pack_b"] = plans.CopyPlan(recs, rank, _ITEMSIZE[dtype], device)
        return ent
    return _CACHE.get(key, meta_dot, build)


def _tds_forward(Adata, Bdata, meta_dot, Areshape, Breshape, Aorder, Border, Dsize):
    Adata, Bdata, dtype = _promote(Adata, Bdata)
    ent = _tds_plans(meta_dot, Areshape, Breshape, Aorder, Border, dtype, Adata.device.index)
    if ent["pack_a"] is not None:
        packed = torch.empty(Adata.numel(), dtype=dtype, device=Adata.device)
        _run_copy(ent["pack_a"], Adata, packed, zero=False)   # blocks not taking part stay uninitialised and unread
        Adata = packed
    if ent["pack_b"] is not None:
        packed = torch.empty(Bdata.numel(), dtype=dtype, device=Bdata.device)
        _run_copy(ent["pack_b"], Bdata, packed, zero=False)
        Bdata = packed
    out = torch.empty(Dsize, dtype=dtype, device=Adata.device)
    _run_gemm(ent["gemm"], Adata, Bdata, out)
    return out


class _TransposeDotSum(torch.autograd.Function):
    @staticmethod
    def forward(Adata, Bdata, meta_dot, Areshape, Breshape, Aorder, Border, Dsize):
        return _tds_forward(Adata, Bdata, meta_dot, Areshape, Breshape, Aorder, Border, Dsize)

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.save_for_backward(inputs[0], inputs[1])
        ctx.args = inputs[2:]

    @staticmethod
    def backward(ctx, grad):
        # adjoint through explicit packing: At_b = C_b @ Bt^H, Bt_b = At^H @ C_b, then un-permute
        Adata, Bdata = ctx.saved_tensors
        meta_dot, Areshape, Breshape, Aorder, Border, Dsize = ctx.args
        in_dtypes = (Adata.dtype, Bdata.dtype)
        Adata, Bdata, dtype = _promote(Adata, Bdata)
        grad = grad.to(dtype) if grad.dtype != dtype else grad
        dev = Adata.device
        key = ("tdsb", id(meta_dot), id(Areshape), id(Breshape), tuple(Aorder), tuple(Border), dtype, dev.index)

        def build():
            ra, rka = plans.pack_records(Areshape, Aorder)
            rb, rkb = plans.pack_records(Breshape, Border)
            # packed-layout forward meta: C[sl] = sum_pairs At[ia] (Dl x K) @ Bt[ib] (K x Dr)
            recs = []
            for (sl, (Dl, Dr), pairs) in meta_dot:
                for ia, ib in pairs:
                    sla, _, _, K = Areshape[ia]
                    slb = Breshape[ib][0]
                    recs.append((sl, (Dl, Dr), sla, (Dl, K), slb, (K, Dr)))
            pa, sa, pb, sb = plans.dot_backward_tables(recs)
            isz, code = _ITEMSIZE[dtype], _DTYPE_CODE[dtype]
            return {"pack_a": plans.CopyPlan(ra, rka, isz, dev.index), "pack_b": plans.CopyPlan(rb, rkb, isz, dev.index),
                    "unpack_a": plans.CopyPlan(plans.reverse_records(ra, rka), rka, isz, dev.index),
                    "unpack_b": plans.CopyPlan(plans.reverse_records(rb, rkb), rkb, isz, dev.index),
                    "ga": plans.GemmPlan(pa, sa, code, dev.index), "gb": plans.GemmPlan(pb, sb, code, dev.index)}
        ent = _CACHE.get(key, meta_dot, build)
        At = torch.zeros(Adata.numel(), dtype=dtype, device=dev)
        Bt = torch.zeros(Bdata.numel(), dtype=dtype, device=dev)
        _run_copy(ent["pack_a"], Adata, At, zero=False)
        _run_copy(ent["pack_b"], Bdata, Bt, zero=False)
        gAt = torch.zeros(Adata.numel(), dtype=dtype, device=dev)
        gBt = torch.zeros(Bdata.numel(), dtype=dtype, device=dev)
        _run_gemm(ent["ga"], grad, Bt, gAt, conj_b=True)
        _run_gemm(ent["gb"], At, grad, gBt, conj_a=True)
        gA = torch.zeros(Adata.numel(), dtype=dtype, device=dev)
        gB = torch.zeros(Bdata.numel(), dtype=dtype, device=dev)
        _run_copy(ent["unpack_a"], gAt, gA, zero=False)
        _run_copy(ent["unpack_b"], gBt, gB, zero=False)
        if in_dtypes[0] != dtype:
            gA = gA.real.to(in_dtypes[0]) if not in_dtypes[0].is_complex else gA.to(in_dtypes[0])
        if in_dtypes[1] != dtype:
            gB = gB.real.to(in_dtypes[1]) if not in_dtypes[1].is_complex else gB.to(in_dtypes[1])
        return gA, gB, None, None, None, None, None, None


# -------------------------------------------------------------------------------------------------
# public backend functions (names and signatures of yastn.backend.backend_torch)
# -------------------------------------------------------------------------------------------------

def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t.requires_grad for t in tensors)


# autograd.Function.apply costs ~10 us of host time per call; without gradients the forward is called directly

def transpose_and_merge(data, order, meta_new, meta_mrg, Dsize):
    _check(data, "transpose_and_merge")
    if _needs_grad(data):
        return _TransposeAndMerge.apply(data, order, meta_new, meta_mrg, Dsize)
    return _TransposeAndMerge.forward(data, order, meta_new, meta_mrg, Dsize)


def unmerge(data, meta):
    _check(data, "unmerge")
    if _needs_grad(data):
        return _Unmerge.apply(data, meta)
    return _Unmerge.forward(data, meta)


def transpose(data, axes, meta_transpose):
    _check(data, "transpose")
    if _needs_grad(data):
        return _Transpose.apply(data, axes, meta_transpose)
    return _Transpose.forward(data, axes, meta_transpose)


def dot(Adata, Bdata, meta_dot, Dsize):
    _check(Adata, "dot")
    _check(Bdata, "dot")
    if _needs_grad(Adata, Bdata):
        return _Dot.apply(Adata, Bdata, meta_dot, Dsize)
    return _Dot.forward(Adata, Bdata, meta_dot, Dsize)


def dot_unmerge(Adata, Bdata, meta_dot, Dsize, meta_unmerge, out=None, dst_shift=None):
    """``unmerge(dot(Adata, Bdata, meta_dot, Dsize), meta_unmerge)`` in one launch (fused scatter epilogue).

    Multi-GPU (forward only): ``out`` is the result buffer inside a ``peer.PeerArena`` and ``dst_shift`` an int64 array with
    one entry per record of ``meta_unmerge`` — the element offset (``PeerArena.shift``) of the rank that multiplies that
    output block next, 0 for this rank.  The epilogue then stores every block straight into its next owner's HBM over
    NVLink while the remaining tiles are still being multiplied: GEMM, unmerge and redistribution in one launch."""
    _check(Adata, "dot_unmerge")
    _check(Bdata, "dot_unmerge")
    if _needs_grad(Adata, Bdata):
        if out is not None or dst_shift is not None:
            raise ValueError("yastn_b200.dot_unmerge: out= / dst_shift= are forward-only")
        return _DotUnmerge.apply(Adata, Bdata, meta_dot, Dsize, meta_unmerge)
    return _dot_unmerge_forward(Adata, Bdata, meta_dot, Dsize, meta_unmerge, out, dst_shift)


def dot_into(Adata, Bdata, meta_dot, out):
    """``dot`` into the caller's buffer: only the C blocks of ``meta_dot`` are written (a rank's row panels of a sharded
    contraction, yastn_b200.spmd).  Forward only."""
    _check(Adata, "dot_into")
    _check(Bdata, "dot_into")
    Adata, Bdata, dtype = _promote(Adata, Bdata)
    if out.dtype != dtype or out.device != Adata.device or not out.is_contiguous():
        raise ValueError("yastn_b200.dot_into: out must be a contiguous 1-D tensor of the result's dtype and device")
    _run_gemm(_dot_plans(meta_dot, dtype, Adata.device.index)["fwd"], Adata, Bdata, out)
    return out


def transpose_and_merge_partial(data, order, meta_new, meta_mrg, Dsize):
    """``transpose_and_merge`` of a subset of the merged blocks (a rank's share of a sharded contraction): the result has the
    full size ``Dsize`` but only the blocks of ``meta_new`` are defined — cells of those blocks that no source block covers are
    zero-filled by the plan's own records, everything else stays uninitialised and must not be read.  Forward only."""
    _check(data, "transpose_and_merge_partial")
    ent = _merge_plans(data, order, meta_new, meta_mrg, Dsize)
    out = torch.empty(Dsize, dtype=data.dtype, device=data.device)
    kept = sum(hi - lo for (_, _, (lo, hi)) in meta_new)
    _run_copy(ent["fwd"], data, out, zero=ent["fwd"].covered < kept)     # the zero-fill records were not built: clear everything
    return out


def transpose_dot_sum(Adata, Bdata, meta_dot, Areshape, Breshape, Aorder, Border, Dsize):
    _check(Adata, "transpose_dot_sum")
    _check(Bdata, "transpose_dot_sum")
    if _needs_grad(Adata, Bdata):
        return _TransposeDotSum.apply(Adata, Bdata, meta_dot, Areshape, Breshape, Aorder, Border, Dsize)
    return _tds_forward(Adata, Bdata, meta_dot, Areshape, Breshape, Aorder, Border, Dsize)


def vdot(Adata, Bdata, meta):
    """``sum_ii dot(Adata[sla_ii], Bdata[slb_ii])`` over the common blocks of two tensors (yastn/backend/backend_torch.py:537-546,
    called by yastn.vdot, yastn/tensor/_contractions.py:590-630, in every Lanczos step): the reference runs one ``torch.dot``
    per block; here all blocks are 1 x 1 problems of ONE grouped-GEMM launch (the long contraction index is shared out over the
    CTAs by stream-K) followed by one sum.  Conjugation arrives as torch's lazy conj bit, as for ``dot``.  Forward only."""
    _check(Adata, "vdot")
    _check(Bdata, "vdot")
    Adata, Bdata, dtype = _promote(Adata, Bdata)
    n = len(meta)
    if n == 0:
        return torch.zeros((), dtype=dtype, device=Adata.device)
    dev = Adata.device.index
    key = ("vdot", id(meta), dtype, dev)

    def build():
        problems, segments = plans.vdot_tables(meta)
        return {"fwd": plans.GemmPlan(problems, segments, _DTYPE_CODE[dtype], dev)}
    ent = _CACHE.get(key, meta, build)
    tmp = torch.empty(n, dtype=dtype, device=Adata.device)
    _run_gemm(ent["fwd"], Adata, Bdata, tmp)
    return torch.sum(tmp)


# -------------------------------------------------------------------------------------------------
# block-wise elementwise operations: one launch each instead of a Python loop over blocks
# -------------------------------------------------------------------------------------------------

def _run_ew(plan, dst, srcs, aux=None):
    if _recorder is not None:
        _recorder.unsupported("elementwise plan")
    _on_device(dst.device, lambda st: plan.run(dst.data_ptr(), [s.data_ptr() for s in srcs], None if aux is None else aux.data_ptr(), st))


def _plain(t, dtype):
    """Contiguous tensor of ``dtype`` with the conj bit resolved (elementwise kernels read raw storage)."""
    if t.dtype != dtype:
        t = t.to(dtype)
    if t.is_conj():
        t = t.resolve_conj()
    return t if t.is_contiguous() else t.contiguous()


def _lincomb(datas, metas, Dsize, signs, name):
    for d in datas:
        _check(d, name)
    dtype = datas[0].dtype
    for d in datas[1:]:
        dtype = torch.promote_types(dtype, d.dtype)
    dev = datas[0].device
    key = (name, id(metas), signs, dtype, dev.index)

    def build():
        rounds = plans.add_tables(metas, signs)
        covered = sum(hi - lo for (lo, hi) in {sl_c for meta in metas for (sl_c, _) in meta})
        return {"rounds": [(plans.EwPlan(recs, _ITEMSIZE[dtype], dev.index), slots) for recs, slots in rounds], "covered": covered}
    ent = _CACHE.get(key, metas, build)
    datas = [_plain(d, dtype) for d in datas]
    # the reference starts from zeros; every output block is written by at least one operand, so only gaps would need them
    out = torch.empty(Dsize, dtype=dtype, device=dev) if ent["covered"] >= Dsize else torch.zeros(Dsize, dtype=dtype, device=dev)
    for plan, slots in ent["rounds"]:
        _run_ew(plan, out, [out if k < 0 else datas[k] for k in slots])
    return out


def add(datas, metas, Dsize):
    """``sum_k datas[k]`` block by block (yastn/backend/backend_torch.py:518-524): one launch for up to four operands."""
    return _lincomb(list(datas), metas, Dsize, (1,) * len(datas), "add")


def sub(Adata, Bdata, metas, Dsize):
    """``Adata - Bdata`` block by block (yastn/backend/backend_torch.py:527-534)."""
    return _lincomb([Adata, Bdata], metas, Dsize, (1, -1), "sub")


def _negate_forward(Adata, slices):
    key = ("neg", id(slices), Adata.numel(), Adata.dtype, Adata.device.index)

    def build():
        return {"fwd": plans.EwPlan(plans.negate_tables(slices, Adata.numel()), _ITEMSIZE[Adata.dtype], Adata.device.index)}
    ent = _CACHE.get(key, slices, build)
    A = _plain(Adata, Adata.dtype)
    out = torch.empty_like(A)
    _run_ew(ent["fwd"], out, [A])
    return out


class _NegateBlocks(torch.autograd.Function):
    @staticmethod
    def forward(Adata, slices):
        return _negate_forward(Adata, slices)

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.slices = inputs[1]

    @staticmethod
    def backward(ctx, grad):
        return _negate_forward(grad, ctx.slices), None


def negate_blocks(Adata, slices):
    """Copy of ``Adata`` with the listed blocks multiplied by -1 (yastn/backend/_backend_torch_backwards.py:229-248; every
    fermionic ``swap_gate``): one launch instead of a clone plus one in-place multiply per block."""
    _check(Adata, "negate_blocks")
    if _needs_grad(Adata):
        return _NegateBlocks.apply(Adata, slices)
    return _negate_forward(Adata, slices)


def dot_diag(Adata, Bdata, meta, Dsize, axis, a_ndim):
    """Diagonal tensor times one leg of a tensor (yastn/backend/backend_torch.py:557-564, ``yastn.broadcast``): one launch.
    Forward only (the YASTN module hands inputs that require grad to the reference's differentiable loop)."""
    _check(Adata, "dot_diag")
    _check(Bdata, "dot_diag")
    dtype = torch.promote_types(Adata.dtype, Bdata.dtype)
    dev = Bdata.device
    key = ("ddiag", id(meta), axis, a_ndim, dtype, dev.index)

    def build():
        return {"fwd": plans.EwPlan(plans.dot_diag_tables(meta, axis, a_ndim), _ITEMSIZE[dtype], dev.index)}
    ent = _CACHE.get(key, meta, build)
    out = torch.empty(Dsize, dtype=dtype, device=dev)
    _run_ew(ent["fwd"], out, [_plain(Bdata, dtype)], _plain(Adata, dtype))
    return out


def _mask_index(mask, order, device):
    idx = [mask[tm] for tm in order]
    idx = [torch.as_tensor(v, device=device).to(torch.int64).reshape(-1) for v in idx]
    return idx[0].contiguous() if len(idx) == 1 else torch.cat(idx)


def _whole_blocks(mask):
    """True when every mask entry is ``slice(None)``: yastn.flip_charges moves whole blocks through embed_mask with
    ``mask = {0: slice(None)}`` (yastn/tensor/_single.py:224)."""
    vals = list(mask.values())
    return bool(vals) and all(isinstance(v, slice) and v == slice(None) for v in vals)


def _mask_run(data, mask, meta, Dsize, axis, ndim, scatter):
    """GATHER (scatter=False) / SCATTER (True) of the records of ``meta``; SCATTER leaves the unselected positions zero."""
    whole = _whole_blocks(mask)
    key = ("mask", id(meta), axis, ndim, scatter, whole, data.dtype, data.device.index)

    def build():
        if whole:
            return {"fwd": plans.EwPlan(plans.block_copy_tables(meta), _ITEMSIZE[data.dtype], data.device.index), "order": None}
        recs, order = plans.mask_tables(meta, axis, ndim, scatter)
        return {"fwd": plans.EwPlan(recs, _ITEMSIZE[data.dtype], data.device.index), "order": order}
    ent = _CACHE.get(key, meta, build)
    out = (torch.zeros if scatter else torch.empty)(Dsize, dtype=data.dtype, device=data.device)
    if len(meta):
        _run_ew(ent["fwd"], out, [_plain(data, data.dtype)], None if whole else _mask_index(mask, ent["order"], data.device))
    return out


def _swap_meta(meta):
    """Records of the adjoint: source and destination blocks trade places."""
    return tuple((sla, Da, sln, Dn, tm) for sln, Dn, sla, Da, tm in meta)


class _ApplyMask(torch.autograd.Function):
    @staticmethod
    def forward(Adata, mask, meta, Dsize, axis, ndim):
        return _mask_run(Adata, mask, meta, Dsize, axis, ndim, scatter=False)

    @staticmethod
    def setup_context(ctx, inputs, output):
        Adata, ctx.mask, ctx.meta, _, ctx.axis, ctx.ndim = inputs
        ctx.n = Adata.numel()

    @staticmethod
    def backward(ctx, grad):
        key = ("maskadj", id(ctx.meta))
        adj = _CACHE.get(key, ctx.meta, lambda: _swap_meta(ctx.meta))
        return _mask_run(grad, ctx.mask, adj, ctx.n, ctx.axis, ctx.ndim, scatter=True), None, None, None, None, None


class _EmbedMask(torch.autograd.Function):
    @staticmethod
    def forward(Adata, mask, meta, Dsize, axis, ndim):
        return _mask_run(Adata, mask, meta, Dsize, axis, ndim, scatter=True)

    @staticmethod
    def setup_context(ctx, inputs, output):
        Adata, ctx.mask, ctx.meta, _, ctx.axis, ctx.ndim = inputs
        ctx.n = Adata.numel()

    @staticmethod
    def backward(ctx, grad):
        key = ("maskadj", id(ctx.meta))
        adj = _CACHE.get(key, ctx.meta, lambda: _swap_meta(ctx.meta))
        out = _mask_run(grad, ctx.mask, adj, ctx.n, ctx.axis, ctx.ndim, scatter=False)
        return out, None, None, None, None, None


def apply_mask(Adata, mask, meta, Dsize, axis, ndim):
    """Keep the positions ``mask[tm]`` along one leg of every block (yastn/backend/_backend_torch_backwards.py:251-279; the
    truncation after every SVD): one gather launch (+ one ``cat`` of the index vectors) instead of an index-select per block."""
    _check(Adata, "apply_mask")
    if _needs_grad(Adata):
        return _ApplyMask.apply(Adata, mask, meta, Dsize, axis, ndim)
    return _mask_run(Adata, mask, meta, Dsize, axis, ndim, scatter=False)


def embed_mask(Adata, mask, meta, Dsize, axis, ndim):
    """Inverse placement: blocks go to the positions ``mask[tm]`` of a zero tensor (:282-310)."""
    _check(Adata, "embed_mask")
    if _needs_grad(Adata):
        return _EmbedMask.apply(Adata, mask, meta, Dsize, axis, ndim)
    return _mask_run(Adata, mask, meta, Dsize, axis, ndim, scatter=True)


def trace(data, order, meta, Dsize):
    """Partial trace over pairs of legs (yastn/backend/backend_torch.py:268-275): one launch.  Forward only."""
    _check(data, "trace")
    key = ("trace", id(meta), tuple(order), data.dtype, data.device.index)

    def build():
        recs, traces = plans.trace_tables(order, meta)
        covered = sum(sln[1] - sln[0] for sln, _ in meta)
        return {"fwd": plans.EwPlan(recs, _ITEMSIZE[data.dtype], data.device.index, traces), "covered": covered}
    ent = _CACHE.get(key, meta, build)
    out = (torch.empty if ent["covered"] >= Dsize else torch.zeros)(Dsize, dtype=data.dtype, device=data.device)
    _run_ew(ent["fwd"], out, [_plain(data, data.dtype)])
    return out


_BS_CACHE = plans.PlanCache(maxsize=1024)


def kernel_tensordot_bs(a, b, NSYM, a_struct_t, a_slices, a_t_per_mode, a_D_per_mode, nout_a, nin_a,
                        b_struct_t, b_slices, b_t_per_mode, b_D_per_mode, nout_b, nin_b,
                        c_size, c_struct_t, c_slices, profile=False):
    """Whole block-sparse contraction from raw block tables in one call: same signature and meaning as the reference's
    single-call boundary (yastn/backend/backend_torch_cpp.py:173-228, reached from yastn/tensor/_contractions.py:199-242
    when ``BACKEND_ID == 'torch_cpp'`` under the ``no_fusion`` policy; there it rebuilds a cuTENSOR plan on every call).
    Here the block-pair join is a cached, vectorised meta pass (plans.bs_to_tds) and the contraction is ONE grouped-GEMM
    launch (plus one packing copy per operand whose permuted blocks are not strided matrices).  The per-mode charge /
    dimension lists are redundant with the ``D`` of the slice records and are not needed."""
    _check(a, "kernel_tensordot_bs")
    _check(b, "kernel_tensordot_bs")
    if c_size == 0:
        return torch.zeros(0, dtype=torch.promote_types(a.dtype, b.dtype), device=a.device)
    a_struct_t, b_struct_t, c_struct_t = tuple(a_struct_t), tuple(b_struct_t), tuple(c_struct_t)
    a_slices, b_slices, c_slices = tuple(a_slices), tuple(b_slices), tuple(c_slices)
    nout_a, nin_a, nout_b, nin_b = tuple(nout_a), tuple(nin_a), tuple(nout_b), tuple(nin_b)
    key = (a_struct_t, a_slices, nout_a, nin_a, b_struct_t, b_slices, nout_b, nin_b, c_struct_t, c_slices)
    ent = _BS_CACHE._d.get(key)
    if ent is None:
        ent = (None, plans.bs_to_tds(a_struct_t, a_slices, nout_a, nin_a, b_struct_t, b_slices, nout_b, nin_b, c_struct_t, c_slices))
        _BS_CACHE._d[key] = ent
        _BS_CACHE.misses += 1
        if len(_BS_CACHE._d) > _BS_CACHE.maxsize:
            _BS_CACHE._d.popitem(last=False)
    else:
        _BS_CACHE.hits += 1
    meta_dot, Areshape, Breshape = ent[1]
    return transpose_dot_sum(a, b, meta_dot, Areshape, Breshape, nout_a + nin_a, nin_b + nout_b, c_size)


HOT_FUNCTIONS = ("transpose_and_merge", "unmerge", "transpose", "dot", "transpose_dot_sum")
