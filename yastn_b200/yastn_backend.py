"""Plug the B200 kernels into YASTN through YASTN's own backend-module interface.

YASTN selects its numerics with ``yastn.make_config(backend=<module>)``: a backend is a Python module exposing
~90 free functions plus ``BACKEND_ID`` and ``DTYPE`` (yastn/backend/backend_torch.py:32-58), and module objects
pass through ``make_config`` untouched (yastn/tensor/_initialize.py:106-115).  Two equivalent install modes:

  (A) ``cfg = yastn.make_config(backend=yastn_b200.yastn_backend.module(), default_device='cuda', ...)``
      — a module object that re-exports the stock torch backend and overrides the five hot functions (plus ``vdot`` and the
      sector-parallel ``svd`` / ``svdvals`` / ``eigh`` / ``qr`` of yastn_b200.decomp);
  (B) ``yastn_b200.yastn_backend.activate()`` — rebinds the same functions on
      ``yastn.backend.backend_torch`` in place, after which ``make_config(backend='torch')`` (and the reference's
      unmodified test-suite run with ``--backend torch --device cuda``) drives the B200 kernels.

``BACKEND_ID`` stays ``"torch"`` so tensors remain combinable / serialisable with the stock torch backend
(yastn/tensor/_tests.py:38-39, _output.py:72).

Scope of the replacement: CUDA tensors of dtype float64 / complex128.  There is no CPU path: a CPU tensor
reaching a hot function raises.  CUDA tensors of the other YASTN dtypes (float32 / complex64 / bool, which the
reference's tests touch in a few places) are handed to the reference's own torch kernels on the same device when
``delegate_other_dtypes=True`` (default) — that is the reference's GPU code, not a fallback of ours — and raise
otherwise.
"""
import types

import torch

from . import backend_b200 as _bk
from . import decomp as _decomp

_HOT = _bk.HOT_FUNCTIONS
_EWISE = ("add", "sub", "negate_blocks", "dot_diag", "apply_mask", "embed_mask", "trace")
_DECOMP = ("svd", "svdvals", "eigh", "qr")
_NATIVE = (torch.float64, torch.complex128)
_state = {"module": {}, "saved": None, "cpu_passthrough": False, "saved_f2m": None, "saved_decomp": None,
          "calls": {name: 0 for name in _HOT + ("dot_unmerge", "kernel_tensordot_bs", "vdot") + _EWISE},
          "delegated": {name: 0 for name in _HOT + _EWISE}}


def _stock():
    try:
        import yastn.backend.backend_torch as stock
    except ImportError as e:  # loud: the boundary needs the host library it plugs into
        raise ImportError("yastn_b200.yastn_backend needs the yastn package importable (pip install yastn, or put "
                          "the reference checkout on sys.path)") from e
    return stock


def _native(*tensors):
    """True: float64 / complex128 CUDA tensors (our kernels).  False: the call is handed to the reference's own torch code
    (other dtypes; CPU tensors only when the stock module was patched with ``activate(cpu_passthrough=True)``)."""
    for t in tensors:
        if not t.is_cuda:
            if _state["cpu_passthrough"]:
                return False
            raise TypeError(f"yastn_b200: CPU tensor reached a hot backend function (device {t.device}); "
                            "this backend has no CPU path — build the config with default_device='cuda'")
    return all(t.dtype in _NATIVE for t in tensors)


def _on_gpu(*tensors):
    """float64 / complex128 CUDA tensors only (the vector operations never raise for other inputs, see ``ewise`` below)."""
    return all(t.is_cuda and t.dtype in _NATIVE for t in tensors)


def _make_hot(stock_fns, delegate):
    """The five overriding functions; ``stock_fns`` are the reference's own implementations (for delegated dtypes)."""
    calls, delegated = _state["calls"], _state["delegated"]

    def other(name, *args):
        if not delegate and all(t.is_cuda for t in args if isinstance(t, torch.Tensor)):
            raise TypeError(f"yastn_b200.{name}: dtype not float64/complex128 and delegate_other_dtypes=False")
        delegated[name] += 1
        return stock_fns[name](*args)

    def transpose_and_merge(data, order, meta_new, meta_mrg, Dsize):
        if not _native(data):
            return other("transpose_and_merge", data, order, meta_new, meta_mrg, Dsize)
        calls["transpose_and_merge"] += 1
        return _bk.transpose_and_merge(data, order, meta_new, meta_mrg, Dsize)

    def unmerge(data, meta):
        if not _native(data):
            return other("unmerge", data, meta)
        calls["unmerge"] += 1
        return _bk.unmerge(data, meta)

    def transpose(data, axes, meta_transpose):
        if not _native(data):
            return other("transpose", data, axes, meta_transpose)
        calls["transpose"] += 1
        return _bk.transpose(data, axes, meta_transpose)

    def dot(Adata, Bdata, meta_dot, Dsize):
        if not _native(Adata, Bdata):
            return other("dot", Adata, Bdata, meta_dot, Dsize)
        calls["dot"] += 1
        return _bk.dot(Adata, Bdata, meta_dot, Dsize)

    def transpose_dot_sum(Adata, Bdata, meta_dot, Areshape, Breshape, Aorder, Border, Dsize):
        if not _native(Adata, Bdata):
            return other("transpose_dot_sum", Adata, Bdata, meta_dot, Areshape, Breshape, Aorder, Border, Dsize)
        calls["transpose_dot_sum"] += 1
        return _bk.transpose_dot_sum(Adata, Bdata, meta_dot, Areshape, Breshape, Aorder, Border, Dsize)

    def dot_unmerge(Adata, Bdata, meta_dot, Dsize, meta_unmerge):
        """dot + unmerge in one launch; called by the fused fuse_to_matrix tensordot (enable_fused_tensordot)."""
        if not _native(Adata, Bdata):
            return unmerge(dot(Adata, Bdata, meta_dot, Dsize), meta_unmerge)
        calls["dot_unmerge"] += 1
        return _bk.dot_unmerge(Adata, Bdata, meta_dot, Dsize, meta_unmerge)

    def kernel_tensordot_bs(a, b, *args, **kwargs):
        """Single-call boundary of the reference's torch_cpp backend (yastn/backend/backend_torch_cpp.py:173-188)."""
        if not _native(a, b):
            raise TypeError("yastn_b200.kernel_tensordot_bs: float64 / complex128 CUDA tensors only")
        calls["kernel_tensordot_bs"] += 1
        return _bk.kernel_tensordot_bs(a, b, *args, **kwargs)

    def vdot(Adata, Bdata, meta):
        """Block loop of torch.dot calls (backend_torch.py:537-546) as one grouped-GEMM launch; inputs that require grad, or are
        not float64 / complex128 CUDA tensors, keep the reference's differentiable loop."""
        grad = torch.is_grad_enabled() and (Adata.requires_grad or Bdata.requires_grad)
        if grad or len(meta) < 2 or not _native(Adata, Bdata):
            return stock_fns["vdot"](Adata, Bdata, meta)
        calls["vdot"] += 1
        return _bk.vdot(Adata, Bdata, meta)

    # ---- block-wise elementwise operations.  add / sub / dot_diag / trace are differentiated by torch itself in the
    # reference (plain tensor ops): inputs that require grad keep that code; negate_blocks / apply_mask / embed_mask carry
    # their own backward like the reference's autograd.Functions.
    def _grad(*tensors):
        return torch.is_grad_enabled() and any(t.requires_grad for t in tensors)

    def ewise(name, tensors, args, needs_forward_only):
        # these are the reference's general-purpose vector operations (also used on small host-side tensors while a model is
        # set up): anything that is not a float64 / complex128 CUDA tensor stays with the reference's own code
        if (needs_forward_only and _grad(*tensors)) or not _on_gpu(*tensors):
            delegated[name] += 1
            return stock_fns[name](*args)
        calls[name] += 1
        return getattr(_bk, name)(*args)

    def add(datas, metas, Dsize):
        return ewise("add", tuple(datas), (datas, metas, Dsize), True)

    def sub(Adata, Bdata, metas, Dsize):
        return ewise("sub", (Adata, Bdata), (Adata, Bdata, metas, Dsize), True)

    def negate_blocks(Adata, slices):
        return ewise("negate_blocks", (Adata,), (Adata, slices), False)

    def dot_diag(Adata, Bdata, meta, Dsize, axis, a_ndim):
        return ewise("dot_diag", (Adata, Bdata), (Adata, Bdata, meta, Dsize, axis, a_ndim), True)

    def _partial_slices(mask):
        """Masks given as slices other than slice(None) (no caller in the reference) keep the reference's code."""
        return any(isinstance(v, slice) and v != slice(None) for v in mask.values()) or \
            (any(isinstance(v, slice) for v in mask.values()) and not all(isinstance(v, slice) for v in mask.values()))

    def apply_mask(Adata, mask, meta, Dsize, axis, ndim):
        if _partial_slices(mask):
            delegated["apply_mask"] += 1
            return stock_fns["apply_mask"](Adata, mask, meta, Dsize, axis, ndim)
        return ewise("apply_mask", (Adata,), (Adata, mask, meta, Dsize, axis, ndim), False)

    def embed_mask(Adata, mask, meta, Dsize, axis, ndim):
        if _partial_slices(mask):
            delegated["embed_mask"] += 1
            return stock_fns["embed_mask"](Adata, mask, meta, Dsize, axis, ndim)
        return ewise("embed_mask", (Adata,), (Adata, mask, meta, Dsize, axis, ndim), False)

    def trace(data, order, meta, Dsize):
        return ewise("trace", (data,), (data, order, meta, Dsize), True)

    return {"transpose_and_merge": transpose_and_merge, "unmerge": unmerge, "transpose": transpose, "dot": dot,
            "transpose_dot_sum": transpose_dot_sum, "dot_unmerge": dot_unmerge, "kernel_tensordot_bs": kernel_tensordot_bs,
            "vdot": vdot, "add": add, "sub": sub, "negate_blocks": negate_blocks, "dot_diag": dot_diag, "apply_mask": apply_mask,
            "embed_mask": embed_mask, "trace": trace}


def _hook_cache_control():
    """``yastn.clear_cache()`` / ``yastn.set_cache_maxsize()`` (yastn/tensor/_control_lru.py:22-63) drop or re-wrap YASTN's
    lru-cached ``_meta_*`` functions; the device plans keyed on the identity of those metas can never be hit again, so the
    same calls also empty the plan cache (their tables go back to the device pool)."""
    if _state.get("lru_hooked"):
        return
    import sys
    import yastn
    ctl = sys.modules.get("yastn.tensor._control_lru")
    if ctl is None:
        return
    saved = {n: getattr(ctl, n) for n in ("clear_cache", "set_cache_maxsize")}

    def clear_cache():
        saved["clear_cache"]()
        _bk.clear_plan_cache()

    def set_cache_maxsize(maxsize=0):
        saved["set_cache_maxsize"](maxsize)
        _bk.clear_plan_cache()
    clear_cache.__doc__, set_cache_maxsize.__doc__ = saved["clear_cache"].__doc__, saved["set_cache_maxsize"].__doc__
    for mod in (ctl, sys.modules.get("yastn.tensor"), yastn):
        for name, fn in (("clear_cache", clear_cache), ("set_cache_maxsize", set_cache_maxsize)):
            if mod is not None and getattr(mod, name, None) is saved[name]:
                setattr(mod, name, fn)
    _state["lru_hooked"] = True


def _make_decomp(stock):
    """Sector-parallel svd / svdvals / eigh / qr (yastn_b200.decomp) on top of the reference's own implementations."""
    if _state["saved_decomp"] is None:
        _state["saved_decomp"] = types.SimpleNamespace(**{n: getattr(stock, n) for n in _DECOMP})
    return _decomp.make(_state["saved_decomp"])


def module(delegate_other_dtypes=True, bs_boundary=False):
    """Backend module object for ``yastn.make_config(backend=...)`` (install mode A).  One instance per variant and
    process: ``_config`` is an lru_cache key inside YASTN, so the module identity must be stable.

    ``bs_boundary=True`` returns the variant whose ``BACKEND_ID`` is ``"torch_cpp"``: YASTN then routes ``no_fusion``
    contractions through the single-call ``kernel_tensordot_bs`` boundary (yastn/tensor/_contractions.py:199-242)."""
    variant = (bool(bs_boundary), bool(delegate_other_dtypes))
    if variant not in _state["module"]:
        stock = _stock()
        _hook_cache_control()
        saved = _state["saved"] or {n: getattr(stock, n) for n in _HOT + ("vdot",) + _EWISE}
        mod = types.ModuleType("yastn_b200_backend" + ("_bs" if bs_boundary else ""),
                               "stock yastn torch backend with the B200 contraction kernels")
        for name in dir(stock):
            if not name.startswith("__"):
                setattr(mod, name, getattr(stock, name))
        for name, fn in _make_hot(saved, delegate_other_dtypes).items():
            setattr(mod, name, fn)
        for name, fn in _make_decomp(stock).items():
            setattr(mod, name, fn)
        mod.BACKEND_ID = "torch_cpp" if bs_boundary else "torch"
        mod.clear_plan_cache = _bk.clear_plan_cache
        mod.plan_cache_stats = _bk.plan_cache_stats
        _state["module"][variant] = mod
    return _state["module"][variant]


def activate(delegate_other_dtypes=True, cpu_passthrough=False):
    """Install mode B: rebind the five hot functions, ``vdot`` and the decompositions on ``yastn.backend.backend_torch`` itself
    (and add ``dot_unmerge`` / ``kernel_tensordot_bs``).

    ``activate`` patches the process-wide stock module, so configs built with ``backend='torch'`` on the CPU would reach our
    functions too.  By default that raises (this backend has no CPU path and never pretends to); with
    ``cpu_passthrough=True`` CPU tensors are handed back to the functions that were bound on the stock module before
    ``activate`` — the reference's own code, counted under ``call_counts()['delegated']``."""
    stock = _stock()
    _hook_cache_control()
    _state["cpu_passthrough"] = bool(cpu_passthrough)
    if _state["saved"] is None:
        _state["saved"] = {n: getattr(stock, n) for n in _HOT + ("vdot",) + _EWISE}
    for name, fn in _make_hot(_state["saved"], delegate_other_dtypes).items():
        if name in _state["saved"] or name in ("dot_unmerge", "kernel_tensordot_bs"):
            setattr(stock, name, fn)
    for name, fn in _make_decomp(stock).items():
        setattr(stock, name, fn)
    return stock


def deactivate():
    """Undo :func:`activate`."""
    _state["cpu_passthrough"] = False
    if _state["saved"] is not None:
        stock = _stock()
        for name, fn in _state["saved"].items():
            setattr(stock, name, fn)
        _state["saved"] = None
        for name in ("dot_unmerge", "kernel_tensordot_bs"):       # added by activate(), not part of the stock module
            if hasattr(stock, name):
                delattr(stock, name)
        if _state["saved_decomp"] is not None:
            for name in _DECOMP:
                setattr(stock, name, getattr(_state["saved_decomp"], name))


def enable_fused_tensordot():
    """Route YASTN's fuse_to_matrix tensordot through ``backend.dot_unmerge`` when the backend offers it: the grouped GEMM
    writes straight into the unmerged block layout, i.e. 3 launches per tensordot instead of 4 and one pass less over the
    result.  This is the optional 6-line change to ``_tensordot_f2m`` (yastn/tensor/_contractions.py:139-156) described in
    INTEGRATION.md, applied from outside; backends without ``dot_unmerge`` keep the reference's call sequence."""
    import yastn.tensor._contractions as C
    from yastn.tensor._merging import _no_change_in_unmerge
    if _state["saved_f2m"] is not None:
        return
    _state["saved_f2m"] = C._tensordot_f2m

    def tensordot_f2m(a, b, nout_a, nin_a, nin_b, nout_b, s_c):
        backend = a.config.backend
        fused = getattr(backend, "dot_unmerge", None)
        if fused is None:
            return _state["saved_f2m"](a, b, nout_a, nin_a, nin_b, nout_b, s_c)
        ind_a, ind_b = C._common_inds(a.struct.t, b.struct.t, nin_a, nin_b, a.ndim_n, b.ndim_n, a.config.sym.NSYM)
        data_a, struct_a, slices_a, ls_l, ls_ac = C._merge_to_matrix(a, (nout_a, nin_a), ind_a)
        data_b, struct_b, slices_b, ls_bc, ls_r = C._merge_to_matrix(b, (nin_b, nout_b), ind_b)
        if ls_ac != ls_bc:
            raise C.YastnError('Bond dimensions do not match.')
        meta_dot, struct_m, slices_m = C._meta_tensordot_f2m(struct_a, slices_a, struct_b, slices_b)
        meta_unmerge, struct_c, slices_c = C._meta_unmerge_matrix(a.config, struct_m, slices_m, ls_l, ls_r, s_c)
        if _no_change_in_unmerge(meta_unmerge):       # result already in block layout: plain dot (reference fast path)
            return backend.dot(data_a, data_b, meta_dot, struct_m.size), struct_c, slices_c
        return fused(data_a, data_b, meta_dot, struct_m.size, meta_unmerge), struct_c, slices_c
    C._tensordot_f2m = tensordot_f2m


def disable_fused_tensordot():
    if _state["saved_f2m"] is not None:
        import yastn.tensor._contractions as C
        C._tensordot_f2m = _state["saved_f2m"]
        _state["saved_f2m"] = None


def call_counts():
    """How many hot calls ran on the B200 kernels / were delegated to the reference's torch kernels."""
    return {"native": dict(_state["calls"]), "delegated": dict(_state["delegated"])}
