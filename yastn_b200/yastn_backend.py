"""Plug the B200 kernels into YASTN through YASTN's own backend-module interface.

YASTN selects its numerics with ``yastn.make_config(backend=<module>)``: a backend is a Python module exposing
~90 free functions plus ``BACKEND_ID`` and ``DTYPE`` (yastn/backend/backend_torch.py:32-58), and module objects
pass through ``make_config`` untouched (yastn/tensor/_initialize.py:106-115).  Two equivalent install modes:

  (A) ``cfg = yastn.make_config(backend=yastn_b200.yastn_backend.module(), default_device='cuda', ...)``
      — a module object that re-exports the stock torch backend and overrides the five hot functions;
  (B) ``yastn_b200.yastn_backend.activate()`` — rebinds the five hot functions on
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

_HOT = _bk.HOT_FUNCTIONS
_NATIVE = (torch.float64, torch.complex128)
_state = {"module": None, "saved": None, "calls": {name: 0 for name in _HOT}, "delegated": {name: 0 for name in _HOT}}


def _stock():
    try:
        import yastn.backend.backend_torch as stock
    except ImportError as e:  # loud: the boundary needs the host library it plugs into
        raise ImportError("yastn_b200.yastn_backend needs the yastn package importable (pip install yastn, or put "
                          "the reference checkout on sys.path)") from e
    return stock


def _native(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise TypeError(f"yastn_b200: CPU tensor reached a hot backend function (device {t.device}); "
                            "this backend has no CPU path — build the config with default_device='cuda'")
    return all(t.dtype in _NATIVE for t in tensors)


def _make_hot(stock_fns, delegate):
    """The five overriding functions; ``stock_fns`` are the reference's own implementations (for delegated dtypes)."""
    calls, delegated = _state["calls"], _state["delegated"]

    def other(name, *args):
        if not delegate:
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

    return {"transpose_and_merge": transpose_and_merge, "unmerge": unmerge, "transpose": transpose, "dot": dot,
            "transpose_dot_sum": transpose_dot_sum}


def module(delegate_other_dtypes=True):
    """Backend module object for ``yastn.make_config(backend=...)`` (install mode A).  One instance per process:
    ``_config`` is an lru_cache key inside YASTN, so the module identity must be stable."""
    if _state["module"] is None:
        stock = _stock()
        saved = _state["saved"] or {n: getattr(stock, n) for n in _HOT}
        mod = types.ModuleType("yastn_b200_backend", "stock yastn torch backend with the B200 contraction kernels")
        for name in dir(stock):
            if not name.startswith("__"):
                setattr(mod, name, getattr(stock, name))
        for name, fn in _make_hot(saved, delegate_other_dtypes).items():
            setattr(mod, name, fn)
        mod.BACKEND_ID = "torch"
        mod.clear_plan_cache = _bk.clear_plan_cache
        mod.plan_cache_stats = _bk.plan_cache_stats
        _state["module"] = mod
    return _state["module"]


def activate(delegate_other_dtypes=True):
    """Install mode B: rebind the five hot functions on ``yastn.backend.backend_torch`` itself."""
    stock = _stock()
    if _state["saved"] is None:
        _state["saved"] = {n: getattr(stock, n) for n in _HOT}
    for name, fn in _make_hot(_state["saved"], delegate_other_dtypes).items():
        setattr(stock, name, fn)
    return stock


def deactivate():
    """Undo :func:`activate`."""
    if _state["saved"] is not None:
        stock = _stock()
        for name, fn in _state["saved"].items():
            setattr(stock, name, fn)
        _state["saved"] = None


def call_counts():
    """How many hot calls ran on the B200 kernels / were delegated to the reference's torch kernels."""
    return {"native": dict(_state["calls"]), "delegated": dict(_state["delegated"])}
