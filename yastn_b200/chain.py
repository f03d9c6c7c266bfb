"""Tensordot chains: several contractions in a row recorded once and replayed by ONE library call.

The reference applies the effective Hamiltonian of DMRG as four ``yastn.tensordot`` calls (yastn/tn/mps/_env.py:512-518
``Heff2``; three in ``Heff1`` :506-510 and in ``update_env_to_last`` / ``update_env_to_first`` :496-504) and ``eigs`` calls it
again and again on operands whose block structure does not change between Lanczos iterations.  Every call walks YASTN's Python
meta functions (~0.2 ms per tensordot) before the first kernel is launched — at small bond dimension that, not the GPU, is the
run time.

``trace(name, fn, tensors)`` runs ``fn(*tensors)`` normally the first time it meets a tuple of operand *structures* and records
the launches the backend makes (copy and grouped-GEMM plans + which buffer feeds which).  When the data flow is closed — every
launch reads only the operands or the output of an earlier launch, and the result is the output of the last one — the chain
becomes a ``yb_chain`` (include/yastn_b200.h): later calls with operands of the same structure allocate the result and one
scratch arena (intermediates share it by lifetime) and replay all launches with one ``yb_chain_run`` — no YASTN metadata pass,
no per-step Python.  Anything else (a torch operation between the steps, an elementwise plan, inputs that require grad, an
empty result) is never replayed: ``fn`` simply runs as the reference would run it.

``enable()`` installs chains for ``Env_mps_mpo_mps.Heff1 / Heff2 / update_env_to_last / update_env_to_first``, ``enable_peps()``
for the double-layer contractions of CTMRG (``append_vec_*``, yastn/tn/fpeps/envs/_env_contractions.py:211-365); the chain can
additionally be replayed from a CUDA graph (``YASTN_B200_CHAIN_GRAPH=1``: operands are staged into static buffers, one graph
launch per application) — measured slower than the direct replay at every size, see DESIGN.md.
"""
import ctypes
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from . import backend_b200 as _bk

_ALIGN = 512
_MAX_CHAINS = 1024
_cache = OrderedDict()          # key -> _Chain | None (None: recorded, but not replayable)
_stats = {"recorded": 0, "replayed": 0, "rejected": 0, "bypassed": 0, "launches_replayed": 0}
_GRAPH = os.environ.get("YASTN_B200_CHAIN_GRAPH", "0") == "1"


class _Recorder:
    """Collects the launches of one traced call (installed as backend_b200._recorder while ``fn`` runs)."""

    def __init__(self):
        self.steps = []          # (kind, plan, (a, b, c) tensors, flags)
        self.bad = None

    def copy(self, plan, src, dst, flags):
        self.steps.append((_lib.YB_CHAIN_COPY, plan, (src, None, dst), flags))

    def gemm(self, plan, A, B, C, flags):
        self.steps.append((_lib.YB_CHAIN_GEMM, plan, (A, B, C), flags))

    def unsupported(self, what):
        self.bad = what


class _NativeChain:
    """Owns the yb_chain handle (the CPU test shim swaps this class for a table interpreter)."""

    def __init__(self, table, plans, nslots):
        self._lib = _lib.load()
        self.handle = ctypes.c_void_p()
        table = table.copy()
        table[:, 2] = [p.handle.value for p in plans]
        _lib.check(self._lib.yb_chain_create(table.ctypes.data, table.shape[0], nslots, ctypes.byref(self.handle)))
        self.nslots = nslots

    def run(self, ptrs, stream):
        rc = self._lib.yb_chain_run(self.handle, (ctypes.c_void_p * self.nslots)(*ptrs), self.nslots, stream)
        if rc:
            _lib.check(rc)

    def __del__(self):
        try:
            if self.handle:
                self._lib.yb_chain_destroy(self.handle)
        except Exception:
            pass


class _Chain:
    def __init__(self, runner, plans, n_in, nsteps, arena_bytes, out_elems, out_conj, dtype, template):
        self.runner, self.plans = runner, plans          # the chain borrows the plans: keep them alive
        self.n_in, self.nsteps, self.arena_bytes = n_in, nsteps, arena_bytes
        self.out_elems, self.out_conj, self.dtype, self.template = out_elems, out_conj, dtype, template
        self.graph = None


def _usable(d):
    return d.is_cuda


def _span(t):
    p = t.data_ptr()
    return p, p + t.numel() * t.element_size()


def _build(rec, ins, res):
    """_Chain of a recording whose data flow is closed, else None."""
    if rec.bad is not None or not rec.steps or res.numel() == 0:
        return None
    n_in = len(ins)
    ARENA, OUT = n_in, n_in + 1
    in_spans = [_span(t) for t in ins]
    if any(lo == hi for lo, hi in in_spans):
        return None
    out_ptr = res.data_ptr()
    made = []                    # per produced buffer: [lo, hi, birth step, last reading step]

    def locate(t, step):
        """(kind, index, byte offset) of the buffer holding tensor t: ('in', k, off) or ('buf', j, off)."""
        lo, hi = _span(t)
        if lo == hi:
            return None
        for j in range(len(made) - 1, -1, -1):          # latest producer first
            if made[j][0] <= lo and hi <= made[j][1]:
                made[j][3] = max(made[j][3], step)
                return ("buf", j, lo - made[j][0])
        for k, (ilo, ihi) in enumerate(in_spans):
            if ilo <= lo and hi <= ihi:
                return ("in", k, lo - ilo)
        return None

    flow = []
    for i, (kind, plan, (a, b, c), flags) in enumerate(rec.steps):
        la = locate(a, i)
        lb = locate(b, i) if kind == _lib.YB_CHAIN_GEMM else ("in", 0, 0)
        if la is None or lb is None or c.numel() == 0:
            return None
        clo, chi = _span(c)
        if any(lo < chi and clo < hi for lo, hi in in_spans):     # a launch writing into an operand: not a pure function
            return None
        made.append([clo, chi, i, i])
        flow.append((la, lb, len(made) - 1))
    # the result must be exactly the output of a recorded launch, and the last writer of that buffer
    out_buf = [j for j in range(len(made)) if made[j][0] == out_ptr and made[j][1] - made[j][0] == res.numel() * res.element_size()]
    if not out_buf:
        return None
    out_buf = out_buf[-1]
    # arena layout: buffers share space by lifetime (first fit over the blocks whose last reader has run)
    offs, free, live, end = {}, [], [], 0
    for j, (lo, hi, birth, death) in enumerate(made):
        for item in [x for x in live if x[0] < birth]:
            live.remove(item)
            free.append(item[1])
        if j == out_buf:
            continue
        need = -(-(hi - lo) // _ALIGN) * _ALIGN
        pick = next((f for f in free if f[1] >= need), None)
        if pick is not None:
            free.remove(pick)
            if pick[1] > need:
                free.append((pick[0] + need, pick[1] - need))
            offs[j] = pick[0]
        else:
            offs[j] = end
            end += need
        live.append((death, (offs[j], need)))
    table = np.zeros((len(flow), 10), dtype=np.int64)

    def slot(loc):
        kind, idx, off = loc
        if kind == "in":
            return idx, off
        if idx == out_buf:
            return OUT, off
        return ARENA, offs[idx] + off
    for i, ((la, lb, jc), (kind, plan, tens, flags)) in enumerate(zip(flow, rec.steps)):
        sa, oa = slot(la)
        sb, ob = slot(lb)
        sc, oc = slot(("buf", jc, 0))
        table[i] = (kind, flags, i, sa, oa, sb, ob, sc, oc, tens[2].numel())
    plans = [s[1] for s in rec.steps]
    return _NativeChain(table, plans, n_in + 2), plans, n_in, len(flow), end


def _sig(t):
    d = t._data
    return (t.struct, t.slices, t.hfs, t.mfs, t._trans, d.dtype, d.is_conj())


def _replay(ch, datas, dev):
    out = torch.empty(ch.out_elems, dtype=ch.dtype, device=dev)
    arena = torch.empty(max(ch.arena_bytes, 1), dtype=torch.uint8, device=dev)
    ptrs = [d.data_ptr() for d in datas] + [arena.data_ptr(), out.data_ptr()]
    _bk._on_device(dev, lambda st: ch.runner.run(ptrs, st))
    return out


def _replay_graph(ch, datas, dev):
    """Replay through a CUDA graph: operands are staged into static buffers, the graph holds all launches."""
    if ch.graph is None:
        static_in = [torch.empty_like(d) for d in datas]
        static_out = torch.empty(ch.out_elems, dtype=ch.dtype, device=dev)
        arena = torch.empty(max(ch.arena_bytes, 1), dtype=torch.uint8, device=dev)
        ptrs = [d.data_ptr() for d in static_in] + [arena.data_ptr(), static_out.data_ptr()]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            ch.runner.run(ptrs, torch.cuda.current_stream(dev).cuda_stream)
        ch.graph = (g, static_in, static_out, arena)
    g, static_in, static_out, _ = ch.graph
    for s, d in zip(static_in, datas):
        s.copy_(d)
    g.replay()
    return static_out.clone()


def trace(name, fn, tensors):
    """``fn(*tensors)`` for YASTN tensors, replayed from a recorded chain when the operand structures have been seen before."""
    datas = [t._data for t in tensors]
    d0 = datas[0]
    native = all(_usable(d) and d.dtype in _bk._DTYPE_CODE and d.is_contiguous() and d.device == d0.device for d in datas) and \
        not (torch.is_grad_enabled() and any(d.requires_grad for d in datas)) and _bk._recorder is None
    if not native:
        _stats["bypassed"] += 1
        return fn(*tensors)
    # operands that share their data (bra is ket in an expectation value) are one buffer to the recorder: the chain is only
    # valid for calls with the same sharing pattern, so the pattern is part of the key; partial overlaps are never chained
    spans = [_span(d) for d in datas]
    alias = tuple(next(j for j in range(i + 1) if spans[j] == spans[i]) for i in range(len(spans)))
    for i in range(len(spans)):
        for j in range(i):
            if alias[i] != alias[j] and spans[i][0] < spans[j][1] and spans[j][0] < spans[i][1]:
                _stats["bypassed"] += 1
                return fn(*tensors)
    key = (name, tensors[0].config, d0.device.index, alias) + tuple(_sig(t) for t in tensors)
    ch = _cache.get(key, False)
    if ch is None:               # known not to be replayable
        return fn(*tensors)
    if ch is not False:
        _cache.move_to_end(key)
        phys = [d.conj() if d.is_conj() else d for d in datas]
        out = (_replay_graph if _GRAPH else _replay)(ch, phys, d0.device)
        _stats["replayed"] += 1
        _stats["launches_replayed"] += ch.nsteps
        return ch.template._replace(data=out.conj() if ch.out_conj else out)
    rec = _Recorder()
    _bk._recorder = rec
    try:
        res = fn(*tensors)
    finally:
        _bk._recorder = None
    rd = res._data
    phys_in = [d.conj() if d.is_conj() else d for d in datas]
    built = _build(rec, phys_in, rd.conj() if rd.is_conj() else rd) if rd.dtype in _bk._DTYPE_CODE else None
    if built is None:
        _cache[key] = None
        _stats["rejected"] += 1
    else:
        runner, plans, n_in, nsteps, arena_bytes = built
        # the template carries the result's structure; its data is dropped (the recorded result itself goes to the caller)
        _cache[key] = _Chain(runner, plans, n_in, nsteps, arena_bytes, rd.numel(), rd.is_conj(), rd.dtype, res._replace(data=None))
        _stats["recorded"] += 1
    if len(_cache) > _MAX_CHAINS:
        _cache.popitem(last=False)
    return res


def clear():
    _cache.clear()


def stats():
    return dict(_stats, chains=sum(1 for v in _cache.values() if v is not None), rejected_cached=sum(1 for v in _cache.values() if v is None))


# -------------------------------------------------------------------------------------------------
# YASTN call sites (installed from outside, like enable_fused_tensordot)
# -------------------------------------------------------------------------------------------------
_saved = {}


def enable():
    """Route the tensordot sequences of ``yastn.tn.mps`` environments (yastn/tn/mps/_env.py:496-518) through chains.  The bodies
    below restate those reference lines — same contractions, same order, same axes — as functions of their operands."""
    if _saved:
        return
    import yastn.tn.mps._env as E
    from yastn import tensordot
    cls = E.Env_mps_mpo_mps
    for name in ("Heff1", "Heff2", "update_env_to_last", "update_env_to_first"):
        _saved[name] = getattr(cls, name)

    def heff2_body(AA, FR, W2, W1, FL):
        tmp = AA @ FR
        tmp = tensordot(W2, tmp, axes=((2, 3), (3, 2)))
        tmp = tensordot(W1, tmp, axes=((2, 3), (0, 3)))
        return tensordot(FL, tmp, axes=((0, 1), (3, 0)))

    def heff1_body(A, FR, W, FL):
        tmp = A @ FR
        tmp = tensordot(W, tmp, axes=((2, 3), (2, 1)))
        return tensordot(FL, tmp, axes=((0, 1), (2, 0)))

    def to_last_body(vecL, bra, W, ket):
        tmp = vecL @ bra.conj()
        tmp = tensordot(W, tmp, axes=((0, 1), (1, 2)))
        return tensordot(ket, tmp, axes=((0, 1), (2, 1)))

    def to_first_body(vecR, ket, W, bra):
        tmp = ket @ vecR
        tmp = tensordot(tmp, W, axes=((2, 1), (2, 3)))
        return tensordot(tmp, bra.conj(), axes=((3, 1), (1, 2)))

    def Heff2(self, AA, bd):
        n1, n2 = bd if bd[0] < bd[1] else bd[::-1]
        tmp = trace("Heff2", heff2_body, (AA, self.F[n2 + 1, n2], self.op.A[n2], self.op.A[n1], self.F[n1 - 1, n1]))
        return tmp * self.op.factor

    def Heff1(self, A, n):
        tmp = trace("Heff1", heff1_body, (A, self.F[n + 1, n], self.op.A[n], self.F[n - 1, n]))
        return tmp * self.op.factor

    def update_env_to_last(self, vecL, n):
        return trace("to_last", to_last_body, (vecL, self.bra.A[n], self.op.A[n], self.ket.A[n]))

    def update_env_to_first(self, vecR, n):
        return trace("to_first", to_first_body, (vecR, self.ket.A[n], self.op.A[n], self.bra.A[n]))

    cls.Heff2, cls.Heff1 = Heff2, Heff1
    cls.update_env_to_last, cls.update_env_to_first = update_env_to_last, update_env_to_first


_PEPS_FNS = ("append_vec_tl", "append_vec_br", "append_vec_tr", "append_vec_bl")
_saved_peps = {}


def enable_peps():
    """Chains for the double-layer contractions of CTMRG: ``append_vec_tl / _br / _tr / _bl``
    (yastn/tn/fpeps/envs/_env_contractions.py:211-365, reached from ``DoublePepsTensor.tensordot``,
    yastn/tn/fpeps/_doublePepsTensor.py:246-283) attach the bra and ket PEPS tensors to a corner / edge vector through fuse_legs,
    unfuse_legs, two tensordots and a final fuse — ~12 backend calls walked in Python for operands whose block structure is the
    same in every CTM sweep once the environment dimension has saturated.  The reference functions themselves are traced
    (nothing is restated); calls with an operator inserted (``op``) run as written, and on fermionic lattices the swap gates in
    between are elementwise plans, so those recordings are rejected and the functions keep running as written."""
    if _saved_peps:
        return
    import inspect
    import sys
    import yastn.tn.fpeps.envs._env_contractions as EC
    import yastn.tn.fpeps  # noqa: F401  (loads the modules that imported the functions by name)
    for name in _PEPS_FNS:
        orig = getattr(EC, name)
        sig = inspect.signature(orig)

        def make(orig, sig, name):
            def wrapped(*args, **kwargs):
                b = sig.bind(*args, **kwargs)
                b.apply_defaults()
                a = b.arguments
                if a.get("op") is not None:
                    return orig(*args, **kwargs)
                mode, in_b, out_a = a["mode"], tuple(a["in_b"]), tuple(a["out_a"])
                keys = list(a)[:3]          # (Ac, A, vector)
                return trace((name, mode, in_b, out_a), lambda x, y, v: orig(x, y, v, op=None, mode=mode, in_b=in_b, out_a=out_a),
                             tuple(a[k] for k in keys))
            wrapped.__wrapped__ = orig
            wrapped.__doc__ = orig.__doc__
            return wrapped
        new = make(orig, sig, name)
        _saved_peps[name] = (orig, new)
    for mod in list(sys.modules.values()):
        if mod is None or not getattr(mod, "__name__", "").startswith("yastn.tn.fpeps"):
            continue
        for name, (orig, new) in _saved_peps.items():
            if getattr(mod, name, None) is orig:
                setattr(mod, name, new)


def disable():
    if _saved:
        import yastn.tn.mps._env as E
        for name, fn in _saved.items():
            setattr(E.Env_mps_mpo_mps, name, fn)
        _saved.clear()
    if _saved_peps:
        import sys
        for mod in list(sys.modules.values()):
            if mod is None or not getattr(mod, "__name__", "").startswith("yastn.tn.fpeps"):
                continue
            for name, (orig, new) in _saved_peps.items():
                if getattr(mod, name, None) is new:
                    setattr(mod, name, orig)
        _saved_peps.clear()
    clear()
