"""Golden calls of the reference's block-wise vector operations (run in the authoring container only).

    python tests/golden/make_golden_ewise.py        # writes tests/golden/calls_ewise.json.gz, calls_ewise.npz

Same recipe as make_golden.py: the UNMODIFIED reference (/root/reference, numpy backend) is driven through its public API
(``a + b``, ``a - b``, ``yastn.add``, ``swap_gate``, ``broadcast``, ``apply_mask``, ``trace``, ``svd_with_truncation`` ...) while
a recorder captures every call of  add / sub / negate_blocks / dot_diag / apply_mask / embed_mask / trace  with its exact
arguments and result (yastn/backend/backend_np.py; the torch twins are backend_torch.py:268-275,518-534,557-564 and
_backend_torch_backwards.py:229-310).  ``mask`` arguments are dicts {charge: index array}; they are stored as a list of
(charge, array) pairs.
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import yastn, ref_np, _plain  # noqa: E402  (imports the reference with the opt_einsum stub)

OPS = {
    "add": ("datas", "metas", "Dsize"),
    "sub": ("Adata", "Bdata", "metas", "Dsize"),
    "negate_blocks": ("Adata", "slices"),
    "dot_diag": ("Adata", "Bdata", "meta", "Dsize", "axis", "a_ndim"),
    "apply_mask": ("Adata", "mask", "meta", "Dsize", "axis", "ndim"),
    "embed_mask": ("Adata", "mask", "meta", "Dsize", "axis", "ndim"),
    "trace": ("data", "order", "meta", "Dsize"),
}


class Recorder:
    def __init__(self):
        self.calls, self.orig = [], {}

    def __enter__(self):
        for name in OPS:
            self.orig[name] = getattr(ref_np, name)
            setattr(ref_np, name, self._wrap(name))
        return self

    def __exit__(self, *exc):
        for name in OPS:
            setattr(ref_np, name, self.orig[name])

    def _wrap(self, name):
        def f(*args):
            out = self.orig[name](*args)
            self.calls.append((name, args, np.array(out, copy=True)))
            return out
        return f


def cases():
    def algebra(cfg, dtype):
        leg = yastn.Leg(cfg, s=1, t=(-1, 0, 1), D=(2, 3, 4))
        a = yastn.rand(cfg, legs=[leg.conj(), leg, leg], dtype=dtype)
        b = yastn.rand(cfg, legs=[leg.conj(), leg, leg], dtype=dtype)
        _ = a + b
        _ = a - b
        # different block structures: union of blocks, holes on either side
        c = yastn.Tensor(config=cfg, s=(-1, 1, 1), dtype=dtype)
        c.set_block(ts=(0, 0, 0), Ds=(3, 3, 3), val='rand')
        c.set_block(ts=(1, 0, 1), Ds=(4, 3, 4), val='rand')
        d = yastn.Tensor(config=cfg, s=(-1, 1, 1), dtype=dtype)
        d.set_block(ts=(1, 0, 1), Ds=(4, 3, 4), val='rand')
        d.set_block(ts=(-1, 0, -1), Ds=(2, 3, 2), val='rand')
        d.set_block(ts=(1, 1, 0), Ds=(4, 4, 3), val='rand')
        _ = c + d
        _ = c - d
        _ = d - c
        # linear combination of 6 tensors (two launches of four sources) with mixed structures
        _ = yastn.add(a, b, c, d, a, c, amplitudes=(1, -2, 0.5, None, 3, 1))
        _ = yastn.add(a, b, a, amplitudes=(0.5, 0.25, None))
        # diagonal tensors
        e = yastn.rand(cfg, legs=[leg.conj(), leg], isdiag=True, dtype=dtype)
        f = yastn.rand(cfg, legs=[leg.conj(), leg], isdiag=True, dtype=dtype)
        _ = e + f
        _ = e - f

    def diag_and_mask(cfg, dtype):
        l1 = yastn.Leg(cfg, s=1, t=(-1, 0, 1), D=(5, 6, 7))
        l2 = yastn.Leg(cfg, s=1, t=(-1, 0, 2), D=(2, 3, 4))
        a = yastn.rand(cfg, legs=[l1.conj(), l2, l1, l2.conj()], dtype=dtype)
        d = yastn.rand(cfg, legs=[l1.conj(), l1], isdiag=True, dtype=dtype)
        for ax in (0, 2):
            yastn.broadcast(d, a, axes=ax)
        yastn.broadcast(d, a.transpose((3, 2, 1, 0)), axes=1)
        dd = yastn.rand(cfg, legs=[l1.conj(), l1], isdiag=True, dtype=dtype)
        yastn.broadcast(d, dd, axes=0)
        yastn.tensordot(a, d, axes=(2, 0))          # contraction with a diagonal tensor goes through broadcast
        # masks: keep entries above a threshold
        m = yastn.rand(cfg, legs=[l1.conj(), l1], isdiag=True, dtype='float64')
        mask = m > 0.1 if hasattr(m, '__gt__') else m
        for ax in (0, 2):
            yastn.apply_mask(mask, a, axes=ax)
        yastn.apply_mask(mask, a.transpose((2, 1, 0, 3)), axes=0)
        yastn.apply_mask(mask, dd, axes=0)
        yastn.apply_mask(mask, a, a, axes=(0, 2))

    def truncation(cfg, dtype):
        l1 = yastn.Leg(cfg, s=1, t=(-1, 0, 1), D=(5, 6, 7))
        l2 = yastn.Leg(cfg, s=1, t=(-1, 0, 1), D=(2, 3, 4))
        a = yastn.rand(cfg, legs=[l1.conj(), l2, l1, l2.conj()], dtype=dtype)
        yastn.svd_with_truncation(a, axes=((0, 1), (2, 3)), D_total=11)
        yastn.svd_with_truncation(a, axes=((0, 1), (2, 3)), D_block=3)

    def embed(cfg, dtype):
        # addition of tensors whose fused legs differ -> _embed_tensor -> embed_mask
        a = yastn.Tensor(config=cfg, s=(-1, 1, 1, -1), dtype=dtype)
        a.set_block(ts=(1, 1, 0, 0), Ds=(2, 3, 4, 5), val='rand')
        a.set_block(ts=(1, 0, 1, 0), Ds=(2, 6, 7, 5), val='rand')
        b = yastn.Tensor(config=cfg, s=(-1, 1, 1, -1), dtype=dtype)
        b.set_block(ts=(1, 1, 0, 0), Ds=(2, 3, 4, 5), val='rand')
        b.set_block(ts=(-1, -1, 2, 2), Ds=(3, 4, 8, 9), val='rand')
        fa = a.fuse_legs(axes=((0, 1), (2, 3)), mode='hard')
        fb = b.fuse_legs(axes=((0, 1), (2, 3)), mode='hard')
        _ = fa + fb
        _ = fa - fb
        yastn.vdot(fa, fb)

    def traces(cfg, dtype):
        l1 = yastn.Leg(cfg, s=1, t=(-1, 0, 1), D=(2, 3, 4))
        l2 = yastn.Leg(cfg, s=1, t=(0, 1), D=(5, 2))
        a = yastn.rand(cfg, legs=[l1.conj(), l2, l1, l2.conj(), l2], dtype=dtype)
        yastn.trace(a, axes=(0, 2))
        yastn.trace(a, axes=(3, 1))
        yastn.trace(a, axes=((0, 1), (2, 3)))
        yastn.trace(a.transpose((4, 3, 2, 1, 0)), axes=((1, 2), (3, 4)))
        b = yastn.rand(cfg, legs=[l1.conj(), l1], dtype=dtype)
        yastn.trace(b, axes=(0, 1))

    def fermions(cfg, dtype):
        l1 = yastn.Leg(cfg, s=1, t=(0, 1), D=(3, 4))
        l2 = yastn.Leg(cfg, s=1, t=(0, 1), D=(2, 5))
        a = yastn.rand(cfg, legs=[l1.conj(), l2, l1, l2.conj()], n=1, dtype=dtype)
        a.swap_gate(axes=(0, 1))
        a.swap_gate(axes=((0, 1), (2, 3)))
        a.swap_gate(axes=(1, 3), charge=(1,))
        a.transpose((3, 1, 0, 2)).swap_gate(axes=(0, 2, 1, 3))

    return [("U1", False, "algebra", algebra), ("U1", False, "diag_and_mask", diag_and_mask), ("U1", False, "truncation", truncation),
            ("U1", False, "embed", embed), ("U1", False, "traces", traces), ("Z2", True, "fermions", fermions),
            ("U1xU1", False, "algebra_u1u1", None)]


def algebra_u1u1(cfg, dtype):
    L = yastn.gaussian_leg(cfg, s=1, n=(0, 0), sigma=1.0, D_total=24, method='round')
    p = yastn.Leg(cfg, s=1, t=((0, 0), (1, 0), (0, 1), (1, 1)), D=(1, 1, 1, 1))
    A = yastn.rand(cfg, legs=[L.conj(), p, L], n=(0, 0), dtype=dtype)
    B = yastn.rand(cfg, legs=[L.conj(), p, L], n=(0, 0), dtype=dtype)
    _ = yastn.add(A, B, A, amplitudes=(0.3, None, -1.5))
    d = yastn.rand(cfg, legs=[L.conj(), L], isdiag=True, dtype=dtype)
    yastn.broadcast(d, A, axes=2)
    yastn.trace(yastn.tensordot(A, A, axes=(1, 1), conj=(0, 1)), axes=((0, 1), (2, 3)))


def _store(val, key, arrays):
    """JSON-able form of an argument; arrays go to the npz under `key`."""
    if isinstance(val, np.ndarray):
        arrays[key] = val
        return "@" + key
    if isinstance(val, dict):       # mask: {charge: index array}
        out = []
        for i, (k, v) in enumerate(val.items()):
            arrays[f"{key}_m{i}"] = np.asarray(v, dtype=np.int64)
            out.append([_plain(k), "@" + f"{key}_m{i}"])
        return {"__mask__": out}
    if isinstance(val, (list, tuple)) and len(val) and all(isinstance(v, np.ndarray) for v in val):
        out = []
        for i, v in enumerate(val):
            arrays[f"{key}_{i}"] = v
            out.append("@" + f"{key}_{i}")
        return {"__arrays__": out}
    return _plain(val)


def main():
    index, arrays = [], {}
    for sym, fermionic, name, fn in cases():
        fn = fn or algebra_u1u1
        for dtype in ("float64", "complex128"):
            cfg = yastn.make_config(sym=sym, backend='np', fermionic=fermionic, default_fusion='hard')
            cfg.backend.random_seed(5)
            yastn.clear_cache()
            with Recorder() as rec:
                fn(cfg, dtype)
            for fname, args, out in rec.calls:
                k = len(index)
                entry = {"fn": fname, "case": name, "sym": sym, "dtype": dtype, "args": {}}
                for an, av in zip(OPS[fname], args):
                    entry["args"][an] = _store(av, f"c{k}_{an}", arrays)
                arrays[f"c{k}_out"] = out
                index.append(entry)
    with gzip.open(os.path.join(HERE, "calls_ewise.json.gz"), "wt") as f:
        json.dump(index, f, separators=(",", ":"))
    np.savez_compressed(os.path.join(HERE, "calls_ewise.npz"), **arrays)
    by_fn = {}
    for e in index:
        by_fn[e["fn"]] = by_fn.get(e["fn"], 0) + 1
    print("calls_ewise:", len(index), by_fn, "array bytes", sum(a.nbytes for a in arrays.values()))


if __name__ == "__main__":
    main()
