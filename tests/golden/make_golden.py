"""Generate golden fixtures from the UNMODIFIED reference (run in the authoring container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.json.gz, *.npz

The reference (/root/reference, yastn @ 60af786a) is imported read-only with a 2-file stub for the
missing ``opt_einsum`` package (only ``import yastn`` needs it; nothing on the hot path uses it).
Its numpy backend is wrapped by a recorder, so that every ``transpose_and_merge`` / ``dot`` /
``unmerge`` / ``transpose`` / ``transpose_dot_sum`` call made by ``yastn.tensordot``, ``fuse_legs``,
``unfuse_legs`` and ``consume_transpose`` is captured with its exact arguments and result.

Outputs
  calls_small.json.gz + calls_small.npz   recorded backend calls (metas + input/output data) of small
                                          cases: dense/U1/Z2/U1xU1/Z2xU1, fp64/c128, all 3 policies.
  structs_bench.json.gz                   structure-only fixtures (block tables of operands, recorded
                                          metas per policy, result struct/slices) for the benchmark
                                          workloads of SURVEY.md section 8(d); data is synthetic at run time.
"""
import gzip
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("YASTN_REF", "/root/reference")


def _import_reference():
    stub = tempfile.mkdtemp(prefix="oe_stub_")
    os.makedirs(os.path.join(stub, "opt_einsum"))
    with open(os.path.join(stub, "opt_einsum", "__init__.py"), "w") as f:
        f.write("from . import contract\n")
    with open(os.path.join(stub, "opt_einsum", "contract.py"), "w") as f:
        f.write("class PathInfo:\n    pass\n_VALID_CONTRACT_KWARGS = set()\n")
    sys.path[:0] = [stub, REF]
    import yastn  # noqa
    return yastn


yastn = _import_reference()
import yastn.backend.backend_np as ref_np  # noqa: E402

HOT = ("transpose_and_merge", "dot", "unmerge", "transpose", "transpose_dot_sum")


def _plain(x):
    """Nested tuples/lists/numpy ints -> nested lists of Python ints."""
    if isinstance(x, (tuple, list)):
        return [_plain(y) for y in x]
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, np.ndarray):
        return _plain(x.tolist())
    if hasattr(x, "_asdict"):  # _slc / _struct NamedTuples
        return {k: _plain(v) for k, v in x._asdict().items()}
    return x


class Recorder:
    """Wraps the five hot functions of the reference numpy backend."""

    ARGNAMES = {
        "transpose_and_merge": ("data", "order", "meta_new", "meta_mrg", "Dsize"),
        "dot": ("Adata", "Bdata", "meta_dot", "Dsize"),
        "unmerge": ("data", "meta"),
        "transpose": ("data", "axes", "meta_transpose"),
        "transpose_dot_sum": ("Adata", "Bdata", "meta_dot", "Areshape", "Breshape", "Aorder", "Border", "Dsize"),
    }

    def __init__(self):
        self.calls = []
        self.orig = {}

    def __enter__(self):
        for name in HOT:
            self.orig[name] = getattr(ref_np, name)
            setattr(ref_np, name, self._wrap(name))
        return self

    def __exit__(self, *exc):
        for name in HOT:
            setattr(ref_np, name, self.orig[name])

    def _wrap(self, name):
        def f(*args):
            out = self.orig[name](*args)
            self.calls.append((name, args, out))
            return out
        return f


def struct_dict(a):
    return {"s": _plain(a.struct.s), "n": _plain(a.struct.n), "t": _plain(a.struct.t), "D": _plain(a.struct.D),
            "size": int(a.struct.size), "slices": [list(x.slcs[0]) for x in a.slices],
            "trans": _plain(a.trans), "nsym": int(a.config.sym.NSYM), "sym": a.config.sym.SYM_ID}


# ----------------------------------------------------------------------------------------------
# small data-carrying cases
# ----------------------------------------------------------------------------------------------

def small_cases():
    """Yield (name, callable(cfg) -> None) exercising the hot path on the reference."""
    def u1_r4(cfg, dtype):
        a = yastn.rand(config=cfg, s=(-1, 1, 1, -1), t=((-1, 1, 2), (-1, 1, 2), (-1, 1, 2), (-1, 1, 2)),
                       D=((1, 2, 3), (4, 5, 6), (3, 2, 4), (5, 3, 2)), dtype=dtype)
        b = yastn.rand(config=cfg, s=(1, -1, 1), t=((-1, 1, 2), (-1, 1, 2), (-1, 0, 1)),
                       D=((1, 2, 3), (4, 5, 6), (3, 7, 5)), dtype=dtype)
        yastn.tensordot(a, b, axes=((0, 1), (0, 1)))
        yastn.tensordot(a, b, axes=(0, 0))
        yastn.tensordot(b, a, axes=((1, 0), (1, 0)), conj=(1, 1))
        yastn.tensordot(a.transpose((3, 1, 0, 2)), b.transpose((2, 0, 1)), axes=((2, 1), (1, 2)))
        # outer product (one K=1 GEMM per sector) on small operands
        c = yastn.rand(config=cfg, s=(-1, 1), t=((-1, 0, 1), (-1, 0, 1)), D=((2, 3, 4), (2, 3, 4)), dtype=dtype)
        d = yastn.rand(config=cfg, s=(1, -1, 1), t=((0, 1), (-1, 1), (0, 1, 2)), D=((2, 3), (4, 2), (1, 2, 3)), dtype=dtype)
        yastn.tensordot(c, d, axes=((), ()))

    def u1_missing(cfg, dtype):
        # charges present on one side only -> _common_inds filters blocks and the merge pads with zeros
        a = yastn.rand(config=cfg, s=(-1, 1, 1), t=((-2, 0, 2), (-1, 1), (-3, -1, 1, 3)), D=((2, 3, 4), (5, 6), (1, 2, 3, 4)), dtype=dtype)
        b = yastn.rand(config=cfg, s=(-1, 1, 1), t=((-3, -1, 5), (0, 1), (-1, 0, 1)), D=((1, 2, 7), (3, 2), (2, 3, 4)), dtype=dtype)
        yastn.tensordot(a, b, axes=(2, 0))
        yastn.tensordot(b, a, axes=(0, 2))
        a2 = yastn.Tensor(config=cfg, s=(-1, 1, 1, -1), dtype=dtype)
        a2.set_block(ts=(1, 1, 0, 0), Ds=(2, 3, 4, 5), val='rand')
        a2.set_block(ts=(1, 0, 1, 0), Ds=(2, 6, 7, 5), val='rand')
        a2.set_block(ts=(-1, -1, 2, 2), Ds=(3, 4, 8, 9), val='rand')
        b2 = yastn.Tensor(config=cfg, s=(1, -1, 1), dtype=dtype)
        b2.set_block(ts=(1, 1, 0), Ds=(2, 3, 5), val='rand')
        b2.set_block(ts=(1, 0, -1), Ds=(2, 6, 2), val='rand')
        b2.set_block(ts=(2, 2, 0), Ds=(7, 8, 5), val='rand')
        yastn.tensordot(a2, b2, axes=((0, 1), (0, 1)))
        yastn.tensordot(a2, b2, axes=(0, 0))
        # empty result
        a3 = yastn.rand(config=cfg, s=(-1, 1), t=((0,), (0,)), D=((2,), (3,)), dtype=dtype)
        b3 = yastn.rand(config=cfg, s=(-1, 1), t=((1,), (1,)), D=((4,), (5,)), dtype=dtype)
        yastn.tensordot(a3, b3, axes=(1, 0))

    def fuse_cases(cfg, dtype):
        a = yastn.rand(config=cfg, s=(-1, 1, 1, -1, 1), t=((-1, 0, 1), (0, 1), (-1, 1), (0, 1, 2), (-1, 0)),
                       D=((2, 3, 2), (3, 2), (2, 4), (1, 2, 3), (2, 2)), dtype=dtype)
        f = a.fuse_legs(axes=((0, 2), 1, (4, 3)), mode='hard')
        f.unfuse_legs(axes=(0, 2))
        f2 = f.fuse_legs(axes=((0, 1), 2), mode='hard')
        f2.unfuse_legs(axes=0).unfuse_legs(axes=(0, 2))
        a.transpose((4, 2, 0, 3, 1)).consume_transpose()
        b = yastn.rand(config=cfg, s=(1, -1, -1), t=((-1, 0, 1), (-1, 1), (0, 1)), D=((2, 3, 2), (2, 4), (3, 2)), dtype=dtype)
        yastn.tensordot(f, b.fuse_legs(axes=((0, 1), 2), mode='hard'), axes=(0, 0))

    def dense(cfg, dtype):
        a = yastn.rand(config=cfg, s=(-1, 1, 1, -1), D=(2, 3, 4, 5), dtype=dtype)
        b = yastn.rand(config=cfg, s=(1, -1, 1), D=(2, 3, 5), dtype=dtype)
        yastn.tensordot(a, b, axes=((0, 3), (0, 2)))
        yastn.tensordot(b, a, axes=((2, 0), (3, 0)), conj=(1, 1))
        yastn.tensordot(a, b, axes=((), ()))
        a.transpose((2, 0, 3, 1)).consume_transpose()

    def z2xu1(cfg, dtype):
        t1 = ((0, -1), (0, 1), (1, -1), (1, 1))
        a = yastn.rand(config=cfg, s=(-1, 1, 1, -1), t=(t1, t1, t1, t1), D=((1, 2, 2, 4), (9, 4, 3, 2), (5, 6, 7, 8), (7, 8, 9, 10)), dtype=dtype)
        b = yastn.rand(config=cfg, s=(1, -1, 1), t=(t1, t1, t1), D=((1, 2, 2, 4), (9, 4, 3, 2), (4, 5, 6, 3)), dtype=dtype)
        yastn.tensordot(a, b, axes=((0, 1), (0, 1)))
        yastn.tensordot(a.transpose((1, 3, 0, 2)), b, axes=((2, 0), (0, 1)))

    def u1xu1(cfg, dtype):
        L = yastn.gaussian_leg(cfg, s=1, n=(0, 0), sigma=1.0, D_total=24, method='round')
        p = yastn.Leg(cfg, s=1, t=((0, 0), (1, 0), (0, 1), (1, 1)), D=(1, 1, 1, 1))
        A = yastn.rand(cfg, legs=[L.conj(), p, L], n=(0, 0), dtype=dtype)
        W = yastn.rand(cfg, legs=[L.conj(), p.conj(), p, L], n=(0, 0), dtype=dtype)
        yastn.tensordot(A, W, axes=(2, 0))
        yastn.tensordot(A, A, axes=((0, 1), (0, 1)), conj=(1, 0))

    def z2(cfg, dtype):
        L = yastn.Leg(cfg, s=1, t=(0, 1), D=(7, 9))
        p = yastn.Leg(cfg, s=1, t=(0, 1), D=(1, 1))
        A = yastn.rand(cfg, legs=[L.conj(), p, p, L], n=0, dtype=dtype)
        F = yastn.rand(cfg, legs=[L.conj(), p, L], n=1, dtype=dtype)
        yastn.tensordot(A, F, axes=(3, 0))
        yastn.tensordot(A.transpose((2, 0, 3, 1)), A, axes=((1, 2), (3, 0)))

    return [("U1", "u1_r4", u1_r4), ("U1", "u1_missing", u1_missing), ("U1", "fuse_cases", fuse_cases),
            ("none", "dense", dense), ("Z2xU1", "z2xu1", z2xu1), ("U1xU1", "u1xu1", u1xu1), ("Z2", "z2", z2)]


def make_small():
    index, arrays = [], {}
    for sym, name, fn in small_cases():
        for policy in ("fuse_to_matrix", "fuse_contracted", "no_fusion"):
            for dtype in ("float64", "complex128"):
                if dtype == "complex128" and name in ("fuse_cases", "z2xu1", "z2"):
                    continue
                if policy != "fuse_to_matrix" and name == "fuse_cases":
                    continue
                sym_ = {"none": "dense", "Z2xU1": yastn.sym.sym_Z2xU1}.get(sym, sym)
                cfg = yastn.make_config(sym=sym_, backend='np', tensordot_policy=policy, default_fusion='hard')
                cfg.backend.random_seed(11)
                yastn.clear_cache() if hasattr(yastn, "clear_cache") else None
                with Recorder() as rec:
                    fn(cfg, dtype)
                for fname, args, out in rec.calls:
                    k = len(index)
                    entry = {"fn": fname, "case": name, "sym": sym, "policy": policy, "dtype": dtype, "args": {}}
                    for an, av in zip(Recorder.ARGNAMES[fname], args):
                        if isinstance(av, np.ndarray):
                            arrays[f"c{k}_{an}"] = av
                            entry["args"][an] = f"@c{k}_{an}"
                        else:
                            entry["args"][an] = _plain(av)
                    arrays[f"c{k}_out"] = out
                    index.append(entry)
    with gzip.open(os.path.join(HERE, "calls_small.json.gz"), "wt") as f:
        json.dump(index, f, separators=(",", ":"))
    np.savez_compressed(os.path.join(HERE, "calls_small.npz"), **arrays)
    by_fn = {}
    for e in index:
        by_fn[e["fn"]] = by_fn.get(e["fn"], 0) + 1
    print("calls_small:", len(index), by_fn, "arrays bytes", sum(a.nbytes for a in arrays.values()))


# ----------------------------------------------------------------------------------------------
# structure-only fixtures of the benchmark workloads
# ----------------------------------------------------------------------------------------------

def _record_tensordot(a, b, axes, policies=("fuse_to_matrix", "fuse_contracted", "no_fusion")):
    """Run tensordot on structure-identical tiny-dtype... no: on the real tensors; record metas only."""
    out = {"a": struct_dict(a), "b": struct_dict(b), "axes": _plain(axes)}
    for policy in policies:
        cfg = a.config._replace(tensordot_policy=policy)
        a1, b1 = a._replace(config=cfg), b._replace(config=cfg)
        with Recorder() as rec:
            c = yastn.tensordot(a1, b1, axes=axes)
        names = [x[0] for x in rec.calls]
        key = {"fuse_to_matrix": "f2m", "fuse_contracted": "fc", "no_fusion": "nf"}[policy]
        st = {"struct_c": struct_dict(c)}
        if policy == "no_fusion":
            (_, args, _), = rec.calls
            st["tds"] = {k: _plain(v) for k, v in zip(Recorder.ARGNAMES["transpose_dot_sum"][2:], args[2:])}
        else:
            # identify which merge calls happened: a's merge comes first; fast-path skips leave no call.
            merges = [c_ for c_ in rec.calls if c_[0] == "transpose_and_merge"]
            dots = [c_ for c_ in rec.calls if c_[0] == "dot"]
            unm = [c_ for c_ in rec.calls if c_[0] == "unmerge"]
            assert len(dots) == 1 and len(unm) <= 1 and names.index("dot") == len(merges)
            na, nb = len(a1._data), len(b1._data)
            sides = {"merge_a": None, "merge_b": None}
            for _, args, res in merges:
                rec_ = {"order": _plain(args[1]), "meta_new": _plain(args[2]), "meta_mrg": _plain(args[3]), "Dsize": int(args[4])}
                side = "merge_a" if (args[0] is a1._data and sides["merge_a"] is None) else "merge_b"
                sides[side] = rec_
            st.update(sides)
            st["dot"] = {"meta_dot": _plain(dots[0][1][2]), "Dsize": int(dots[0][1][3])}
            st["unmerge"] = {"meta": _plain(unm[0][1][1])} if unm else None
        out[key] = st
    return out


def bench_structs():
    cases = {}

    def synth(sym, n0, sigma, D, tag, pt, pD, wt, wD, policies=("fuse_to_matrix", "fuse_contracted", "no_fusion")):
        cfg = yastn.make_config(sym=sym, backend='np')
        cfg.backend.random_seed(0)
        L = yastn.gaussian_leg(cfg, s=1, n=n0, sigma=sigma, D_total=D, method='round')
        p = yastn.Leg(cfg, s=1, t=pt, D=pD)
        w = yastn.Leg(cfg, s=1, t=wt, D=wD)
        # structure only: tiny dtype is not available in yastn, so build with zeros of float64 lazily
        A = yastn.zeros(cfg, legs=[L.conj(), p, p, L], n=n0)
        F = yastn.zeros(cfg, legs=[L.conj(), w, L], n=n0)
        B4 = yastn.zeros(cfg, legs=[L.conj(), p.conj(), p.conj(), L], n=n0)
        cases[f"{tag}_P1"] = _record_tensordot(A, F, (3, 0), policies)
        cases[f"{tag}_P2"] = _record_tensordot(A, B4, ((1, 2, 3), (2, 1, 0)), policies)
        cases[f"{tag}_P3"] = _record_tensordot(A.transpose((2, 0, 3, 1)), B4, ((1, 2), (3, 0)), policies)
        print(tag, "sectors", len(L.t), "blocks A/F/B4", len(A.struct.t), len(F.struct.t), len(B4.struct.t),
              "sizes", A.size, F.size, B4.size)

    u1 = dict(pt=(-1, 1), pD=(1, 1), wt=(-2, 0, 2), wD=(1, 3, 1))
    synth('U1', 0, 0.7, 64, "U1_D64", **u1)
    synth('U1', 0, 1.0, 1024, "U1_D1024", **u1)
    synth('U1', 0, 1.5, 2048, "U1_D2048", **u1)
    synth('U1', 0, 2.5, 4096, "U1_D4096", **u1)
    synth('U1', 0, 4.0, 8192, "U1_D8192", policies=("fuse_to_matrix",), **u1)
    synth('U1', 0, 6.0, 16384, "U1_D16384", policies=("fuse_to_matrix",), **u1)
    synth('Z2', 0, 1.0, 512, "Z2_D512", pt=(0, 1), pD=(1, 1), wt=(0, 1), wD=(2, 2))
    synth('U1xU1', (0, 0), 2.0, 4096, "U1xU1_D4096", policies=("fuse_to_matrix",),
          pt=((0, 0), (1, 0), (0, 1), (1, 1)), pD=(1, 1, 1, 1), wt=((0, 0), (1, 0), (-1, 0), (0, 1), (0, -1)), wD=(2, 1, 1, 1, 1))
    with gzip.open(os.path.join(HERE, "structs_bench.json.gz"), "wt") as f:
        json.dump(cases, f, separators=(",", ":"))
    print("structs_bench:", len(cases), "cases,", os.path.getsize(os.path.join(HERE, "structs_bench.json.gz")), "bytes")


def chain_structs():
    """Two consecutive contractions that group the shared tensor by DIFFERENT legs (multi-GPU redistribution fixture,
    SURVEY 8e): step 1  C = tensordot(A, F, (3, 0)) shards by the charge of A's right leg, step 2  E = tensordot(G, C, (2, 0))
    shards by the charge of C's left leg, so the blocks of C change owner between the two."""
    cases = {}
    for sigma, D in ((1.0, 1024), (2.5, 4096)):
        cfg = yastn.make_config(sym='U1', backend='np')
        L = yastn.gaussian_leg(cfg, s=1, n=0, sigma=sigma, D_total=D, method='round')
        p = yastn.Leg(cfg, s=1, t=(-1, 1), D=(1, 1))
        w = yastn.Leg(cfg, s=1, t=(-2, 0, 2), D=(1, 3, 1))
        A = yastn.zeros(cfg, legs=[L.conj(), p, p, L], n=0)
        F = yastn.zeros(cfg, legs=[L.conj(), w, L], n=0)
        G = yastn.zeros(cfg, legs=[L.conj(), w, L], n=0)
        C = yastn.tensordot(A, F, axes=(3, 0))
        cases[f"U1_D{D}_chain"] = {"step1": _record_tensordot(A, F, (3, 0), ("fuse_to_matrix",)),
                                   "step2": _record_tensordot(G, C, (2, 0), ("fuse_to_matrix",))}
        print("chain", D, "C blocks", len(C.struct.t), "size", C.size)
    with gzip.open(os.path.join(HERE, "structs_chain.json.gz"), "wt") as f:
        json.dump(cases, f, separators=(",", ":"))
    print("structs_chain:", os.path.getsize(os.path.join(HERE, "structs_chain.json.gz")), "bytes")


def transpose_structs():
    """Contractions whose merge is a genuine transposition of large blocks (copy-kernel tiled path at scale):
    T1 = tensordot(A, conj(F), axes=(0, 0)) merges A[l*,p,p,r] as (p p r) x l, i.e. every (Dl x Dr) block is transposed."""
    cases = {}
    for sigma, D in ((2.5, 4096), (6.0, 16384)):
        cfg = yastn.make_config(sym='U1', backend='np')
        L = yastn.gaussian_leg(cfg, s=1, n=0, sigma=sigma, D_total=D, method='round')
        p = yastn.Leg(cfg, s=1, t=(-1, 1), D=(1, 1))
        w = yastn.Leg(cfg, s=1, t=(-2, 0, 2), D=(1, 3, 1))
        A = yastn.zeros(cfg, legs=[L.conj(), p, p, L], n=0)
        F = yastn.zeros(cfg, legs=[L.conj(), w, L], n=0)
        cases[f"U1_D{D}_T1"] = _record_tensordot(A, F.conj(), (0, 0), ("fuse_to_matrix",))
        print("transpose", D, "order", cases[f"U1_D{D}_T1"]["f2m"]["merge_a"]["order"])
    with gzip.open(os.path.join(HERE, "structs_transpose.json.gz"), "wt") as f:
        json.dump(cases, f, separators=(",", ":"))
    print("structs_transpose:", os.path.getsize(os.path.join(HERE, "structs_transpose.json.gz")), "bytes")


if __name__ == "__main__":
    import sys
    if "--transpose-only" in sys.argv:
        transpose_structs()
    elif "--chain-only" in sys.argv:
        chain_structs()
    else:
        make_small()
        bench_structs()
        chain_structs()
        transpose_structs()
