"""Loaders for the committed golden fixtures (tests/golden/, made by make_golden.py from the reference)."""
import gzip
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def tup(x):
    """JSON lists -> nested tuples (the reference hands tuples of Python ints to the backend)."""
    if isinstance(x, list):
        return tuple(tup(y) for y in x)
    return x


_cache = {}


def small_calls():
    """List of recorded backend calls: dict(fn, case, sym, policy, dtype, args{name: value|ndarray}, out)."""
    if "small" not in _cache:
        with gzip.open(os.path.join(GOLDEN, "calls_small.json.gz"), "rt") as f:
            index = json.load(f)
        arrays = np.load(os.path.join(GOLDEN, "calls_small.npz"))
        calls = []
        for k, e in enumerate(index):
            args = {}
            for name, val in e["args"].items():
                args[name] = arrays[val[1:]] if isinstance(val, str) and val.startswith("@") else tup(val)
            calls.append({**{x: e[x] for x in ("fn", "case", "sym", "policy", "dtype")}, "args": args, "out": arrays[f"c{k}_out"]})
        _cache["small"] = calls
    return _cache["small"]


def ewise_calls():
    """Recorded calls of add / sub / negate_blocks / dot_diag / apply_mask / embed_mask / trace (make_golden_ewise.py):
    dict(fn, case, sym, dtype, args{name: value}, out); ``mask`` arguments come back as {charge tuple: int64 index array},
    lists of arrays (the operands of ``add``) as tuples of arrays."""
    if "ewise" not in _cache:
        with gzip.open(os.path.join(GOLDEN, "calls_ewise.json.gz"), "rt") as f:
            index = json.load(f)
        arrays = np.load(os.path.join(GOLDEN, "calls_ewise.npz"))

        def get(val):
            if isinstance(val, str) and val.startswith("@"):
                return arrays[val[1:]]
            if isinstance(val, dict) and "__mask__" in val:
                return {tup(k): arrays[v[1:]] for k, v in val["__mask__"]}
            if isinstance(val, dict) and "__arrays__" in val:
                return tuple(arrays[v[1:]] for v in val["__arrays__"])
            return tup(val)
        _cache["ewise"] = [{**{x: e[x] for x in ("fn", "case", "sym", "dtype")}, "args": {n: get(v) for n, v in e["args"].items()},
                            "out": arrays[f"c{k}_out"]} for k, e in enumerate(index)]
    return _cache["ewise"]


def bench_structs():
    """Dict name -> structure fixture (operand block tables, recorded metas per policy, result structure)."""
    if "bench" not in _cache:
        with gzip.open(os.path.join(GOLDEN, "structs_bench.json.gz"), "rt") as f:
            raw = json.load(f)
        extra = os.path.join(GOLDEN, "structs_transpose.json.gz")      # T1 cases: merges that transpose large blocks
        if os.path.exists(extra):
            with gzip.open(extra, "rt") as f:
                raw.update(json.load(f))

        def conv(d):
            if isinstance(d, dict):
                return {k: conv(v) for k, v in d.items()}
            return tup(d)
        _cache["bench"] = {k: conv(v) for k, v in raw.items()}
    return _cache["bench"]


def chain_structs():
    """Dict name -> {"step1", "step2"}: two consecutive contractions grouping the shared tensor by different legs."""
    if "chain" not in _cache:
        with gzip.open(os.path.join(GOLDEN, "structs_chain.json.gz"), "rt") as f:
            raw = json.load(f)

        def conv(d):
            if isinstance(d, dict):
                return {k: conv(v) for k, v in d.items()}
            return tup(d)
        _cache["chain"] = {k: conv(v) for k, v in raw.items()}
    return _cache["chain"]


def ref_merge_fixtures():
    """The reference's own captured transpose_and_merge argument sets (experimental/main_d7chi98_1d.py) + output hashes."""
    if "refmerge" not in _cache:
        with gzip.open(os.path.join(GOLDEN, "ref_merge_fixtures.json.gz"), "rt") as f:
            raw = json.load(f)
        _cache["refmerge"] = {k: {x: tup(y) for x, y in v.items()} for k, v in raw.items()}
    return _cache["refmerge"]


def closed_form_input(n, dtype="float64"):
    """data[i] = ((i * 2654435761) mod 2^32) / 2^32 (imaginary part: the same sequence reversed), as in make_ref_fixtures.py."""
    i = np.arange(n, dtype=np.uint64)
    x = ((i * np.uint64(2654435761)) % np.uint64(2 ** 32)).astype(np.float64) / 2.0 ** 32
    return x + 1j * x[::-1] if dtype == "complex128" else x
