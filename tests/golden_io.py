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


def bench_structs():
    """Dict name -> structure fixture (operand block tables, recorded metas per policy, result structure)."""
    if "bench" not in _cache:
        with gzip.open(os.path.join(GOLDEN, "structs_bench.json.gz"), "rt") as f:
            raw = json.load(f)

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
