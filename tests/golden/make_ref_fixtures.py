#!/usr/bin/env python
"""Fixtures from the reference's own captured merge metas (authoring container only: reads /root/reference).

    python tests/golden/make_ref_fixtures.py     ->  tests/golden/ref_merge_fixtures.json.gz

``experimental/main_d7chi98_1d.py:18-61`` holds three ``transpose_and_merge`` argument sets captured from real runs
(SURVEY.md 8c): ``a`` — a D=7, chi=98 CTMRG merge (rank-6 source, 281 blocks, order (0,2,4,5,1,3), 922 753 -> 954 145
elements incl. zero padding, five very skinny target matrices), ``b`` — a small rank-5 merge, ``c`` — a rank-5 source fused
into a 3-D target (fuse_legs).  The script's assignments are evaluated as they stand (nothing is copied into this repo), the
reference's own numpy implementation of the op (``experimental/backend_np_1d.py:32-40``) is run on a closed-form input
(``data[i] = ((i * 2654435761) mod 2^32) / 2^32``, reproducible anywhere without a random generator), and the fixture stores
the metas plus the SHA-256 of the expected output bytes — the op is a pure permutation with zero padding, so parity is
bit-exact equality of the hash.
"""
import ast
import gzip
import hashlib
import importlib.util
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("YASTN_REF", "/root/reference")


def closed_form_input(n):
    i = np.arange(n, dtype=np.uint64)
    return ((i * np.uint64(2654435761)) % np.uint64(2 ** 32)).astype(np.float64) / 2.0 ** 32


def main():
    src = open(os.path.join(REF, "experimental", "main_d7chi98_1d.py")).read()
    tree = ast.parse(src)
    env = {}
    for node in tree.body:                              # top-level constant assignments only
        if isinstance(node, ast.Assign):
            exec(compile(ast.Module([node], []), "main_d7chi98_1d", "exec"), env)
    spec = importlib.util.spec_from_file_location("ref_backend_np_1d", os.path.join(REF, "experimental", "backend_np_1d.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    for k in ("a", "b", "c"):
        n, order, meta_new, meta_mrg, Dsize = (env[f"{k}_{x}"] for x in ("data_size", "order", "meta_new", "meta_mrg", "Dsize"))
        case = {"data_size": n, "order": order, "meta_new": meta_new, "meta_mrg": meta_mrg, "Dsize": Dsize}
        for dtype in ("float64", "complex128"):
            x = closed_form_input(n)
            if dtype == "complex128":
                x = x + 1j * closed_form_input(n)[::-1]
            y = ref.transpose_and_merge(x, order, meta_new, meta_mrg, Dsize)
            case["sha256_" + dtype] = hashlib.sha256(np.ascontiguousarray(y).tobytes()).hexdigest()
            case["nonzero_" + dtype] = int(np.count_nonzero(y))
        out[k] = case
        print(k, "blocks", len(meta_mrg), "rank", len(order), n, "->", Dsize, case["sha256_float64"][:16])
    with gzip.open(os.path.join(HERE, "ref_merge_fixtures.json.gz"), "wt") as f:
        json.dump(out, f, separators=(",", ":"))
    print("bytes", os.path.getsize(os.path.join(HERE, "ref_merge_fixtures.json.gz")))


if __name__ == "__main__":
    main()
