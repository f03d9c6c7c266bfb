"""CPU restatement of the reference's sector pairing (TEST INFRASTRUCTURE — not product code).

Follows yastn/tensor/_contractions.py:281-298 (``_meta_tensordot_f2m``) and :301-346 (``_meta_tensordot_fc``) on plain block
tables ((t, D, (lo, hi)), ...): pair blocks on the contracted charge, sort result blocks by output charge, assign result
slices by accumulating block sizes.  Pinned by tests/test_oracle_golden.py against the ``meta_dot`` tuples recorded from the
reference (tests/golden/structs_bench.json.gz).  Only tests may import this module.
"""
from itertools import accumulate, groupby, product
from operator import itemgetter


def meta_dot_f2m(blocks_a, blocks_b, nsym):
    """blocks of 2-leg (merged) operands; returns (meta_dot, t_c, D_c, size)."""
    a_sorted = sorted((t[nsym:], t, D, sl) for t, D, sl in blocks_a)
    meta = []
    for (tar, ta, Da, sla), (tb, Db, slb) in zip(a_sorted, blocks_b):
        assert tar == tb[:nsym]
        meta.append((ta[:nsym] + tb[nsym:], (Da[0], Db[1]), sla, Da, slb, Db))
    return _finish(sorted(meta), 1)


def meta_dot_fc(blocks_a, blocks_b, nsym):
    """blocks with only the contracted legs fused: last leg of A / first leg of B; returns (meta_dot, t_c, D_c, size)."""
    def prod(x):
        p = 1
        for v in x:
            p *= v
        return p
    ra = sorted((t[len(t) - nsym:], t[:len(t) - nsym], D[-1], prod(D[:-1]), D[:-1], sl) for t, D, sl in blocks_a)
    rb = [(t[:nsym], t[nsym:], D[0], prod(D[1:]), D[1:], sl) for t, D, sl in blocks_b]
    meta = []
    for (tar, ga), (tbl, gb) in zip(groupby(ra, key=itemgetter(0)), groupby(rb, key=itemgetter(0))):
        assert tar == tbl
        for (_, toa, Dca, Dopa, Doa, sla), (_, tob, Dcb, Dopb, Dob, slb) in product(list(ga), list(gb)):
            meta.append((toa + tob, Doa + Dob, Dopa * Dopb, (Dopa, Dopb), sla, (Dopa, Dca), slb, (Dcb, Dopb)))
    return _finish(sorted(meta), 3)


def _finish(meta, first):
    t_c = tuple(x[0] for x in meta)
    D_c = tuple(x[1] for x in meta)
    Dp = tuple(x[2] for x in meta) if first == 3 else tuple(D[0] * D[1] for D in D_c)
    slices = tuple((stop - dp, stop) for stop, dp in zip(accumulate(Dp), Dp))
    meta_dot = tuple((sl, *m[first:]) for sl, m in zip(slices, meta))
    return meta_dot, t_c, D_c, sum(Dp)
