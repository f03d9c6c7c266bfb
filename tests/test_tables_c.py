"""Host logic (no GPU): the meta pass in C (csrc/yb_tables.cu) produces the same tables, entry for entry, as its numpy
statement in yastn_b200.plans (which the other CPU tests check against the oracle) on every recorded meta."""
import numpy as np
import pytest

from golden_io import small_calls, ewise_calls, bench_structs, ref_merge_fixtures
from yastn_b200 import plans

CALLS = small_calls()


def _same(a, b):
    if isinstance(a, (tuple, list)):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            _same(x, y)
    elif isinstance(a, np.ndarray):
        assert a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b)
    else:
        assert a == b


def test_merge_tables_match_numpy_on_golden_calls():
    n = 0
    for c in CALLS:
        if c["fn"] != "transpose_and_merge":
            continue
        a = c["args"]
        for zero in (True, False):
            _same(plans.merge_records(a["order"], a["meta_new"], a["meta_mrg"], zero),
                  plans.merge_records_np(a["order"], a["meta_new"], a["meta_mrg"], zero))
        n += 1
    assert n > 50


def test_merge_tables_match_numpy_on_bench_and_reference_fixtures():
    for name, st in bench_structs().items():
        if "D16384" in name or "D8192" in name:
            continue                                  # same code path, seconds of numpy each
        for pol in ("f2m", "fc"):
            for side in ("merge_a", "merge_b"):
                m = st.get(pol, {}).get(side)
                if m is None:
                    continue
                _same(plans.merge_records(m["order"], m["meta_new"], m["meta_mrg"]), plans.merge_records_np(m["order"], m["meta_new"], m["meta_mrg"]))
    for name, fx in ref_merge_fixtures().items():
        _same(plans.merge_records(fx["order"], fx["meta_new"], fx["meta_mrg"]), plans.merge_records_np(fx["order"], fx["meta_new"], fx["meta_mrg"]))


def test_merge_tables_holes_and_group_fallback():
    # two merged blocks; the first has a hole in the middle of a 3 x 3 grid, the second no source at all
    meta_new = (((0,), (6, 9), (0, 54)), ((1,), (4, 5), (54, 74)))
    cells = [(r, c) for r in range(3) for c in range(3) if (r, c) != (1, 1)]
    meta_mrg, off = [], 100
    for r, c in cells:
        meta_mrg.append(((0,), (off, off + 6), (2, 3), ((2 * r, 2 * r + 2), (3 * c, 3 * c + 3)), (2, 3)))
        off += 6
    meta_mrg = tuple(meta_mrg)
    x, y = plans.merge_records((0, 1), meta_new, meta_mrg), plans.merge_records_np((0, 1), meta_new, meta_mrg)
    _same(x, y)
    assert x[2] == 74 and (x[0][:, 0] == plans.SRC_ZERO).sum() == 2
    # records handed over in another order than meta_new: the dictionary fallback
    shuffled = (meta_new[1], meta_new[0])
    _same(plans.merge_records((0, 1), shuffled, meta_mrg), plans.merge_records_np((0, 1), shuffled, meta_mrg))
    with pytest.raises(ValueError):
        plans.merge_records((0, 1), meta_new, (((0,), (0, 6), (2, 3), ((0, 2), (0, 3)), (3, 2)),))     # reshape groups do not align


def test_scatter_tables_match_numpy():
    n = 0
    for name, st in bench_structs().items():
        m = st.get("f2m")
        if not m or m.get("unmerge") is None:
            continue
        md, mu = m["dot"]["meta_dot"], m["unmerge"]["meta"]
        _same(plans.unmerge_scatter_tables(md, mu), plans.unmerge_scatter_tables_np(md, mu))
        shift = np.arange(len(mu), dtype=np.int64) * 1000
        _same(plans.unmerge_scatter_tables(md, mu, shift), plans.unmerge_scatter_tables_np(md, mu, shift))
        n += 1
    assert n >= 5
    with pytest.raises(ValueError):          # rectangles that do not tile the block
        plans.unmerge_scatter_tables((((0, 12), (3, 4), (0, 6), (3, 2), (0, 8), (2, 4)),), (((0, 6), (3, 2), (0, 12), (3, 4), ((0, 3), (0, 2))),))


def test_add_tables_match_python():
    n = 0
    for c in ewise_calls():
        if c["fn"] not in ("add", "sub"):
            continue
        metas = c["args"]["metas"]
        signs = (1, -1) if c["fn"] == "sub" else None
        _same(plans.add_tables(metas, signs), plans.add_tables_np(metas, signs))
        n += 1
    assert n >= 4
    # six operands: a second round that reads the running sum; an operand that skips a block; an empty interval
    metas = tuple(tuple(((lo, lo + 5), (k + lo, k + lo + 5)) for lo in (0, 5, 20) if (k + lo) % 3) + (((7, 7), (0, 0)),) for k in range(6))
    _same(plans.add_tables(metas), plans.add_tables_np(metas))
    with pytest.raises(ValueError):
        plans.add_tables(((((0, 4), (0, 4)), ((2, 6), (4, 8))),))
