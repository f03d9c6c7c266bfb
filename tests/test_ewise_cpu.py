"""Host logic of the block-wise elementwise operations (no GPU): the int64 tables built by yastn_b200.plans, executed by
the numpy interpreter of the C-ABI records (tests/table_exec.py), reproduce every call recorded from the reference
(tests/golden/make_golden_ewise.py) — bit for bit where the operation only moves or negates data, to 1e-14 where it adds."""
import numpy as np
import pytest

from golden_io import ewise_calls
from table_exec import exec_ew
from yastn_b200 import plans

CALLS = ewise_calls()


def _sel(*fns):
    return [c for c in CALLS if c["fn"] in fns]


def _ids(calls):
    return [f"{c['fn']}-{c['case']}-{c['dtype']}-{k}" for k, c in enumerate(calls)]


def _close(out, ref, exact):
    assert out.shape == ref.shape
    if exact:
        assert np.array_equal(out, ref)
    else:
        assert np.linalg.norm(out - ref) <= 1e-14 * max(1.0, np.linalg.norm(ref))


@pytest.mark.parametrize("call", _sel("add", "sub"), ids=_ids(_sel("add", "sub")))
def test_add_sub_tables(call):
    a = call["args"]
    if call["fn"] == "add":
        datas, signs = list(a["datas"]), None
    else:
        datas, signs = [a["Adata"], a["Bdata"]], (1, -1)
    dt = np.result_type(*[d.dtype for d in datas])
    datas = [d.astype(dt) for d in datas]
    out = np.full(a["Dsize"], np.nan, dtype=dt)
    covered = sum(hi - lo for (lo, hi) in {sl for meta in a["metas"] for (sl, _) in meta})
    if covered < a["Dsize"]:
        out[:] = 0
    for recs, slots in plans.add_tables(a["metas"], signs):
        assert len(slots) <= 4
        exec_ew(recs, None, out, [out if k < 0 else datas[k] for k in slots] + [None] * (4 - len(slots)), None)
    _close(out, call["out"], exact=len(datas) <= 2)


@pytest.mark.parametrize("call", _sel("negate_blocks"), ids=_ids(_sel("negate_blocks")))
def test_negate_tables(call):
    a = call["args"]
    out = exec_ew(plans.negate_tables(a["slices"], a["Adata"].size), None, np.full_like(a["Adata"], np.nan), [a["Adata"], None, None, None], None)
    _close(out, call["out"], exact=True)


@pytest.mark.parametrize("call", _sel("dot_diag"), ids=_ids(_sel("dot_diag")))
def test_dot_diag_tables(call):
    a = call["args"]
    dt = np.result_type(a["Adata"].dtype, a["Bdata"].dtype)
    recs = plans.dot_diag_tables(a["meta"], a["axis"], a["a_ndim"])
    out = exec_ew(recs, None, np.full(a["Dsize"], np.nan, dtype=dt), [a["Bdata"].astype(dt), None, None, None], a["Adata"].astype(dt))
    _close(out, call["out"].astype(dt), exact=dt.kind != "c")      # complex products may round differently per operand order


@pytest.mark.parametrize("call", _sel("apply_mask", "embed_mask"), ids=_ids(_sel("apply_mask", "embed_mask")))
def test_mask_tables(call):
    a = call["args"]
    scatter = call["fn"] == "embed_mask"
    recs, order = plans.mask_tables(a["meta"], a["axis"], a["ndim"], scatter)
    idx = np.concatenate([np.asarray(a["mask"][tm], dtype=np.int64).reshape(-1) for tm in order]) if order else np.zeros(0, dtype=np.int64)
    out = (np.zeros if scatter else lambda n, dtype: np.full(n, np.nan, dtype=dtype))(a["Dsize"], dtype=a["Adata"].dtype)
    out = exec_ew(recs, None, out, [a["Adata"], None, None, None], idx)
    _close(out, call["out"], exact=True)
    # adjoint tables (the reference's backward, _backend_torch_backwards.py:267-279,297-310): gather <-> scatter
    adj = tuple((sla, Da, sln, Dn, tm) for sln, Dn, sla, Da, tm in a["meta"])
    recs, order2 = plans.mask_tables(adj, a["axis"], a["ndim"], not scatter)
    assert order2 == order
    back = (np.zeros if not scatter else lambda n, dtype: np.full(n, np.nan, dtype=dtype))(a["Adata"].size, dtype=a["Adata"].dtype)
    back = exec_ew(recs, None, back, [call["out"], None, None, None], idx)
    if scatter:      # embed then gather back: identity
        assert np.array_equal(back, a["Adata"])
    else:            # apply then scatter back: the selected entries return, the others are zero
        sel = back != 0
        assert np.array_equal(back[sel], a["Adata"][sel])


@pytest.mark.parametrize("call", _sel("trace"), ids=_ids(_sel("trace")))
def test_trace_tables(call):
    a = call["args"]
    recs, traces = plans.trace_tables(a["order"], a["meta"])
    covered = sum(sln[1] - sln[0] for sln, _ in a["meta"])
    out = np.zeros(a["Dsize"], dtype=a["data"].dtype) if covered < a["Dsize"] else np.full(a["Dsize"], np.nan, dtype=a["data"].dtype)
    out = exec_ew(recs, traces, out, [a["data"], None, None, None], None)
    _close(out, call["out"], exact=False)


def test_every_op_has_calls():
    for fn, least in (("add", 10), ("sub", 6), ("negate_blocks", 6), ("dot_diag", 8), ("apply_mask", 12), ("embed_mask", 4), ("trace", 8)):
        assert len(_sel(fn)) >= least, fn
