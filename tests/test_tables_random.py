"""Randomised host-logic check (no GPU): on block structures drawn from the real YASTN — random symmetric tensors, random
splits of their legs into (row, column) groups, random subsets of blocks taking part (which leaves holes in the merged blocks) —
the meta pass in C builds the same tables as its numpy statement, and executing the merge tables gives the reference result
of the numpy oracle (zero-filled holes included)."""
import itertools

import numpy as np
import pytest

from yastn_loader import load_yastn

yastn = load_yastn()
if yastn is None:
    pytest.skip("yastn not importable (no baseline/_ref, no reference checkout)", allow_module_level=True)

from oracle import backend_oracle as orc  # noqa: E402
from table_exec import exec_copy  # noqa: E402
from yastn_b200 import plans  # noqa: E402
import yastn.tensor._merging as M  # noqa: E402
import yastn.tensor._contractions as C  # noqa: E402


def _same(a, b):
    if isinstance(a, (tuple, list)):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            _same(x, y)
    elif isinstance(a, np.ndarray):
        assert a.shape == b.shape and np.array_equal(a, b)
    else:
        assert a == b


def _random_tensor(rng, sym):
    cfg = yastn.make_config(sym=sym, backend="np")
    ndim = int(rng.integers(3, 6))
    legs = []
    for _ in range(ndim):
        if sym == "U1":
            t = sorted(rng.choice(np.arange(-2, 3), size=int(rng.integers(2, 5)), replace=False).tolist())
        elif sym == "Z2":
            t = sorted(rng.choice([0, 1], size=int(rng.integers(1, 3)), replace=False).tolist())
        else:
            pool = list(itertools.product((-1, 0, 1), (0, 1)))
            t = sorted(pool[i] for i in rng.choice(len(pool), size=int(rng.integers(1, 4)), replace=False))
        legs.append(yastn.Leg(cfg, s=int(rng.choice([-1, 1])), t=t, D=[int(d) for d in rng.integers(1, 5, size=len(t))]))
    nsym = cfg.sym.NSYM
    n = tuple(int(x) for x in (rng.integers(0, 2, size=nsym) if sym != "U1" else rng.integers(-1, 2, size=1)))
    return yastn.rand(config=cfg, legs=legs, n=n if nsym > 1 else n[0])


@pytest.mark.parametrize("sym", ["U1", "Z2", "U1xU1"])
def test_merge_tables_on_random_structures(sym):
    rng = np.random.default_rng({"U1": 11, "Z2": 12, "U1xU1": 13}[sym])
    done = holes = 0
    while done < 40:
        a = _random_tensor(rng, sym)
        if a.size == 0 or len(a.struct.t) == 0:
            continue
        perm = rng.permutation(a.ndim_n).tolist()
        cut = int(rng.integers(0, a.ndim_n + 1)) if done % 4 == 0 else int(rng.integers(1, a.ndim_n))
        axes = (tuple(perm[:cut]), tuple(perm[cut:]))
        nb = len(a.struct.t)
        # mostly a few blocks dropped: the merged blocks keep their shape and get uncovered cells
        keep = nb if rng.random() < 0.3 else max(1, nb - max(1, nb // 4))
        inds = None if keep == nb else tuple(sorted(rng.choice(nb, size=keep, replace=False).tolist()))
        struct, slices, meta_mrg, _, _ = M._meta_merge_to_matrix(a.config, a.struct, a.slices, axes, inds)
        meta_new = tuple((x, y, z.slcs[0]) for x, y, z in zip(struct.t, struct.D, slices))
        order = axes[0] + axes[1]
        got = plans.merge_records(order, meta_new, meta_mrg)
        _same(got, plans.merge_records_np(order, meta_new, meta_mrg))
        recs, rank, covered = got
        data = rng.standard_normal(a.size)
        ref = orc.transpose_and_merge(data, order, meta_new, meta_mrg, struct.size)
        dst = np.full(struct.size, np.nan)
        if covered < struct.size:          # zero-fill records were not built: the caller clears the destination
            dst[:] = 0
        out = exec_copy(recs, rank, data, dst)
        assert np.array_equal(out, ref)
        holes += int((recs[:, 0] == plans.SRC_ZERO).any()) if len(recs) else 0
        done += 1
    assert holes >= (3 if sym == "U1" else 0)     # the U1 draw really produces merged blocks with uncovered cells (the others rarely do)


@pytest.mark.parametrize("sym", ["U1", "Z2", "U1xU1"])
def test_scatter_tables_on_random_contractions(sym):
    rng = np.random.default_rng({"U1": 21, "Z2": 22, "U1xU1": 23}[sym])
    done = 0
    tries = 0
    while done < 25 and tries < 400:
        tries += 1
        a = _random_tensor(rng, sym)
        if a.size == 0:
            continue
        k = int(rng.integers(1, a.ndim_n))
        perm = rng.permutation(a.ndim_n).tolist()
        ain, aout = tuple(perm[:k]), tuple(perm[k:])
        b = a.conj()                                  # contracting with the conjugate always matches charges and dimensions
        try:
            nout_a, nin_a, nin_b, nout_b = aout, ain, ain, aout
            ind_a, ind_b = C._common_inds(a.struct.t, b.struct.t, nin_a, nin_b, a.ndim_n, b.ndim_n, a.config.sym.NSYM)
            sa, sla, _, ls_l, ls_ac = M._meta_merge_to_matrix(a.config, a.struct, a.slices, (nout_a, nin_a), ind_a)
            sb, slb, _, ls_bc, ls_r = M._meta_merge_to_matrix(b.config, b.struct, b.slices, (nin_b, nout_b), ind_b)
            meta_dot, struct_m, slices_m = C._meta_tensordot_f2m(sa, sla, sb, slb)
            s_c = tuple(a.struct.s[i] for i in nout_a) + tuple(b.struct.s[i] for i in nout_b)
            meta_unmerge, _, _ = M._meta_unmerge_matrix(a.config, struct_m, slices_m, ls_l, ls_r, s_c)
        except Exception:
            continue
        if not meta_unmerge or not meta_dot:
            continue
        _same(plans.unmerge_scatter_tables(meta_dot, meta_unmerge), plans.unmerge_scatter_tables_np(meta_dot, meta_unmerge))
        shift = rng.integers(0, 1000, size=len(meta_unmerge)).astype(np.int64)
        _same(plans.unmerge_scatter_tables(meta_dot, meta_unmerge, shift), plans.unmerge_scatter_tables_np(meta_dot, meta_unmerge, shift))
        done += 1
    assert done >= 10
