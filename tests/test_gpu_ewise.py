"""GPU parity of the block-wise elementwise engine (yb_ewise.cu) through the backend functions, against every call recorded
from the reference (tests/golden/make_golden_ewise.py): add / sub / negate_blocks / dot_diag / apply_mask / embed_mask / trace,
their backward passes where the reference defines one, and large synthetic cases against numpy."""
import numpy as np
import pytest
import torch

from golden_io import ewise_calls

pytestmark = pytest.mark.gpu
CALLS = ewise_calls()


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _ids(calls):
    return [f"{c['fn']}-{c['case']}-{c['dtype']}-{k}" for k, c in enumerate(calls)]


@pytest.fixture(scope="module")
def bk():
    from yastn_b200 import backend_b200
    return backend_b200


def _args(call):
    a = dict(call["args"])
    out = []
    for name, v in a.items():
        if isinstance(v, np.ndarray):
            out.append(_dev(v))
        elif name == "datas":
            out.append([_dev(x) for x in v])
        elif name == "mask":
            out.append({k: _dev(np.asarray(x, dtype=np.int64)) for k, x in v.items()})
        else:
            out.append(v)
    return out


@pytest.mark.parametrize("call", CALLS, ids=_ids(CALLS))
def test_golden_calls(bk, call):
    poison = torch.full((max(call["out"].size, 1),), float("nan"), dtype=torch.from_numpy(call["out"]).dtype, device="cuda")
    del poison       # outputs come from torch.empty: a NaN left in the recycled block would show an element nobody wrote
    out = getattr(bk, call["fn"])(*_args(call))
    torch.cuda.synchronize()
    got, ref = out.cpu().numpy(), call["out"]
    assert got.shape == ref.shape and got.dtype == ref.dtype
    if call["fn"] in ("negate_blocks", "apply_mask", "embed_mask") or (call["fn"] in ("sub",) and ref.dtype.kind != "c"):
        assert np.array_equal(got, ref)
    else:
        assert np.linalg.norm(got - ref) <= 1e-14 * max(1.0, np.linalg.norm(ref))


def test_mask_and_negate_backward_match_the_reference_adjoints(bk):
    n = 0
    for call in CALLS:
        if call["fn"] not in ("apply_mask", "embed_mask", "negate_blocks"):
            continue
        args = _args(call)
        x = args[0].clone().requires_grad_(True)
        out = getattr(bk, call["fn"])(x, *args[1:])
        G = torch.randn_like(out)
        out.backward(G)
        g = x.grad.cpu().numpy()
        Gn = G.cpu().numpy()
        a = call["args"]
        ref = np.zeros_like(g)
        if call["fn"] == "negate_blocks":
            ref = Gn.copy()
            for lo, hi in a["slices"]:
                ref[lo:hi] *= -1
        else:
            ax, nd = a["axis"], a["ndim"]
            for sln, Dn, sla, Da, tm in a["meta"]:
                sel = (slice(None),) * ax + (np.asarray(a["mask"][tm]),) + (slice(None),) * max(nd - ax - 1, 0)
                Dn_, Da_ = (Dn if isinstance(Dn, tuple) else (Dn,)), (Da if isinstance(Da, tuple) else (Da,))
                if call["fn"] == "apply_mask":       # _backend_torch_backwards.py:267-279
                    ref[sla[0]:sla[1]].reshape(Da_)[sel] = Gn[sln[0]:sln[1]].reshape(Dn_)
                else:                                # :297-310
                    ref[sla[0]:sla[1]].reshape(Da_)[...] = Gn[sln[0]:sln[1]].reshape(Dn_)[sel]
        assert np.array_equal(g, ref), (call["fn"], call["case"])
        n += 1
    assert n >= 20


@pytest.mark.parametrize("cplx", [False, True], ids=["f64", "c128"])
def test_large_lincomb_and_negate(bk, cplx):
    """Krylov-sized vectors: 5 operands of 2e7 elements with different block subsets, and a sign flip of every third block."""
    rng = np.random.default_rng(3)
    sizes = rng.integers(1, 200_000, 200)
    bounds = np.concatenate(([0], np.cumsum(sizes)))
    N = int(bounds[-1])
    blocks = [(int(bounds[i]), int(bounds[i + 1])) for i in range(len(sizes))]
    datas, metas = [], []
    for k in range(5):
        keep = [b for i, b in enumerate(blocks) if (i + k) % 4 != 0]        # every operand misses a quarter of the blocks
        n = sum(hi - lo for lo, hi in keep)
        x = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
        pos, meta = 0, []
        for lo, hi in keep:
            meta.append(((lo, hi), (pos, pos + hi - lo)))
            pos += hi - lo
        datas.append(x)
        metas.append(tuple(meta))
    metas = tuple(metas)
    ref = np.zeros(N, dtype=datas[0].dtype)
    for x, meta in zip(datas, metas):
        for (c0, c1), (a0, a1) in meta:
            ref[c0:c1] += x[a0:a1]
    out = bk.add([_dev(x) for x in datas], metas, N)
    assert np.linalg.norm(out.cpu().numpy() - ref) <= 1e-14 * np.linalg.norm(ref)
    sl = tuple(blocks[::3])
    neg = bk.negate_blocks(_dev(ref), sl)
    expect = ref.copy()
    for lo, hi in sl:
        expect[lo:hi] *= -1
    assert np.array_equal(neg.cpu().numpy(), expect)
