"""Device-side sector matching: build the `dot` tables of a contraction on the GPU from raw block tables.

Mirrors what the reference computes on the host in ``_meta_tensordot_f2m`` / ``_meta_tensordot_fc``
(yastn/tensor/_contractions.py:281-346): which (A block, B block) pairs share the contracted charge, the shape and offset
of every result block, the total result size.  The kernel (csrc/yb_match.cu) emits the int64 problem / segment tables of
``yb_gemm_plan_create`` directly; ``meta_dot_from_tables`` converts them back to the reference's tuple format for parity.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def block_tables_f2m(blocks_a, blocks_b, nsym):
    """Block tables of two operands merged to matrices.  blocks = ((t, D, (lo, hi)), ...) with t = row + column charge."""
    a_key = np.array([t[nsym:] for t, _, _ in blocks_a], dtype=np.int64).reshape(len(blocks_a), nsym)
    b_key = np.array([t[:nsym] for t, _, _ in blocks_b], dtype=np.int64).reshape(len(blocks_b), nsym)
    a_dims = np.array([D for _, D, _ in blocks_a], dtype=np.int64).reshape(len(blocks_a), 2)
    b_dims = np.array([D for _, D, _ in blocks_b], dtype=np.int64).reshape(len(blocks_b), 2)
    a_off = np.array([sl[0] for _, _, sl in blocks_a], dtype=np.int64)
    b_off = np.array([sl[0] for _, _, sl in blocks_b], dtype=np.int64)
    return a_key, a_dims, a_off, b_key, b_dims, b_off


def block_tables_fc(blocks_a, blocks_b, nsym):
    """Block tables of operands with only the contracted legs fused (last leg of A, first leg of B)."""
    a_key = np.array([t[len(t) - nsym:] for t, _, _ in blocks_a], dtype=np.int64).reshape(len(blocks_a), nsym)
    b_key = np.array([t[:nsym] for t, _, _ in blocks_b], dtype=np.int64).reshape(len(blocks_b), nsym)
    a_dims = np.array([(int(np.prod(D[:-1], dtype=np.int64)), D[-1]) for _, D, _ in blocks_a], dtype=np.int64).reshape(len(blocks_a), 2)
    b_dims = np.array([(D[0], int(np.prod(D[1:], dtype=np.int64))) for _, D, _ in blocks_b], dtype=np.int64).reshape(len(blocks_b), 2)
    a_off = np.array([sl[0] for _, _, sl in blocks_a], dtype=np.int64)
    b_off = np.array([sl[0] for _, _, sl in blocks_b], dtype=np.int64)
    return a_key, a_dims, a_off, b_key, b_dims, b_off


def match_sectors(a_key, a_dims, a_off, b_key, b_dims, b_off, device, capacity=None):
    """Run the device join.  Returns (problems[n, 6], segments[n, 7], c_size) as numpy int64 arrays."""
    lib = _lib.load()
    dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    na, nb = int(a_dims.shape[0]), int(b_dims.shape[0])
    kw = int(a_key.shape[1]) if a_key.ndim == 2 else 0
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.int64)).to(dev)
    da = [t(a_key), t(a_dims), t(a_off)]
    db = [t(b_key), t(b_dims), t(b_off)]
    if capacity is None:
        capacity = max(na, 1) * 4
    while True:
        problems = torch.empty((capacity, 6), dtype=torch.int64, device=dev)
        segments = torch.empty((capacity, 7), dtype=torch.int64, device=dev)
        result = torch.zeros(3, dtype=torch.int64, device=dev)
        scratch = torch.empty(int(lib.yb_match_scratch_elems(na, nb)), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            rc = lib.yb_match_sectors(da[0].data_ptr(), da[1].data_ptr(), da[2].data_ptr(), na, db[0].data_ptr(), db[1].data_ptr(),
                                      db[2].data_ptr(), nb, kw, capacity, problems.data_ptr(), segments.data_ptr(), result.data_ptr(),
                                      scratch.data_ptr(), stream)
        if rc:
            _lib.check(rc)
        npairs, csize, status = (int(x) for x in result.cpu())
        if status == 1:
            raise ValueError("Bond dimensions do not match.")
        if status == 2:
            capacity = npairs
            continue
        return problems[:npairs].cpu().numpy(), segments[:npairs].cpu().numpy(), csize


def meta_dot_from_tables(problems, segments):
    """The reference's meta_dot tuple ((slc, Dc, sla, Da, slb, Db), ...) from the device tables."""
    out = []
    for (M, N, offC, ldc, s0, s1), (K, offA, lda, _, offB, ldb, _) in zip(problems.tolist(), segments.tolist()):
        out.append(((offC, offC + M * N), (M, N), (offA, offA + M * K), (M, K), (offB, offB + K * N), (K, N)))
    return tuple(out)
