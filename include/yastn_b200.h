/*
 * yastn_b200 — C ABI of the B200-native block-sparse contraction kernels.
 *
 * This is the drop-in boundary for YASTN's contraction hot path.  Every entry point takes plain
 * pointers and sizes (no torch types); device pointers are raw CUDA device addresses, `stream` is a
 * cudaStream_t passed as void*.  All functions return 0 on success and a negative code on failure;
 * yb_last_error() returns a thread-local message for the last failure.  Calls are asynchronous on
 * `stream`, allocate nothing at run time (plans own their device tables) and are CUDA-graph capturable.
 *
 * Reference interfaces replaced (paths relative to the yastn tree, commit 60af786a):
 *   yb_copy_*  : backend.transpose_and_merge  yastn/backend/backend_torch.py:588 (loop: _backend_torch_backwards.py:340-364)
 *                backend.unmerge              yastn/backend/backend_torch.py:592 (loop: _backend_torch_backwards.py:397-408)
 *                backend.transpose            yastn/backend/backend_torch.py:584 (loop: _backend_torch_backwards.py:315-319)
 *                and their backward passes    _backend_torch_backwards.py:325-335, 377-392, 418-426
 *                prior-art C ABI of the same op: experimental/tm_worker.c:266-297 (tm_worker_parallel_float64/complex128)
 *   yb_gemm_*  : backend.dot                  yastn/backend/backend_torch.py:549 (loop: _backend_torch_backwards.py:100-109, backward 118-138)
 *                backend.transpose_dot_sum    yastn/backend/backend_torch.py:553 (loop: _backend_torch_backwards.py:143-157)
 *                prior art: experimental/torch_mmib.cpp:20-35 ("mm incommensurate batch")
 *   yb_tables_*: the meta pass between the reference's meta tuples and the plan tables (host code, yb_tables.cu)
 *   yb_chain_* : several tensordots in a row (Heff2, environment updates; yastn/tn/mps/_env.py:496-518) as one call
 */
#ifndef YASTN_B200_H
#define YASTN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YB_ABI_VERSION 4

/* element types */
#define YB_F64 0
#define YB_C128 1

/* src_base of a copy record that has no source: the destination box is filled with zeros (cells of a merged block that no
 * source block covers; replaces a memset of the whole destination) */
#define YB_COPY_SRC_ZERO INT64_MIN

/* yb_copy_run flags */
#define YB_COPY_ZERO_DST 1 /* clear dst before scattering (merge with missing blocks)               */
#define YB_COPY_CONJ 2     /* complex conjugate while copying (resolves torch's lazy conj bit)        */

/* yb_gemm_run flags */
#define YB_GEMM_CONJ_A 1
#define YB_GEMM_CONJ_B 2

typedef struct yb_copy_plan yb_copy_plan;
typedef struct yb_gemm_plan yb_gemm_plan;

int yb_abi_version(void);
const char* yb_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Block copy plans.  One record moves one N-d box:  for every index (i_0..i_{rank-1}) with
 * 0 <= i_k < extent[k]:   dst[dst_base + sum_k i_k*dst_stride[k]] = src[src_base + sum_k i_k*src_stride[k]].
 * recs is an nrec x (2 + 3*rank) row-major int64 table
 *     [src_base, dst_base, extent[0..rank), src_stride[0..rank), dst_stride[0..rank)]      (units: elements)
 * Records must write disjoint destinations.  itemsize is 8 (float64) or 16 (complex128).
 * ---------------------------------------------------------------------------------------------- */
int yb_copy_plan_create(const int64_t* recs, int64_t nrec, int rank, int itemsize, int device, yb_copy_plan** out);
/* info[0]=work items, [1]=elements moved (or zero-filled), [2]=records kept after normalisation, [3]=records on the
 * tiled-transpose path, [4]=entries of the run table (rows of small records and zero-fill boxes) */
int yb_copy_plan_info(const yb_copy_plan* plan, int64_t info[5]);
int yb_copy_run(const yb_copy_plan* plan, const void* src, void* dst, int64_t dst_elems, int flags, void* stream);
void yb_copy_plan_destroy(yb_copy_plan* plan);

/* ------------------------------------------------------------------------------------------------
 * Grouped block GEMM plans.  Problem p computes the M x N block
 *     C[offC + m*ldc + n] = sum over its segments s of  sum_k opA(A[offA_s + m*sAm_s + k*sAk_s]) * opB(B[offB_s + k*sBk_s + n*sBn_s])
 * problems: nprob x 6 int64 [M, N, offC, ldc, seg_begin, seg_end)
 * segments: nseg  x 7 int64 [K, offA, sAm, sAk, offB, sBk, sBn]           (units: elements)
 * In every segment one of (sAm, sAk) and one of (sBk, sBn) must be 1 (or the extent along it must be 1);
 * all segments of a plan share the same pair of unit-stride choices.
 * dtype is YB_F64 or YB_C128.  C blocks of different problems must be disjoint.
 * ---------------------------------------------------------------------------------------------- */
int yb_gemm_plan_create(const int64_t* problems, int64_t nprob, const int64_t* segments, int64_t nseg, int dtype,
                        int device, yb_gemm_plan** out);
/* Same, with a fused unmerge epilogue (backend.unmerge, _backend_torch_backwards.py:397-408, applied while the
 * block is still in registers): problem p with scat_index[p] = s >= 0 does not store its M x N block row-major at
 * offC; instead the block is cut by row_cuts[row_ptr[s] .. row_ptr[s+1]) (0 = c_0 < c_1 < .. < c_nrs = M) and
 * col_cuts[col_ptr[s] .. col_ptr[s+1]) and sub-block (i, j) is written contiguously (row-major, its own width) at
 * element offset dst[dst_ptr[s] + i*ncs + j] of C.  scat_index[p] = -1 keeps the plain store. */
int yb_gemm_plan_create_scatter(const int64_t* problems, int64_t nprob, const int64_t* segments, int64_t nseg,
                                const int64_t* scat_index, int64_t nscat, const int64_t* row_ptr, const int64_t* row_cuts,
                                const int64_t* col_ptr, const int64_t* col_cuts, const int64_t* dst_ptr, const int64_t* dst,
                                int dtype, int device, yb_gemm_plan** out);
/* info[0]=tiles, [1]=real multiply-adds (M*N*K summed), [2]=big tiles, [3]=small tiles, [4]=grid (CTAs), [5]=CTAs that
 * start inside a tile (stream-K partials), [6]=warps of the skinny-output kernel, [7]=its partial-sum runs, [8]=work units of
 * the panel kernel, [9]=reserved.
 * Problems whose result block is at most 8 x 8 (complex128: at most 32 entries) are not tiled: they run as HBM-bound
 * reductions over the contraction index in a second kernel of the same yb_gemm_run call (backend.vdot, huge-K / tiny-output
 * contractions, adjoints of tall-and-skinny products); any strides are accepted for them.  Problems with K <= 8 and exactly
 * one of M, N <= 8 (an MPO block applied to an environment: the other extent runs to 10^7) are streaming products and run in
 * a third kernel, also with any strides.
 * Cross-CTA scratch (stream-K partial tiles, partial sums) is looked up per (device, stream) at launch time: plans may run
 * concurrently on different streams.  Launches that wait on partial tiles of other CTAs are cooperative launches. */
int yb_gemm_plan_info(const yb_gemm_plan* plan, int64_t info[10]);
int yb_gemm_run(const yb_gemm_plan* plan, const void* A, const void* B, void* C, int flags, void* stream);
void yb_gemm_plan_destroy(yb_gemm_plan* plan);

/* ------------------------------------------------------------------------------------------------
 * Block-wise elementwise plans: the per-block loops of the reference's vector operations, one launch each.
 *   reference interfaces replaced: backend.add / sub        yastn/backend/backend_torch.py:518-534
 *                                  backend.negate_blocks    yastn/backend/_backend_torch_backwards.py:229-248
 *                                  backend.dot_diag         yastn/backend/backend_torch.py:557-564
 *                                  backend.apply_mask / embed_mask (and backwards)   _backend_torch_backwards.py:251-310
 *                                  backend.trace            yastn/backend/backend_torch.py:268-275
 * recs is an nrec x 16 row-major int64 table (units: elements)
 *     [0] mode  [1] dst  [2] n  [3..6] src offsets (YB_EW_ABSENT: no such source)  [7] negate mask  [8] aux  [9] post  [10] naxis  [11] nfull
 *   mode 0 LINCOMB  dst[d + i] = sum_k (+/-) src_k[s_k + i]                          i < n; bit k of [7] negates source k
 *   mode 1 DIAG     dst[d + e] = src_0[s_0 + e] * aux[a + (e / post) % naxis]          e < n; aux has the element type
 *   mode 2 GATHER   dst[d + e] = src_0[s_0 + (p * nfull + idx[a + j]) * post + q]      e = (p * naxis + j) * post + q < n; aux = int64 idx
 *   mode 3 SCATTER  dst[d + (p * nfull + idx[a + j]) * post + q] = src_0[s_0 + e]      same decomposition of e
 *   mode 4 TRACE    dst[d + e] = sum over rows c in [a, a + nfull) of `traces`, sum_{i < D_c} src_0[base_c + i * dstride_c + off_c(e)]
 *                   with off_c(e) = sum_k idx_k * stride_k for the row-major decomposition of e over ext_c[0..nd_c)
 * traces is an ntrace x 16 int64 table [0] base [1] D [2] dstride [3] nd [4..9] ext [10..15] stride.
 * Destination ranges of different records must be disjoint; dst may alias a source (same index in, same index out).
 * ---------------------------------------------------------------------------------------------- */
#define YB_EW_ABSENT INT64_MIN
#define YB_EW_LINCOMB 0
#define YB_EW_DIAG 1
#define YB_EW_GATHER 2
#define YB_EW_SCATTER 3
#define YB_EW_TRACE 4
typedef struct yb_ew_plan yb_ew_plan;
int yb_ew_plan_create(const int64_t* recs, int64_t nrec, const int64_t* traces, int64_t ntrace, int itemsize, int device,
                      yb_ew_plan** out);
/* info[0]=work pieces, [1]=elements of the iteration space */
int yb_ew_plan_info(const yb_ew_plan* plan, int64_t info[2]);
int yb_ew_run(const yb_ew_plan* plan, void* dst, const void* src0, const void* src1, const void* src2, const void* src3,
              const void* aux, void* stream);
void yb_ew_plan_destroy(yb_ew_plan* plan);

/* ------------------------------------------------------------------------------------------------
 * Batched SVD of small charge sectors (one-sided Jacobi in shared memory, one CTA per sector, one launch per block matrix).
 *   reference interface replaced: the per-sector loop of backend.svd / svdvals for sectors up to 64 x 64
 *   (yastn/backend/_backend_torch_backwards.py:26-39, yastn/backend/backend_torch.py:322-327; torch.linalg.svd per sector).
 * recs is an nrec x 6 int64 table [offA, m, n, offU, offS, offV] (element offsets): A row-major m x n inside `A`, U row-major
 * m x k inside `U`, k singular values (descending, float64) inside `S`, Vh row-major k x n inside `Vh`, k = min(m, n) <= 64.
 * status (nrec int32, device): 0 ok, 1 not converged within max_sweeps, 2 a singular value is exactly zero (the vectors of that
 * sector are not orthonormal; refactorise it with another routine).  vectors = 0: singular values only.
 * ---------------------------------------------------------------------------------------------- */
typedef struct yb_svd_plan yb_svd_plan;
int yb_svd_plan_create(const int64_t* recs, int64_t nrec, int itemsize, int device, yb_svd_plan** out);
int yb_svd_run(const yb_svd_plan* plan, const void* A, void* U, void* S, void* Vh, void* status, int max_sweeps, int vectors,
               void* stream);
void yb_svd_plan_destroy(yb_svd_plan* plan);

/* ------------------------------------------------------------------------------------------------
 * Host-side meta pass (no device calls): the reference's meta tuples, flattened depth-first to int64 tables, become the
 * record tables of the plans above.  A launch-bound run (DMRG at D=64 meets ~20 new block structures per bond) spends its
 * time here, not in the kernels.  Every call leaves its result in a thread-local buffer: read its length with
 * yb_tables_result_size() and copy it out with yb_tables_result_fetch().
 *   yb_tables_merge   : records of backend.transpose_and_merge (loop: _backend_torch_backwards.py:340-364) from
 *                       meta_mrg rows [tn(T), slo(2), Do(r), Dslc(2g), Drsh(g)] and meta_new rows [tn(T), Dn(g), sln(2)]
 *                       (yastn/tensor/_merging.py:137-187), plus zero-fill records for the cells no source block covers.
 *                       Result [status, rank, covered, nrec, recs...]; status 1: records not grouped in meta_new order (pass
 *                       grp_in), 2: zero-fill records not built (clear the destination instead).
 *   yb_tables_scatter : tables of yb_gemm_plan_create_scatter from meta_unmerge rows [sln(2), Dn(gn), slo(2), Do(2), r0, r1,
 *                       c0, c1] (yastn/tensor/_merging.py:528-549) and meta_dot rows (12 wide); shift: optional per-record
 *                       destination shift.  Result [nprob, ng, nrow, ncol, ndst, scat_index, row_ptr, row_cuts, col_ptr,
 *                       col_cuts, dst_ptr, dst].
 *   yb_tables_add     : LINCOMB records of backend.add / sub (yastn/backend/backend_torch.py:518-534) from rows
 *                       [operand, c0, c1, a0].  Result [nrounds, {nrec, nslots, slots[4], recs(nrec x 16)}...].
 * ---------------------------------------------------------------------------------------------- */
int64_t yb_tables_result_size(void);
int yb_tables_result_fetch(int64_t* out, int64_t n);
int yb_tables_merge(const int64_t* mrg, int64_t n, int64_t wm, const int64_t* neu, int64_t nnew, int64_t wn,
                    const int64_t* order, int r, int g, int T, const int64_t* grp_in, int zero_records);
int yb_tables_scatter(const int64_t* um, int64_t n, int gn, const int64_t* md, int64_t nprob, const int64_t* shift);
int yb_tables_add(const int64_t* ops, int64_t nrec, int64_t n_ops, const int64_t* signs);

/* ------------------------------------------------------------------------------------------------
 * Chains: a recorded sequence of copy / grouped-GEMM runs (several tensordots in a row) replayed by one call.
 *   reference call sites: the four tensordots of Env_mps_mpo_mps.Heff2 (yastn/tn/mps/_env.py:512-518), Heff1 (:506-510),
 *   update_env_to_last / _to_first (:496-504) — applied again and again to operands of unchanged block structure inside eigs.
 * steps is an nsteps x 10 int64 table [kind, run flags, plan handle, slotA, offA, slotB, offB, slotC, offC, dst_elems]:
 *   kind YB_CHAIN_COPY: yb_copy_run(plan, slots[slotA] + offA, slots[slotC] + offC, dst_elems, flags)
 *   kind YB_CHAIN_GEMM: yb_gemm_run(plan, slots[slotA] + offA, slots[slotB] + offB, slots[slotC] + offC, flags)
 * (offsets in bytes; slotB is ignored for copies).  slots are base device pointers supplied per run — the caller's operands,
 * a scratch arena holding the intermediates, the result.  The chain borrows the plans: they must outlive it.
 * ---------------------------------------------------------------------------------------------- */
#define YB_CHAIN_COPY 0
#define YB_CHAIN_GEMM 1
typedef struct yb_chain yb_chain;
int yb_chain_create(const int64_t* steps, int64_t nsteps, int64_t nslots, yb_chain** out);
int yb_chain_run(const yb_chain* chain, void* const* slots, int64_t nslots, void* stream);
int64_t yb_chain_steps(const yb_chain* chain);
void yb_chain_destroy(yb_chain* chain);

/* ------------------------------------------------------------------------------------------------
 * Peer arenas (multi-GPU, one process per GPU on one box).  The reference has no multi-GPU path for a
 * single contraction (its only multi-process code farms whole CTM environments out, yastn/tn/fpeps/envs/_env_ctm_dist_mp.py);
 * these calls give the block exchange between sharded contractions a direct NVLink data path: yb_peer_alloc creates a device
 * buffer and its 64-byte CUDA IPC handle, yb_peer_open maps another rank's buffer into this process.  A copy plan whose
 * records use  dst_base = (peer_ptr - local_ptr) / itemsize + offset  (int64 element offsets are unrestricted), or a GEMM
 * scatter table whose destinations are shifted the same way, then writes blocks straight into the peer's HBM.
 * ---------------------------------------------------------------------------------------------- */
int yb_peer_alloc(int64_t bytes, int device, void** ptr, unsigned char handle[64]);
int yb_peer_open(const unsigned char handle[64], int device, void** ptr);
int yb_peer_close(void* ptr);
int yb_peer_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* YASTN_B200_H */
