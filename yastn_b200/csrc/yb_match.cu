// Device-side sector matching: the meta pass of a block-sparse contraction without a host loop over blocks.
//
// Reproduces the pairing and output layout of yastn/tensor/_contractions.py:281-298 (_meta_tensordot_f2m: one B block per
// A block) and :301-346 (_meta_tensordot_fc: every (A block, B block) pair that shares the contracted charge), i.e. the
// `meta_dot` table handed to backend.dot, directly in the int64 problem / segment format of yb_gemm_plan_create.
//
// Inputs are block tables already in the reference's block order: A blocks sorted by charge tuple (outgoing legs first,
// contracted charge last), B blocks sorted by charge tuple (contracted charge first).  Because the total charge is fixed,
// the outgoing charges of A are unique per block, so the reference's final sort by output charge (t_out_a + t_out_b) is
// the order "A blocks in table order, and for each its matching B range in table order": the join is a binary search per
// A block plus two exclusive scans, all on the device (one CTA: this is O(blocks), microseconds).
#include "yb_common.h"

namespace yb {

constexpr int kMatchThreads = 1024;

struct MatchArgs {
    const int64_t* a_key;    // [na, kw] contracted charge of every A block
    const int64_t* a_dims;   // [na, 2]  (M, K)
    const int64_t* a_off;    // [na]     element offset of the block
    const int64_t* b_key;    // [nb, kw] sorted lexicographically
    const int64_t* b_dims;   // [nb, 2]  (K, N)
    const int64_t* b_off;    // [nb]
    int64_t na, nb;
    int kw;
    int64_t capacity;        // rows available in problems / segments
    int64_t* problems;       // [capacity, 6]
    int64_t* segments;       // [capacity, 7]
    int64_t* result;         // [0] pairs, [1] C size (elements), [2] status: 0 ok, 1 contracted dims differ, 2 capacity exceeded
    int64_t* scratch;        // [3 * na + nb + 1]
};

__device__ __forceinline__ int key_cmp(const int64_t* x, const int64_t* y, int kw) {
    for (int k = 0; k < kw; ++k) {
        if (x[k] < y[k]) return -1;
        if (x[k] > y[k]) return 1;
    }
    return 0;
}

// Exclusive scan of v[0..n) in place (single CTA); returns the total to every thread.
__device__ int64_t block_exclusive_scan(int64_t* v, int64_t n, int64_t* sh) {
    const int tid = threadIdx.x;
    const int64_t chunk = (n + kMatchThreads - 1) / kMatchThreads;
    const int64_t lo = min(n, tid * chunk), hi = min(n, lo + chunk);
    int64_t s = 0;
    for (int64_t i = lo; i < hi; ++i) s += v[i];
    sh[tid] = s;
    __syncthreads();
    if (tid == 0) {
        int64_t run = 0;
        for (int t = 0; t < kMatchThreads; ++t) {
            const int64_t x = sh[t];
            sh[t] = run;
            run += x;
        }
        sh[kMatchThreads] = run;
    }
    __syncthreads();
    int64_t run = sh[tid];
    for (int64_t i = lo; i < hi; ++i) {
        const int64_t x = v[i];
        v[i] = run;
        run += x;
    }
    const int64_t total = sh[kMatchThreads];
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(kMatchThreads) match_kernel(const MatchArgs g) {
    __shared__ int64_t sh[kMatchThreads + 1];
    __shared__ int status;
    const int tid = threadIdx.x;
    int64_t* lb = g.scratch;
    int64_t* cnt = g.scratch + g.na;
    int64_t* base = g.scratch + 2 * g.na;
    int64_t* pb = g.scratch + 3 * g.na;   // [nb + 1] prefix of N over B blocks
    if (tid == 0) status = 0;
    for (int64_t j = tid; j <= g.nb; j += kMatchThreads) pb[j] = j < g.nb ? g.b_dims[2 * j + 1] : 0;
    __syncthreads();
    block_exclusive_scan(pb, g.nb + 1, sh);

    for (int64_t i = tid; i < g.na; i += kMatchThreads) {
        const int64_t* key = g.a_key + i * g.kw;
        int64_t lo = 0, hi = g.nb;
        while (lo < hi) {  // lower bound
            const int64_t mid = (lo + hi) >> 1;
            if (key_cmp(g.b_key + mid * g.kw, key, g.kw) < 0) lo = mid + 1; else hi = mid;
        }
        const int64_t first = lo;
        hi = g.nb;
        while (lo < hi) {  // upper bound
            const int64_t mid = (lo + hi) >> 1;
            if (key_cmp(g.b_key + mid * g.kw, key, g.kw) <= 0) lo = mid + 1; else hi = mid;
        }
        lb[i] = first;
        cnt[i] = lo - first;
        const int64_t K = g.a_dims[2 * i + 1];
        for (int64_t j = first; j < lo; ++j)
            if (g.b_dims[2 * j] != K) status = 1;   // 'Bond dimensions do not match.' (_contractions.py:148-149)
        base[i] = g.a_dims[2 * i] * (pb[lo] - pb[first]);
    }
    __syncthreads();
    // cnt -> first pair index of every A block; base -> first C element of every A block
    // (cnt is needed again below: keep the counts implicit as differences of the scanned array and the total)
    const int64_t npairs = block_exclusive_scan(cnt, g.na, sh);
    const int64_t csize = block_exclusive_scan(base, g.na, sh);
    if (tid == 0) {
        g.result[0] = npairs;
        g.result[1] = csize;
        g.result[2] = status ? 1 : (npairs > g.capacity ? 2 : 0);
    }
    if (npairs > g.capacity) return;
    for (int64_t i = tid; i < g.na; i += kMatchThreads) {
        const int64_t p0 = cnt[i], p1 = (i + 1 < g.na) ? cnt[i + 1] : npairs;
        const int64_t M = g.a_dims[2 * i], K = g.a_dims[2 * i + 1], first = lb[i];
        for (int64_t p = p0; p < p1; ++p) {
            const int64_t j = first + (p - p0);
            const int64_t N = g.b_dims[2 * j + 1];
            int64_t* pr = g.problems + p * 6;
            pr[0] = M;
            pr[1] = N;
            pr[2] = base[i] + M * (pb[j] - pb[first]);
            pr[3] = N;
            pr[4] = p;
            pr[5] = p + 1;
            int64_t* sg = g.segments + p * 7;
            sg[0] = K;
            sg[1] = g.a_off[i];
            sg[2] = K;
            sg[3] = 1;
            sg[4] = g.b_off[j];
            sg[5] = N;
            sg[6] = 1;
        }
    }
}

}  // namespace yb

using namespace yb;

extern "C" int64_t yb_match_scratch_elems(int64_t na, int64_t nb) { return 3 * na + nb + 1; }

extern "C" int yb_match_sectors(const int64_t* a_key, const int64_t* a_dims, const int64_t* a_off, int64_t na,
                                const int64_t* b_key, const int64_t* b_dims, const int64_t* b_off, int64_t nb, int key_width,
                                int64_t capacity, int64_t* problems, int64_t* segments, int64_t* result, int64_t* scratch,
                                void* stream) {
    if (na < 0 || nb < 0 || key_width < 0 || capacity < 0) return fail(kErrArg, "yb_match_sectors: negative size");
    if (!result || !scratch) return fail(kErrArg, "yb_match_sectors: null result / scratch");
    if ((na > 0 && (!a_dims || !a_off || (key_width > 0 && !a_key))) || (nb > 0 && (!b_dims || !b_off || (key_width > 0 && !b_key))))
        return fail(kErrArg, "yb_match_sectors: null block table");
    if (capacity > 0 && (!problems || !segments)) return fail(kErrArg, "yb_match_sectors: null output table");
    MatchArgs g;
    g.a_key = a_key;
    g.a_dims = a_dims;
    g.a_off = a_off;
    g.b_key = b_key;
    g.b_dims = b_dims;
    g.b_off = b_off;
    g.na = na;
    g.nb = nb;
    g.kw = key_width;
    g.capacity = capacity;
    g.problems = problems;
    g.segments = segments;
    g.result = result;
    g.scratch = scratch;
    match_kernel<<<1, kMatchThreads, 0, (cudaStream_t)stream>>>(g);
    YB_CUDA(cudaGetLastError());
    return kOk;
}
