// Block-wise elementwise engine: the per-block Python loops of the reference's vector operations as ONE launch each.
//
//   mode 0  LINCOMB  dst[d + i] = sum_k (+/-) src_k[s_k + i]          backend.add / sub (yastn/backend/backend_torch.py:518-534),
//                                                                      negate_blocks (_backend_torch_backwards.py:229-248, every
//                                                                      fermionic swap_gate)
//   mode 1  DIAG     dst[d + e] = src_0[s + e] * aux[a + (e / post) % naxis]      backend.dot_diag (backend_torch.py:557-564;
//                                                                      yastn.broadcast, tensordot with a diagonal tensor)
//   mode 2  GATHER   dst[d + e] = src_0[s + (p * nfull + idx[a + j]) * post + q],  e = (p * naxis + j) * post + q
//                                                                      apply_mask forward / embed_mask backward (:251-310)
//   mode 3  SCATTER  dst[d + (p * nfull + idx[a + j]) * post + q] = src_0[s + e]   embed_mask forward / apply_mask backward
//   mode 4  TRACE    dst[d + e] = sum_c sum_i src_0[base_c + i * dstride_c + off_c(e)]      backend.trace (backend_torch.py:268-275)
//
// The reference runs 2-6 torch launches per block for these; they are memory-bound and tiny per block, so the host loop is
// the cost.  Here the host cuts every record into pieces of <= 2048 elements, one WARP takes one piece at a time (the copy
// engine's execution model, yb_copy.cu), all index arithmetic is 32-bit with precomputed magic-number division.
#include <algorithm>

#include "yb_common.h"

namespace yb {

constexpr int kEwThreads = 256;
constexpr int kEwWarps = kEwThreads / 32;
constexpr uint32_t kEwPiece = 2048;
constexpr int kEwSrc = 4;
constexpr int kTraceDims = 6;
constexpr int64_t kAbsent = INT64_MIN;

enum EwMode : int { kLincomb = 0, kDiag = 1, kGather = 2, kScatter = 3, kTrace = 4 };

struct alignas(16) EwPiece {
    int64_t dst;
    int64_t src[kEwSrc];
    int64_t aux;
    uint32_t n, e0;
    uint32_t mode, neg;
    uint32_t post, naxis, nfull, pad0;
    uint32_t mul_post, shr_post, mul_axis, shr_axis;
};
static_assert(sizeof(EwPiece) == 96, "EwPiece is staged by one warp, one word per lane");
constexpr int kEwWords = sizeof(EwPiece) / 4;

struct alignas(16) EwTrace {      // one traced source block contributing to an output block
    int64_t base, dstride;
    int64_t str[kTraceDims];
    uint32_t ext[kTraceDims], mul[kTraceDims], shr[kTraceDims];
    uint32_t D, nd;
};

struct EwArgs {
    const EwPiece* pieces;
    const EwTrace* traces;
    int npieces;
    const void* src[kEwSrc];
    const void* aux;
    void* dst;
};

__device__ __forceinline__ uint32_t ew_div(uint32_t n, uint32_t d, uint32_t mul, uint32_t shr) {
    return d == 1 ? n : (__umulhi(n, mul) >> shr);
}
__device__ __forceinline__ double ew_zero(double) { return 0.0; }
__device__ __forceinline__ double2 ew_zero(double2) { return make_double2(0.0, 0.0); }
__device__ __forceinline__ double ew_add(double a, double b, bool neg) { return neg ? a - b : a + b; }
__device__ __forceinline__ double2 ew_add(double2 a, double2 b, bool neg) {
    return neg ? make_double2(a.x - b.x, a.y - b.y) : make_double2(a.x + b.x, a.y + b.y);
}
__device__ __forceinline__ double ew_mul(double a, double b) { return a * b; }
__device__ __forceinline__ double2 ew_mul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

template <typename T>
__global__ void __launch_bounds__(kEwThreads, 4) ewise_kernel(const EwArgs g) {
    __shared__ EwPiece spieces[kEwWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    EwPiece& P = spieces[warp];
    const int nwarps = gridDim.x * kEwWarps;
    T* __restrict__ dst = reinterpret_cast<T*>(g.dst);
    for (int it = blockIdx.x * kEwWarps + warp; it < g.npieces; it += nwarps) {
        __syncwarp();
        if (lane < kEwWords) reinterpret_cast<uint32_t*>(&P)[lane] = reinterpret_cast<const uint32_t*>(g.pieces + it)[lane];
        __syncwarp();
        const uint32_t n = P.n, e0 = P.e0;
        const int mode = (int)P.mode;
        if (mode == kLincomb) {
            const T* s[kEwSrc];
            bool on[kEwSrc];
#pragma unroll
            for (int k = 0; k < kEwSrc; ++k) {
                on[k] = P.src[k] != kAbsent;
                s[k] = reinterpret_cast<const T*>(g.src[k]) + (on[k] ? P.src[k] : 0);
            }
            T* d = dst + P.dst;
            const uint32_t neg = P.neg;
            for (uint32_t i = lane; i < n; i += 32 * 4) {
                T v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = ew_zero(T{});
#pragma unroll
                for (int k = 0; k < kEwSrc; ++k) {
                    if (on[k]) {      // warp-uniform
                        T x[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            x[u] = ew_zero(T{});
                            if (i + 32 * u < n) x[u] = s[k][i + 32 * u];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = ew_add(v[u], x[u], (neg >> k) & 1);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i + 32 * u < n) d[i + 32 * u] = v[u];
            }
        } else if (mode == kDiag) {
            const T* s = reinterpret_cast<const T*>(g.src[0]) + P.src[0];
            const T* a = reinterpret_cast<const T*>(g.aux) + P.aux;
            T* d = dst + P.dst;
            for (uint32_t i = lane; i < n; i += 32) {
                const uint32_t e = e0 + i;
                const uint32_t t = ew_div(e, P.post, P.mul_post, P.shr_post);
                const uint32_t j = t - ew_div(t, P.naxis, P.mul_axis, P.shr_axis) * P.naxis;
                d[e] = ew_mul(s[e], a[j]);
            }
        } else if (mode == kGather || mode == kScatter) {
            const T* s = reinterpret_cast<const T*>(g.src[0]) + P.src[0];
            const int64_t* idx = reinterpret_cast<const int64_t*>(g.aux) + P.aux;
            T* d = dst + P.dst;
            for (uint32_t i = lane; i < n; i += 32) {
                const uint32_t e = e0 + i;
                const uint32_t t = ew_div(e, P.post, P.mul_post, P.shr_post);
                const uint32_t q = e - t * P.post;
                const uint32_t p = ew_div(t, P.naxis, P.mul_axis, P.shr_axis);
                const uint32_t j = t - p * P.naxis;
                const int64_t other = ((int64_t)p * P.nfull + idx[j]) * P.post + q;
                if (mode == kGather) d[e] = s[other];
                else d[other] = s[e];
            }
        } else {   // kTrace
            const T* s = reinterpret_cast<const T*>(g.src[0]);
            T* d = dst + P.dst;
            for (uint32_t i = lane; i < n; i += 32) {
                const uint32_t e = e0 + i;
                T acc = ew_zero(T{});
                for (uint32_t c = 0; c < P.nfull; ++c) {
                    const EwTrace& R = g.traces[P.aux + c];
                    uint32_t rem = e;
                    int64_t off = R.base;
                    for (int k = (int)R.nd - 1; k >= 0; --k) {
                        const uint32_t qd = ew_div(rem, R.ext[k], R.mul[k], R.shr[k]);
                        off += (int64_t)(rem - qd * R.ext[k]) * R.str[k];
                        rem = qd;
                    }
                    for (uint32_t x = 0; x < R.D; ++x) acc = ew_add(acc, s[off + (int64_t)x * R.dstride], false);
                }
                d[e] = acc;
            }
        }
    }
}

}  // namespace yb

using namespace yb;

struct yb_ew_plan {
    int itemsize = 0, device = 0;
    int npieces = 0, grid = 0;
    int64_t elems = 0;
    DeviceTable pieces, traces;
};

// recs: nrec x 16 int64 rows
//   [0] mode  [1] dst  [2] n  [3..6] src offsets (INT64_MIN: absent)  [7] negate mask  [8] aux  [9] post  [10] naxis  [11] nfull
//   TRACE: [8] = first row of `traces`, [11] = number of rows;   traces: ntrace x 16 int64 rows
//   [0] base  [1] D  [2] dstride  [3] nd  [4..9] ext  [10..15] stride
extern "C" int yb_ew_plan_create(const int64_t* recs, int64_t nrec, const int64_t* traces, int64_t ntrace, int itemsize, int device,
                                 yb_ew_plan** out) {
    if (!out) return fail(kErrArg, "yb_ew_plan_create: out is null");
    *out = nullptr;
    if (nrec < 0 || ntrace < 0 || (nrec > 0 && !recs) || (ntrace > 0 && !traces)) return fail(kErrArg, "yb_ew_plan_create: bad table");
    if (itemsize != 8 && itemsize != 16) return fail(kErrUnsupported, "yb_ew_plan_create: itemsize %d (8 or 16)", itemsize);
    std::vector<EwPiece> pieces;
    int64_t elems = 0;
    const int64_t lim = (1ll << 31) - 1;
    for (int64_t r = 0; r < nrec; ++r) {
        const int64_t* q = recs + r * 16;
        const int64_t mode = q[0], n = q[2];
        if (mode < 0 || mode > kTrace) return fail(kErrArg, "yb_ew_plan_create: record %lld has mode %lld", (long long)r, (long long)mode);
        if (n < 0 || n > lim) return fail(kErrUnsupported, "yb_ew_plan_create: record %lld has %lld elements", (long long)r, (long long)n);
        EwPiece p;
        memset(&p, 0, sizeof(p));
        p.mode = (uint32_t)mode;
        p.neg = (uint32_t)q[7];
        p.aux = q[8];
        const int64_t post = std::max<int64_t>(q[9], 1), naxis = std::max<int64_t>(q[10], 1);
        if (post > lim || naxis > lim || q[11] > lim || q[11] < 0) return fail(kErrUnsupported, "yb_ew_plan_create: record %lld extents out of range", (long long)r);
        const FastDiv fp = make_fastdiv((uint32_t)post), fa = make_fastdiv((uint32_t)naxis);
        p.post = (uint32_t)post;
        p.naxis = (uint32_t)naxis;
        p.nfull = (uint32_t)q[11];
        p.mul_post = fp.mul;
        p.shr_post = fp.shr;
        p.mul_axis = fa.mul;
        p.shr_axis = fa.shr;
        if (mode == kTrace && (q[8] < 0 || q[8] + q[11] > ntrace)) return fail(kErrArg, "yb_ew_plan_create: record %lld trace range", (long long)r);
        for (int64_t e0 = 0; e0 < n; e0 += kEwPiece) {
            EwPiece c = p;
            c.n = (uint32_t)std::min<int64_t>(kEwPiece, n - e0);
            if (mode == kLincomb) {      // offsets advance with the piece, e0 is not used
                c.dst = q[1] + e0;
                for (int k = 0; k < kEwSrc; ++k) c.src[k] = q[3 + k] == kAbsent ? kAbsent : q[3 + k] + e0;
                c.e0 = 0;
            } else {
                c.dst = q[1];
                for (int k = 0; k < kEwSrc; ++k) c.src[k] = q[3 + k];
                c.e0 = (uint32_t)e0;
            }
            pieces.push_back(c);
        }
        elems += n;
    }
    std::vector<EwTrace> ht((size_t)ntrace);
    for (int64_t t = 0; t < ntrace; ++t) {
        const int64_t* q = traces + t * 16;
        EwTrace& R = ht[(size_t)t];
        memset(&R, 0, sizeof(R));
        R.base = q[0];
        R.dstride = q[2];
        if (q[1] < 0 || q[1] > lim || q[3] < 0 || q[3] > kTraceDims) return fail(kErrUnsupported, "yb_ew_plan_create: trace row %lld out of range", (long long)t);
        R.D = (uint32_t)q[1];
        R.nd = (uint32_t)q[3];
        for (int k = 0; k < kTraceDims; ++k) {
            const int64_t e = k < (int)R.nd ? q[4 + k] : 1;
            if (e < 1 || e > lim) return fail(kErrUnsupported, "yb_ew_plan_create: trace row %lld extent", (long long)t);
            const FastDiv f = make_fastdiv((uint32_t)e);
            R.ext[k] = (uint32_t)e;
            R.mul[k] = f.mul;
            R.shr[k] = f.shr;
            R.str[k] = k < (int)R.nd ? q[10 + k] : 0;
        }
    }
    yb_ew_plan* plan = new yb_ew_plan();
    plan->itemsize = itemsize;
    plan->device = device;
    plan->npieces = (int)pieces.size();
    plan->elems = elems;
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = kOk;
    if (cudaSetDevice(device) != cudaSuccess) rc = fail(kErrCuda, "yb_ew_plan_create: cudaSetDevice(%d) failed", device);
    if (rc == kOk) {
        TableBatch up;
        up.add(plan->pieces, pieces.data(), pieces.size() * sizeof(EwPiece));
        up.add(plan->traces, ht.data(), ht.size() * sizeof(EwTrace));
        rc = up.commit();
    }
    if (rc == kOk) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        plan->grid = std::max(1, std::min((plan->npieces + kEwWarps - 1) / kEwWarps, sms * 4));
    }
    cudaSetDevice(prev);
    if (rc != kOk) {
        plan->pieces.release();
        plan->traces.release();
        delete plan;
        return rc;
    }
    *out = plan;
    return kOk;
}

extern "C" int yb_ew_plan_info(const yb_ew_plan* plan, int64_t info[2]) {
    if (!plan || !info) return fail(kErrArg, "yb_ew_plan_info: null argument");
    info[0] = plan->npieces;
    info[1] = plan->elems;
    return kOk;
}

extern "C" int yb_ew_run(const yb_ew_plan* plan, void* dst, const void* src0, const void* src1, const void* src2, const void* src3,
                         const void* aux, void* stream) {
    if (!plan) return fail(kErrArg, "yb_ew_run: plan is null");
    if (plan->npieces == 0) return kOk;
    if (!dst) return fail(kErrArg, "yb_ew_run: dst is null");
    EwArgs a;
    a.pieces = (const EwPiece*)plan->pieces.ptr;
    a.traces = (const EwTrace*)plan->traces.ptr;
    a.npieces = plan->npieces;
    a.src[0] = src0;
    a.src[1] = src1;
    a.src[2] = src2;
    a.src[3] = src3;
    a.aux = aux;
    a.dst = dst;
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->itemsize == 8)
        ewise_kernel<double><<<plan->grid, kEwThreads, 0, st>>>(a);
    else
        ewise_kernel<double2><<<plan->grid, kEwThreads, 0, st>>>(a);
    YB_CUDA(cudaGetLastError());
    return kOk;
}

extern "C" void yb_ew_plan_destroy(yb_ew_plan* plan) {
    if (!plan) return;
    plan->pieces.release();
    plan->traces.release();
    delete plan;
}
