// Panel products: a tiny matrix (at most 8 x 8) times a very long one — C[M x N] = A[M x K] B[K x N] with K <= 8 and one of
// M, N <= 8 while the other runs to 10^5 .. 10^7.
//
// This is what applying an MPO tensor to an environment or to a two-site tensor looks like after fuse_to_matrix
// (yastn/tn/mps/_env.py:512-518 Heff2, :496-504 update_env_): per charge sector the MPO block is a 1..4 x 1..4 matrix and
// the other operand has one row (column) per element of the three remaining legs.  On 64 x 64 DMMA tiles with a 16-wide
// k-step, 99.6 % of the tensor pipe works on padding and every tile costs a full pipeline fill for one k-iteration: in a
// D = 4096 Hubbard sweep these calls took 4.5 of 7.3 s of GEMM time at 0.05 TFLOP/s (profiles/dmrg_gemm_classes_r02.json)
// although they are pure streaming — read the long operand once, write the result once, K*X multiply-adds per element on
// the CUDA cores (far below the FP64 ridge).  One WARP takes 1024 consecutive streaming indices of one problem at a time;
// the small matrix sits in shared memory; every lane keeps U independent columns in flight.
#include <algorithm>

#include "yb_gemm_types.h"

namespace yb {

constexpr int kPnThreads = 256;
constexpr int kPnWarps = kPnThreads / 32;
constexpr int kPnPart = 1024;        // streaming indices per work unit

struct PnProb {
    int32_t prob;        // index into the GEMM problem table
    int32_t big_is_b;    // 1: M <= 8, the streaming index is n (C = S B);  0: N <= 8, the streaming index is m (C = A S)
};

struct PnArgs {
    const PnProb* probs;
    const int64_t* pstart;   // [nprob + 1] first work unit of every panel problem
    const GemmProblem* problems;
    const GemmSegment* segs;
    ScatterTables scat;
    const char* A;
    const char* B;
    char* C;
    int nprob;
    int64_t nparts;
    int flags;
};

template <bool CPLX>
struct PnT {
    using T = typename std::conditional<CPLX, double2, double>::type;
};
__device__ __forceinline__ double pn_zero(double) { return 0.0; }
__device__ __forceinline__ double2 pn_zero(double2) { return make_double2(0.0, 0.0); }
__device__ __forceinline__ void pn_fma(double& acc, double a, double b) { acc = fma(a, b, acc); }
__device__ __forceinline__ void pn_fma(double2& acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ double pn_conj(double v, bool) { return v; }
__device__ __forceinline__ double2 pn_conj(double2 v, bool c) { return c ? make_double2(v.x, -v.y) : v; }

template <bool CPLX, int XT, int KT>
__global__ void __launch_bounds__(kPnThreads) panel_kernel(const PnArgs g) {
    using T = typename PnT<CPLX>::T;
    // independent columns per lane: enough loads in flight (>= 128 bytes per lane for the KT x U operand elements) without
    // spilling the XT x U accumulators
    constexpr int kUk = (CPLX ? 8 : 16) / KT, kUx = (CPLX ? 16 : 32) / XT;
    constexpr int U = kUk < kUx ? (kUk < 1 ? 1 : kUk) : (kUx < 1 ? 1 : kUx);
    __shared__ T smat[kPnWarps][XT * KT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T* S = smat[warp];
    const T* __restrict__ A = reinterpret_cast<const T*>(g.A);
    const T* __restrict__ B = reinterpret_cast<const T*>(g.B);
    T* __restrict__ C = reinterpret_cast<T*>(g.C);
    const bool conjA = CPLX && (g.flags & YB_GEMM_CONJ_A), conjB = CPLX && (g.flags & YB_GEMM_CONJ_B);
    const int64_t nwarps = (int64_t)gridDim.x * kPnWarps;
    for (int64_t part = (int64_t)blockIdx.x * kPnWarps + warp; part < g.nparts; part += nwarps) {
        // problem of this work unit: largest i with pstart[i] <= part
        int lo = 0, hi = g.nprob;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (g.pstart[mid] <= part) lo = mid;
            else hi = mid;
        }
        const PnProb pp = g.probs[lo];
        const GemmProblem P = g.problems[pp.prob];
        const bool rb = pp.big_is_b != 0;
        const int X = rb ? P.M : P.N;                       // small output extent
        const int64_t slen = rb ? P.N : P.M;                // streaming extent
        const int64_t s0 = (part - g.pstart[lo]) * kPnPart;
        const int64_t s1 = min(slen, s0 + kPnPart);
        const bool one_seg = P.seg_end - P.seg_begin == 1;
        auto load_small = [&](const GemmSegment& Sg) {      // S[x][k], zero outside the X x K matrix: padded products vanish
            const T* sm = (rb ? A : B) + (rb ? Sg.offA : Sg.offB);
            const int64_t sx = rb ? Sg.sAm : Sg.sBn, sk = rb ? Sg.sAk : Sg.sBk;
            const bool conjS = rb ? conjA : conjB;
            __syncwarp();
            for (int e = lane; e < XT * KT; e += 32) {
                const int x = e / KT, k = e % KT;
                S[e] = (x < X && k < Sg.K) ? pn_conj(sm[x * sx + k * sk], conjS) : pn_zero(T{});
            }
            __syncwarp();
        };
        if (one_seg) load_small(g.segs[P.seg_begin]);
        for (int64_t sb = s0; sb < s1; sb += 32 * U) {
            T acc[U][XT];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int x = 0; x < XT; ++x) acc[u][x] = pn_zero(T{});
            for (int sg = P.seg_begin; sg < P.seg_end; ++sg) {
                const GemmSegment Sg = g.segs[sg];
                const int K = Sg.K;
                if (K == 0) continue;
                if (!one_seg) load_small(Sg);
                const T* G = (rb ? B : A) + (rb ? Sg.offB : Sg.offA);       // the long operand G(k, s)
                const int64_t gk = rb ? Sg.sBk : Sg.sAk, gs = rb ? Sg.sBn : Sg.sAm;
                const bool conjG = rb ? conjB : conjA;
                T gv[U][KT];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int64_t s = sb + lane + 32 * u;
                    const int64_t sc = s < s1 ? s : s1 - 1;              // clamped: branch-free loads, the value is discarded below
#pragma unroll
                    for (int k = 0; k < KT; ++k) {
                        const int kc = k < K ? k : K - 1;
                        gv[u][k] = pn_conj(G[kc * gk + sc * gs], conjG);
                    }
                }
#pragma unroll
                for (int x = 0; x < XT; ++x)
#pragma unroll
                    for (int k = 0; k < KT; ++k) {
                        const T sv = S[x * KT + k];
#pragma unroll
                        for (int u = 0; u < U; ++u) pn_fma(acc[u][x], sv, gv[u][k]);
                    }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t s = sb + lane + 32 * u;
                if (s < s1) {
#pragma unroll
                    for (int x = 0; x < XT; ++x)
                        if (x < X) {
                            const int r = rb ? x : (int)s, c = rb ? (int)s : x;
                            C[c_offset(P, g.scat, r, c)] = acc[u][x];
                        }
                }
            }
        }
    }
}

// Problems are bucketed by the padded shape (XT, KT) of their small matrix: a 1 x 1 MPO block inside a launch specialised for
// 4 x 4 would load every operand element four times and do sixteen multiply-adds for one (the charge sectors of a U1xU1 MPO are
// mostly 1 x 1 and 2 x 2: profiles/panel_probe_r02.jsonl).  One launch per bucket; tiny plans keep a single bucket.
struct PanelBucket {
    int xt = 1, kt = 1;
    int nprob = 0, grid = 0;
    int64_t nparts = 0;
    size_t prob_off = 0, pstart_off = 0;     // element offsets into the plan's tables
};

struct PanelPlan {
    bool cplx = false;
    int device = 0;
    int nprob = 0;
    int64_t nparts = 0;
    std::vector<PanelBucket> buckets;
    DeviceTable probs, pstart;
};

namespace {

int p2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

template <bool CPLX, int XT>
int launch_k(int kt, int grid, const PnArgs& a, cudaStream_t st) {
    switch (kt) {
        case 1: panel_kernel<CPLX, XT, 1><<<grid, kPnThreads, 0, st>>>(a); break;
        case 2: panel_kernel<CPLX, XT, 2><<<grid, kPnThreads, 0, st>>>(a); break;
        case 4: panel_kernel<CPLX, XT, 4><<<grid, kPnThreads, 0, st>>>(a); break;
        default: panel_kernel<CPLX, XT, 8><<<grid, kPnThreads, 0, st>>>(a); break;
    }
    YB_CUDA(cudaGetLastError());
    return kOk;
}

template <bool CPLX>
int launch_x(int xt, int kt, int grid, const PnArgs& a, cudaStream_t st) {
    switch (xt) {
        case 1: return launch_k<CPLX, 1>(kt, grid, a, st);
        case 2: return launch_k<CPLX, 2>(kt, grid, a, st);
        case 4: return launch_k<CPLX, 4>(kt, grid, a, st);
        default: return launch_k<CPLX, 8>(kt, grid, a, st);
    }
}

}  // namespace

bool panel_eligible(const GemmProblem& P, const std::vector<GemmSegment>& hs) {
    if (P.M <= 0 || P.N <= 0) return false;
    const bool small_m = P.M <= kPanelMax, small_n = P.N <= kPanelMax;
    if (small_m == small_n) return false;                       // both small: skinny path; both large: tiles
    if ((small_m ? P.N : P.M) < kPanelMinStream) return false;
    for (int s = P.seg_begin; s < P.seg_end; ++s)
        if (hs[(size_t)s].K > kPanelMax) return false;
    return true;
}

int panel_create(const std::vector<GemmProblem>& hp, const std::vector<GemmSegment>& hs, const std::vector<int>& which, bool cplx,
                 int device, PanelPlan** out) {
    *out = nullptr;
    PanelPlan* plan = new PanelPlan();
    plan->cplx = cplx;
    plan->device = device;
    plan->nprob = (int)which.size();
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    // padded shape of every problem; a plan whose streaming extent is small keeps ONE bucket (a launch costs more than the padding)
    struct Shape {
        int idx, xt, kt;
        int64_t slen;
    };
    std::vector<Shape> shapes;
    int64_t total = 0;
    int mx = 1, mk = 1;
    for (int idx : which) {
        const GemmProblem& P = hp[(size_t)idx];
        const bool rb = P.M <= kPanelMax;
        int k = 1;
        for (int sg = P.seg_begin; sg < P.seg_end; ++sg) k = std::max(k, hs[(size_t)sg].K);
        Shape sh = {idx, p2(rb ? P.M : P.N), p2(k), (int64_t)(rb ? P.N : P.M)};
        mx = std::max(mx, sh.xt);
        mk = std::max(mk, sh.kt);
        total += sh.slen;
        shapes.push_back(sh);
    }
    const bool split = total >= (int64_t)1 << 18;
    if (!split)
        for (auto& sh : shapes) {
            sh.xt = mx;
            sh.kt = mk;
        }
    std::stable_sort(shapes.begin(), shapes.end(), [](const Shape& a, const Shape& b) { return a.xt != b.xt ? a.xt < b.xt : a.kt < b.kt; });
    std::vector<PnProb> probs;
    std::vector<int64_t> pstart;
    for (size_t i = 0; i < shapes.size();) {
        PanelBucket bk;
        bk.xt = shapes[i].xt;
        bk.kt = shapes[i].kt;
        bk.prob_off = probs.size();
        bk.pstart_off = pstart.size();
        int64_t parts = 0;
        size_t j = i;
        for (; j < shapes.size() && shapes[j].xt == bk.xt && shapes[j].kt == bk.kt; ++j) {
            const GemmProblem& P = hp[(size_t)shapes[j].idx];
            probs.push_back({shapes[j].idx, P.M <= kPanelMax ? 1 : 0});
            pstart.push_back(parts);
            parts += (shapes[j].slen + kPnPart - 1) / kPnPart;
        }
        pstart.push_back(parts);
        bk.nprob = (int)(j - i);
        bk.nparts = parts;
        bk.grid = (int)std::max<int64_t>(1, std::min<int64_t>((parts + kPnWarps - 1) / kPnWarps, (int64_t)sms * 4));
        plan->nparts += parts;
        plan->buckets.push_back(bk);
        i = j;
    }
    TableBatch up;
    up.add(plan->probs, probs.data(), probs.size() * sizeof(PnProb));
    up.add(plan->pstart, pstart.data(), pstart.size() * sizeof(int64_t));
    int rc = up.commit();
    if (rc != kOk) {
        panel_destroy(plan);
        return rc;
    }
    *out = plan;
    return kOk;
}

int panel_run(const PanelPlan* plan, const GemmProblem* problems, const GemmSegment* segs, const ScatterTables& scat, const void* A,
              const void* B, void* C, int flags, cudaStream_t st) {
    if (plan->nparts == 0) return kOk;
    for (const PanelBucket& bk : plan->buckets) {
        if (bk.nparts == 0) continue;
        PnArgs a;
        a.probs = (const PnProb*)plan->probs.ptr + bk.prob_off;
        a.pstart = (const int64_t*)plan->pstart.ptr + bk.pstart_off;
        a.problems = problems;
        a.segs = segs;
        a.scat = scat;
        a.A = (const char*)A;
        a.B = (const char*)B;
        a.C = (char*)C;
        a.nprob = bk.nprob;
        a.nparts = bk.nparts;
        a.flags = flags;
        int rc = plan->cplx ? launch_x<true>(bk.xt, bk.kt, bk.grid, a, st) : launch_x<false>(bk.xt, bk.kt, bk.grid, a, st);
        if (rc != kOk) return rc;
    }
    return kOk;
}

void panel_destroy(PanelPlan* plan) {
    if (!plan) return;
    plan->probs.release();
    plan->pstart.release();
    delete plan;
}

int64_t panel_parts(const PanelPlan* plan) { return plan ? plan->nparts : 0; }

}  // namespace yb
