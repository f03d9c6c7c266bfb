// Skinny-output block GEMM: problems whose result block is at most 8 x 8 while the contraction index is long.
//
// Callers (reference loops): backend.vdot (yastn/backend/backend_torch.py:537-546, one torch.dot per block, every Lanczos
// step), the environment-overlap contractions whose two big legs are both contracted (SURVEY.md 8d pattern P3: K ~ 10^7,
// M = N <= 4), and the adjoint B_b = A^H @ C_b of backend.dot (yastn/backend/_backend_torch_backwards.py:136-137) for
// tall-and-skinny operands.  On 64 x 64 DMMA tiles these spend > 99 % of the tensor pipe on padding; they are pure
// HBM-bound reductions: every operand element is read once and takes part in at most 8 multiply-adds, far below the
// ~5 FLOP/B ridge of the FP64 pipes, so plain DFMA on the CUDA cores is the right instruction.
//
// One WARP walks a contiguous piece of the work line (all contraction indices of all problems, cut into equal weights on
// the host); lane l takes contraction indices c0 + l, c0 + l + 32, ... so that operands with a unit contraction stride are
// read fully coalesced, and operands stored as rows of X (<= 8) contiguous elements are read with 16-byte loads when
// aligned.  Every lane keeps >= 128 bytes of independent loads in flight.  A run of pieces of one problem ends in a
// butterfly all-reduce; a problem cut over several warps is finished by whichever warp arrives last (atomic counter),
// which adds the partials in RUN ORDER — the result is bit-identical from launch to launch, no warp ever waits for
// another one (no forward-progress assumption), and the counters are left at zero (CUDA-graph replay safe).
#include <algorithm>

#include "yb_gemm_types.h"

namespace yb {

constexpr int kSkThreads = 256;
constexpr int kSkWarps = kSkThreads / 32;
constexpr int kSkSlot = kSkinnyMax * kSkinnyMax;   // entries of one partial result
constexpr int64_t kSkPartOverhead = 512;           // weight of fetching a part's tables, in operand elements
constexpr int64_t kSkMinChunk = 8192;              // least weight of a warp's share

struct SkPart {
    int32_t prob, lp;      // index into the GEMM problem table / dense index among the skinny problems
    int32_t seg;           // segment, -1: no contraction (the block is zero)
    int32_t c0, c1;        // contraction range inside the segment
    int32_t run;           // partial-result slot
    int32_t flags;         // bit0: first part of its run, bit1: last part of its run
    int32_t pad_;
};

struct SkWarp {
    int32_t part_begin, part_end;
};

struct SkRuns {
    int32_t run_begin, run_end;
};

struct SkArgs {
    const SkPart* parts;
    const SkWarp* warps;
    const SkRuns* runs;
    const GemmProblem* problems;
    const GemmSegment* segs;
    ScatterTables scat;
    const char* A;
    const char* B;
    char* C;
    char* ws;
    int* counters;
    int nwarps;
    int flags;
};

template <bool CPLX>
struct SkT {
    using T = typename std::conditional<CPLX, double2, double>::type;
};

__device__ __forceinline__ double sk_zero(double) { return 0.0; }
__device__ __forceinline__ double2 sk_zero(double2) { return make_double2(0.0, 0.0); }
__device__ __forceinline__ void sk_fma(double& acc, double a, double b) { acc = fma(a, b, acc); }
__device__ __forceinline__ void sk_fma(double2& acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ double sk_add(double a, double b) { return a + b; }
__device__ __forceinline__ double2 sk_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double sk_shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double2 sk_shfl_xor(double2 v, int m) {
    return make_double2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
__device__ __forceinline__ double sk_conj(double v, bool) { return v; }
__device__ __forceinline__ double2 sk_conj(double2 v, bool c) { return c ? make_double2(v.x, -v.y) : v; }
__device__ __forceinline__ double sk_ldcg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ double2 sk_ldcg(const double2* p) { return __ldcg(p); }
__device__ __forceinline__ void sk_stcg(double* p, double v) { __stcg(p, v); }
__device__ __forceinline__ void sk_stcg(double2* p, double2 v) { __stcg(p, v); }

// Loads the NT elements (x = 0..NT) of one operand at contraction index c; elements beyond `ext`, and everything when !ok,
// are zero.  Branch-free on purpose: out-of-range indices are CLAMPED to a valid element and the value is discarded by a
// select, so all loads of a trip are issued back to back (with guarded loads the compiler emitted one branch region per
// load and the warp waited for each in turn: ncu showed 38 stall cycles per issue on the long scoreboard, 15 % of HBM).
//   full (sx == 1, ext == NT, float64, rows 16-byte aligned): a row of NT contiguous elements per contraction index is
//   read with 16-byte loads.
template <typename T, int NT, bool FULL>
__device__ __forceinline__ void sk_load(const T* __restrict__ base, int64_t sx, int64_t sc, int64_t c, int ext, bool ok, bool conj, T* out) {
    const T* row = base + c * sc;
    if constexpr (FULL) {
        static_assert(sizeof(T) == 8 && NT >= 2, "16-byte row loads are a float64 path");
#pragma unroll
        for (int x = 0; x < NT; x += 2) {
            const double2 v = *reinterpret_cast<const double2*>(row + x);
            out[x] = ok ? v.x : 0.0;
            out[x + 1] = ok ? v.y : 0.0;
        }
    } else {
#pragma unroll
        for (int x = 0; x < NT; ++x) {
            const int xs = x < ext ? x : ext - 1;
            const T v = sk_conj(row[xs * sx], conj);
            out[x] = (ok && x < ext) ? v : sk_zero(T{});
        }
    }
}

// The trips of one part.  VA / VB (16-byte row loads of A / B) are template parameters so that the loop body is one basic
// block: every lane runs the same number of trips (lanes past the end work on a clamped index and discard the value).
template <bool CPLX, int XT, int YT, bool VA, bool VB>
__device__ __forceinline__ void sk_accumulate(typename SkT<CPLX>::T (&acc)[XT][YT], const typename SkT<CPLX>::T* __restrict__ pa,
                                              const typename SkT<CPLX>::T* __restrict__ pb, const GemmSegment& S, int M, int N,
                                              int64_t c_begin, int64_t c_end, int lane, bool conjA, bool conjB) {
    using T = typename SkT<CPLX>::T;
    constexpr int U0 = (CPLX ? 8 : 16) / (XT + YT);          // >= 128 bytes of loads in flight per lane
    constexpr int U = U0 < 1 ? 1 : (U0 > 8 ? 8 : U0);
    const int64_t clast = c_end - 1;
    for (int64_t c0 = c_begin; c0 < c_end; c0 += 32 * U) {
        T a[U][XT], b[U][YT];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t cc = c0 + lane + 32 * u;
            const bool ok = cc <= clast;
            const int64_t cs = ok ? cc : clast;
            sk_load<T, XT, VA>(pa, S.sAm, S.sAk, cs, M, ok, conjA, a[u]);
            sk_load<T, YT, VB>(pb, S.sBn, S.sBk, cs, N, ok, conjB, b[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int x = 0; x < XT; ++x)
#pragma unroll
                for (int y = 0; y < YT; ++y) sk_fma(acc[x][y], a[u][x], b[u][y]);
    }
}

template <bool CPLX, int XT, int YT>
__global__ void __launch_bounds__(kSkThreads) skinny_kernel(const SkArgs g) {
    using T = typename SkT<CPLX>::T;
    constexpr int NE = XT * YT;
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * kSkWarps + (threadIdx.x >> 5);
    if (w >= g.nwarps) return;
    const SkWarp W = g.warps[w];
    const bool conjA = CPLX && (g.flags & YB_GEMM_CONJ_A), conjB = CPLX && (g.flags & YB_GEMM_CONJ_B);
    const T* __restrict__ A = reinterpret_cast<const T*>(g.A);
    const T* __restrict__ B = reinterpret_cast<const T*>(g.B);
    T* __restrict__ C = reinterpret_cast<T*>(g.C);
    T* ws = reinterpret_cast<T*>(g.ws);

    T acc[XT][YT];
    for (int p = W.part_begin; p < W.part_end; ++p) {
        const SkPart part = g.parts[p];
        const GemmProblem P = g.problems[part.prob];
        if (part.flags & 1) {
#pragma unroll
            for (int x = 0; x < XT; ++x)
#pragma unroll
                for (int y = 0; y < YT; ++y) acc[x][y] = sk_zero(T{});
        }
        if (part.seg >= 0) {
            const GemmSegment S = g.segs[part.seg];
            const T* pa = A + S.offA;
            const T* pb = B + S.offB;
            const bool vecA = !CPLX && XT >= 2 && P.M == XT && S.sAm == 1 && ((reinterpret_cast<uintptr_t>(pa) & 15) == 0) && ((S.sAk & 1) == 0);
            const bool vecB = !CPLX && YT >= 2 && P.N == YT && S.sBn == 1 && ((reinterpret_cast<uintptr_t>(pb) & 15) == 0) && ((S.sBk & 1) == 0);
            constexpr bool CANA = !CPLX && XT >= 2, CANB = !CPLX && YT >= 2;
            if (CANA && CANB && vecA && vecB)
                sk_accumulate<CPLX, XT, YT, CANA, CANB>(acc, pa, pb, S, P.M, P.N, part.c0, part.c1, lane, conjA, conjB);
            else if (CANA && vecA)
                sk_accumulate<CPLX, XT, YT, CANA, false>(acc, pa, pb, S, P.M, P.N, part.c0, part.c1, lane, conjA, conjB);
            else if (CANB && vecB)
                sk_accumulate<CPLX, XT, YT, false, CANB>(acc, pa, pb, S, P.M, P.N, part.c0, part.c1, lane, conjA, conjB);
            else
                sk_accumulate<CPLX, XT, YT, false, false>(acc, pa, pb, S, P.M, P.N, part.c0, part.c1, lane, conjA, conjB);
        }
        if (!(part.flags & 2)) continue;

        // ---- end of a run: butterfly all-reduce (addition commutes, so every lane ends with the same bits)
#pragma unroll
        for (int x = 0; x < XT; ++x)
#pragma unroll
            for (int y = 0; y < YT; ++y) {
                T v = acc[x][y];
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) v = sk_add(v, sk_shfl_xor(v, m));
                acc[x][y] = v;
            }
        const SkRuns R = g.runs[part.lp];
        const bool single = R.run_end - R.run_begin == 1;
        bool last = single;
        if (!single) {
            T* slot = ws + (size_t)part.run * kSkSlot;
#pragma unroll
            for (int r = 0; r < (NE + 31) / 32; ++r) {
                T v = sk_zero(T{});
#pragma unroll
                for (int e = r * 32; e < NE && e < r * 32 + 32; ++e)
                    if (lane == e - r * 32) v = acc[e / YT][e % YT];
                if (r * 32 + lane < NE) sk_stcg(slot + r * 32 + lane, v);
            }
            __threadfence();
            __syncwarp();
            int prev = 0;
            if (lane == 0) prev = atomicAdd(g.counters + part.lp, 1);
            prev = __shfl_sync(0xffffffffu, prev, 0);
            last = prev == R.run_end - R.run_begin - 1;
            if (last) __threadfence();
        }
        if (!last) continue;
#pragma unroll
        for (int r = 0; r < (NE + 31) / 32; ++r) {
            const int e = r * 32 + lane;
            T v = sk_zero(T{});
            if (single) {
#pragma unroll
                for (int q = r * 32; q < NE && q < r * 32 + 32; ++q)
                    if (lane == q - r * 32) v = acc[q / YT][q % YT];
            } else if (e < NE) {
                for (int run = R.run_begin; run < R.run_end; ++run) v = sk_add(v, sk_ldcg(ws + (size_t)run * kSkSlot + e));
            }
            const int x = e / YT, y = e % YT;
            if (e < NE && x < P.M && y < P.N) C[c_offset(P, g.scat, x, y)] = v;
        }
        if (!single && lane == 0) g.counters[part.lp] = 0;   // counters rest at 0 between launches (graph-replay safe)
    }
}

struct SkinnyPlan {
    bool cplx = false;
    int device = 0;
    int xt = 1, yt = 1;
    int nwarps = 0, nruns = 0, nlp = 0;
    DeviceTable parts, warps, runs;
};

namespace {

int pow2ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

template <bool CPLX, int XT, int YT>
int launch_xy(int grid, const SkArgs& a, cudaStream_t st) {
    if constexpr (XT * YT > skinny_entries(CPLX)) {
        return fail(kErrUnsupported, "yb_skinny: %d x %d accumulators do not fit in registers", XT, YT);
    } else {
        skinny_kernel<CPLX, XT, YT><<<grid, kSkThreads, 0, st>>>(a);
        YB_CUDA(cudaGetLastError());
        return kOk;
    }
}

template <bool CPLX, int XT>
int launch_y(int yt, int grid, const SkArgs& a, cudaStream_t st) {
    switch (yt) {
        case 1: return launch_xy<CPLX, XT, 1>(grid, a, st);
        case 2: return launch_xy<CPLX, XT, 2>(grid, a, st);
        case 4: return launch_xy<CPLX, XT, 4>(grid, a, st);
        default: return launch_xy<CPLX, XT, 8>(grid, a, st);
    }
}

template <bool CPLX, int XT, int YT>
int occupancy_xy(int* occ) {
    if constexpr (XT * YT > skinny_entries(CPLX)) {
        *occ = 0;
        return kOk;
    } else {
        YB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, skinny_kernel<CPLX, XT, YT>, kSkThreads, 0));
        return kOk;
    }
}

template <bool CPLX, int XT>
int occupancy_y(int yt, int* occ) {
    switch (yt) {
        case 1: return occupancy_xy<CPLX, XT, 1>(occ);
        case 2: return occupancy_xy<CPLX, XT, 2>(occ);
        case 4: return occupancy_xy<CPLX, XT, 4>(occ);
        default: return occupancy_xy<CPLX, XT, 8>(occ);
    }
}

template <bool CPLX>
int occupancy_x(int xt, int yt, int* occ) {
    static int cache[2][4][4] = {};   // [cplx][log2 xt][log2 yt]; the answer depends on the kernel only
    auto lg = [](int v) { return v == 1 ? 0 : (v == 2 ? 1 : (v == 4 ? 2 : 3)); };
    int& c = cache[CPLX ? 1 : 0][lg(xt)][lg(yt)];
    if (c > 0) {
        *occ = c;
        return kOk;
    }
    int rc;
    switch (xt) {
        case 1: rc = occupancy_y<CPLX, 1>(yt, occ); break;
        case 2: rc = occupancy_y<CPLX, 2>(yt, occ); break;
        case 4: rc = occupancy_y<CPLX, 4>(yt, occ); break;
        default: rc = occupancy_y<CPLX, 8>(yt, occ); break;
    }
    if (rc == kOk) c = *occ;
    return rc;
}

template <bool CPLX>
int launch_x(int xt, int yt, int grid, const SkArgs& a, cudaStream_t st) {
    switch (xt) {
        case 1: return launch_y<CPLX, 1>(yt, grid, a, st);
        case 2: return launch_y<CPLX, 2>(yt, grid, a, st);
        case 4: return launch_y<CPLX, 4>(yt, grid, a, st);
        default: return launch_y<CPLX, 8>(yt, grid, a, st);
    }
}

}  // namespace

int skinny_create(const std::vector<GemmProblem>& hp, const std::vector<GemmSegment>& hs, const std::vector<int>& which,
                  bool cplx, int device, SkinnyPlan** out) {
    *out = nullptr;
    SkinnyPlan* plan = new SkinnyPlan();
    plan->cplx = cplx;
    plan->device = device;
    plan->nlp = (int)which.size();
    int maxM = 1, maxN = 1;
    int64_t W = 0;
    for (int idx : which) {
        const GemmProblem& P = hp[(size_t)idx];
        maxM = std::max(maxM, (int)P.M);
        maxN = std::max(maxN, (int)P.N);
        W += kSkPartOverhead;
        for (int s = P.seg_begin; s < P.seg_end; ++s) W += (int64_t)hs[(size_t)s].K * (P.M + P.N) + kSkPartOverhead;
    }
    plan->xt = pow2ceil(maxM);
    plan->yt = pow2ceil(maxN);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    // one share per RESIDENT warp: the kernels use 64..250 registers, i.e. 1..4 CTAs per SM (a grid of 4 CTAs per SM on a
    // kernel that fits 2 ran as 2.05 waves in the first version)
    int occ = 0;
    int rc0 = cplx ? occupancy_x<true>(plan->xt, plan->yt, &occ) : occupancy_x<false>(plan->xt, plan->yt, &occ);
    if (rc0 != kOk || occ < 1) occ = 1;
    const int64_t max_warps = (int64_t)sms * kSkWarps * occ;
    // shares: cut the work line into pieces of `chunk` weight.  Every warp of the grid must be resident at once (a second,
    // partial wave doubles the run time of this latency-bound kernel: 607 CTAs on 592 slots cost 2x in the first version),
    // so the chunk grows until the number of shares fits.
    int64_t chunk = std::max<int64_t>(kSkMinChunk, (W + max_warps - 1) / max_warps);
    std::vector<SkPart> parts;
    std::vector<SkWarp> warps;
    for (int attempt = 0; attempt < 16; ++attempt) {
        parts.clear();
        warps.clear();
        int64_t cur = 0;
        int warp_begin = 0;
        auto close_warp = [&]() {
            if ((int)parts.size() > warp_begin) {
                warps.push_back({warp_begin, (int)parts.size()});
                warp_begin = (int)parts.size();
            }
            cur = 0;
        };
        for (int lp = 0; lp < plan->nlp; ++lp) {
            const int idx = which[(size_t)lp];
            const GemmProblem& P = hp[(size_t)idx];
            const int64_t per = std::max(1, P.M + P.N);
            bool any = false;
            for (int s = P.seg_begin; s < P.seg_end; ++s) {
                const int64_t K = hs[(size_t)s].K;
                int64_t c = 0;
                while (c < K) {
                    if (cur + kSkPartOverhead >= chunk) close_warp();
                    const int64_t room = chunk - cur - kSkPartOverhead;
                    const int64_t take = std::min(K - c, std::max<int64_t>(32, room / per));
                    parts.push_back({idx, lp, s, (int32_t)c, (int32_t)(c + take), 0, 0, 0});
                    any = true;
                    cur += take * per + kSkPartOverhead;
                    c += take;
                }
            }
            if (!any) {   // nothing to contract: the block is zero
                if (cur + kSkPartOverhead >= chunk) close_warp();
                parts.push_back({idx, lp, -1, 0, 0, 0, 0, 0});
                cur += kSkPartOverhead;
            }
        }
        close_warp();
        if ((int64_t)warps.size() <= max_warps) break;
        chunk = chunk + chunk / 16 + 64;
    }
    std::vector<SkRuns> runs((size_t)plan->nlp, SkRuns{-1, -1});
    // runs: maximal sequences of parts of one problem inside one warp
    int nruns = 0;
    for (const SkWarp& w : warps) {
        for (int p = w.part_begin; p < w.part_end; ++p) {
            const bool first = p == w.part_begin || parts[(size_t)p - 1].lp != parts[(size_t)p].lp;
            const bool last = p + 1 == w.part_end || parts[(size_t)p + 1].lp != parts[(size_t)p].lp;
            if (first) {
                if (runs[(size_t)parts[(size_t)p].lp].run_begin < 0) runs[(size_t)parts[(size_t)p].lp].run_begin = nruns;
                ++nruns;
            }
            parts[(size_t)p].run = nruns - 1;
            parts[(size_t)p].flags = (first ? 1 : 0) | (last ? 2 : 0);
            runs[(size_t)parts[(size_t)p].lp].run_end = nruns;
        }
    }
    plan->nwarps = (int)warps.size();
    plan->nruns = nruns;
    TableBatch up;
    up.add(plan->parts, parts.data(), parts.size() * sizeof(SkPart));
    up.add(plan->warps, warps.data(), warps.size() * sizeof(SkWarp));
    up.add(plan->runs, runs.data(), runs.size() * sizeof(SkRuns));
    int rc = up.commit();
    if (rc != kOk) {
        skinny_destroy(plan);
        return rc;
    }
    *out = plan;
    return kOk;
}

int skinny_run(const SkinnyPlan* plan, const GemmProblem* problems, const GemmSegment* segs, const ScatterTables& scat,
               const void* A, const void* B, void* C, int flags, cudaStream_t st) {
    if (plan->nwarps == 0) return kOk;
    void* ws = nullptr;
    int* counters = nullptr;
    const size_t ws_bytes = (size_t)plan->nruns * kSkSlot * (plan->cplx ? 16 : 8);
    int rc = stream_workspace(plan->device, st, ws_bytes, (size_t)plan->nlp, &ws, &counters);
    if (rc != kOk) return rc;
    SkArgs a;
    a.parts = (const SkPart*)plan->parts.ptr;
    a.warps = (const SkWarp*)plan->warps.ptr;
    a.runs = (const SkRuns*)plan->runs.ptr;
    a.problems = problems;
    a.segs = segs;
    a.scat = scat;
    a.A = (const char*)A;
    a.B = (const char*)B;
    a.C = (char*)C;
    a.ws = (char*)ws;
    a.counters = counters;
    a.nwarps = plan->nwarps;
    a.flags = flags;
    const int grid = (plan->nwarps + kSkWarps - 1) / kSkWarps;
    return plan->cplx ? launch_x<true>(plan->xt, plan->yt, grid, a, st) : launch_x<false>(plan->xt, plan->yt, grid, a, st);
}

void skinny_destroy(SkinnyPlan* plan) {
    if (!plan) return;
    plan->parts.release();
    plan->warps.release();
    plan->runs.release();
    delete plan;
}

void skinny_info(const SkinnyPlan* plan, int64_t* warps, int64_t* runs) {
    *warps = plan ? plan->nwarps : 0;
    *runs = plan ? plan->nruns : 0;
}

}  // namespace yb
