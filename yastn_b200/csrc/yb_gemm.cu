// Grouped variable-shape block GEMM on the FP64 tensor pipe (DMMA.8x8x4) for float64 and complex128.
//
// Replaces the per-sector loop of backend.dot (yastn/backend/_backend_torch_backwards.py:100-109: one cuBLAS
// launch + one slice copy per charge sector) and the per-pair loop of backend.transpose_dot_sum (:143-157)
// by ONE launch over all sectors.  FP64 has no tcgen05/TMEM path on sm_100a (tcgen05.mma kinds are
// f16/tf32/f8f6f4/i8/mx*); the FP64 tensor instruction is the warp-level DMMA.8x8x4 (every wider
// mma.sync f64 shape lowers to it), measured at 37.1 TFLOP/s on B200 (tools/microbench/fp64_pipes.cu),
// and operands are only 8-byte aligned in general, which rules out TMA tensor maps (16-byte base/stride).
// The kernel therefore uses a multi-stage cp.async pipeline into XOR-swizzled shared memory and
// register-blocked DMMA, one CTA per output tile, tiles ordered longest-first so the hardware block
// scheduler does the load balancing across all problems of the group.
//
// complex128 uses the 4M scheme on the same pipe: Cr += Ar*Br - Ai*Bi ; Ci += Ar*Bi + Ai*Br, with
// conjugation of either operand folded into the fragment loads (torch's lazy conj bit).
#include <algorithm>

#include "yb_common.h"

namespace yb {

constexpr int kGemmThreads = 128;   // 2 x 2 warps
constexpr int kStages = 3;
constexpr int kRowBytes = 128;      // bytes of one K-row (KC format): 16 doubles or 8 complex

struct GemmProblem {
    int32_t M, N;
    int32_t seg_begin, seg_end;
    int64_t offC, ldc;
};

struct GemmSegment {
    int64_t offA, offB;
    int64_t sAm, sAk, sBk, sBn;
    int32_t K;
    int32_t align;  // bit0: A rows 16B-aligned, bit1: B rows 16B-aligned (relative to a 16B-aligned base pointer)
};

struct GemmTile {
    int32_t prob, m0, n0, cfg;  // cfg 0: 64x128, 1: 64x64
};

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(uint32_t saddr, const void* g, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(saddr), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}
__device__ __forceinline__ double flip_sign(double v, uint32_t mask) {  // integer pipe, keeps the FP64 pipe for DMMA
    return __hiloint2double(__double2hiint(v) ^ (int)mask, __double2loint(v));
}
__device__ __forceinline__ double lds64(uint32_t saddr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ double2 lds128(uint32_t saddr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(saddr));
    return v;
}

// Shared-memory operand tile formats.  x is the non-contracted index (m for A, n for B).
//   KC: [x][kRowBytes]  — K contiguous (row-major A, or B^T/B^H);   chunk swizzle by x
//   XC: [k][BX * ES]    — x contiguous (row-major B, or A^T/A^H);   chunk swizzle by k
enum Layout : int { KC = 0, XC = 1 };

template <bool CPLX>
struct ElemTraits {
    static constexpr int ES = CPLX ? 16 : 8;            // element bytes
    static constexpr int BK = kRowBytes / ES;           // 16 (f64) or 8 (c128) contraction indices per stage
    static constexpr int KSTEPS = BK / 4;               // DMMA k-steps per stage
};

// Loader of one operand tile (BX non-contracted x BK contracted) for stage `sbase`.
// All 128 threads cooperate; 16-byte chunks, zero-filled outside the problem.
template <bool CPLX, int LAYOUT, int BX>
struct TileLoader {
    using TR = ElemTraits<CPLX>;
    static constexpr int ES = TR::ES, BK = TR::BK;
    static constexpr int CHUNKS = BX * BK * ES / 16;            // 16B chunks per tile
    static constexpr int PER_THREAD = CHUNKS / kGemmThreads;
    static constexpr int CPR = (LAYOUT == KC) ? (kRowBytes / 16) : (BX * ES / 16);  // chunks per smem row
    static constexpr int EPC = 16 / ES;                          // elements per chunk

    // ptr: operand base (already offset by the segment's off); sx / sk: strides of x and k; X,K: extents;
    // x0,k0: tile origin.  aligned16: every chunk start is 16B-aligned in global memory.
    __device__ __forceinline__ static void load(uint32_t sbase, const char* ptr, int64_t sx, int64_t sk, int X, int K,
                                                int x0, int k0, bool aligned16, int tid) {
#pragma unroll
        for (int i = 0; i < PER_THREAD; ++i) {
            const int idx = tid + i * kGemmThreads;
            const int row = idx / CPR, j = idx % CPR;
            int x, k, rem;       // element coordinates of the chunk start, elements remaining along the contiguous dim
            uint32_t soff;
            if (LAYOUT == KC) {
                x = x0 + row;
                k = k0 + j * EPC;
                rem = (x < X) ? (K - k) : 0;
                const int jj = CPLX ? (j ^ ((row & 1) << 2)) : (j ^ ((row & 3) << 1));
                soff = row * kRowBytes + jj * 16;
            } else {
                k = k0 + row;
                x = x0 + j * EPC;
                rem = (k < K) ? (X - x) : 0;
                const int jj = j ^ ((row & 3) << 1);
                soff = row * (BX * ES) + jj * 16;
            }
            rem = rem < 0 ? 0 : rem;
            const char* g = ptr + ((int64_t)x * sx + (int64_t)k * sk) * ES;
            if (rem == 0) g = ptr;  // keep the address valid; nothing is read
            if (CPLX || aligned16) {
                const int bytes = rem >= EPC ? 16 : rem * ES;
                cp_async16(sbase + soff, g, bytes);
            } else {
                cp_async8(sbase + soff, g, rem >= 1 ? 8 : 0);
                cp_async8(sbase + soff + 8, rem >= 2 ? g + 8 : ptr, rem >= 2 ? 8 : 0);
            }
        }
    }
};

// Fragment address of element (x, k) inside a tile (bytes from the tile base).
template <bool CPLX, int LAYOUT, int BX>
__device__ __forceinline__ uint32_t frag_offset(int x, int k) {
    constexpr int ES = ElemTraits<CPLX>::ES;
    if (LAYOUT == KC) {
        const int byte = k * ES;
        const int j = byte >> 4;
        const int jj = CPLX ? (j ^ ((x & 1) << 2)) : (j ^ ((x & 3) << 1));
        return x * kRowBytes + jj * 16 + (byte & 15);
    } else {
        const int byte = x * ES;
        const int j = byte >> 4;
        const int jj = j ^ ((k & 3) << 1);
        return k * (BX * ES) + jj * 16 + (byte & 15);
    }
}

template <bool CPLX, int AL, int BL, int BM, int BN>
struct TileKernel {
    using TR = ElemTraits<CPLX>;
    static constexpr int ES = TR::ES, BK = TR::BK, KSTEPS = TR::KSTEPS;
    static constexpr int WM = BM / 2, WN = BN / 2;     // warp tile
    static constexpr int MT = WM / 8, NT = WN / 8;     // 8x8 DMMA tiles per warp
    static constexpr int A_BYTES = BM * kRowBytes, B_BYTES = BN * kRowBytes;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM_BYTES = STAGE_BYTES * kStages;

    __device__ static void run(const GemmProblem& P, const GemmSegment* __restrict__ segs, const GemmTile& T,
                               const char* __restrict__ A, const char* __restrict__ B, char* __restrict__ C,
                               int flags, bool base_aligned, uint32_t smem) {
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int wm0 = (warp >> 1) * WM, wn0 = (warp & 1) * WN;
        const int lx = lane >> 2, lk = lane & 3;

        // accumulators: real part (and imaginary part for complex)
        double acc[MT][NT][CPLX ? 4 : 2];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int c = 0; c < (CPLX ? 4 : 2); ++c) acc[i][j][c] = 0.0;

        // iteration space: all K-chunks of all segments of this problem
        int total_iters = 0;
        for (int s = P.seg_begin; s < P.seg_end; ++s) total_iters += (segs[s].K + BK - 1) / BK;

        // loader state
        int ld_seg = P.seg_begin, ld_k0 = 0, ld_it = 0;
        auto issue_load = [&](int stage) {
            if (ld_it < total_iters) {
                while (segs[ld_seg].K <= ld_k0) {  // also skips K == 0 segments
                    ++ld_seg;
                    ld_k0 = 0;
                }
                const GemmSegment& S = segs[ld_seg];
                const uint32_t sa = smem + stage * STAGE_BYTES, sb = sa + A_BYTES;
                TileLoader<CPLX, AL, BM>::load(sa, A + S.offA * ES, S.sAm, S.sAk, P.M, S.K, T.m0, ld_k0,
                                               base_aligned && (S.align & 1), tid);
                TileLoader<CPLX, BL, BN>::load(sb, B + S.offB * ES, S.sBn, S.sBk, P.N, S.K, T.n0, ld_k0,
                                               base_aligned && (S.align & 2), tid);
                ld_k0 += BK;
                ++ld_it;
            }
            cp_async_commit();
        };

#pragma unroll
        for (int s = 0; s < kStages - 1; ++s) issue_load(s);

        const uint32_t sgnA = (CPLX && (flags & YB_GEMM_CONJ_A)) ? 0x80000000u : 0u;
        const uint32_t sgnB = (CPLX && (flags & YB_GEMM_CONJ_B)) ? 0x80000000u : 0u;

        for (int it = 0; it < total_iters; ++it) {
            cp_async_wait<kStages - 2>();
            __syncthreads();
            issue_load((it + kStages - 1) % kStages);
            const uint32_t sa = smem + (it % kStages) * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                if constexpr (!CPLX) {
                    double af[MT], bf[NT];
#pragma unroll
                    for (int i = 0; i < MT; ++i) af[i] = lds64(sa + frag_offset<CPLX, AL, BM>(wm0 + i * 8 + lx, ks * 4 + lk));
#pragma unroll
                    for (int j = 0; j < NT; ++j) bf[j] = lds64(sb + frag_offset<CPLX, BL, BN>(wn0 + j * 8 + lx, ks * 4 + lk));
#pragma unroll
                    for (int i = 0; i < MT; ++i)
#pragma unroll
                        for (int j = 0; j < NT; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                } else {
                    double2 af[MT], bf[NT];
                    double naf[MT];
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        af[i] = lds128(sa + frag_offset<CPLX, AL, BM>(wm0 + i * 8 + lx, ks * 4 + lk));
                        af[i].y = flip_sign(af[i].y, sgnA);
                        naf[i] = flip_sign(af[i].y, 0x80000000u);
                    }
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        bf[j] = lds128(sb + frag_offset<CPLX, BL, BN>(wn0 + j * 8 + lx, ks * 4 + lk));
                        bf[j].y = flip_sign(bf[j].y, sgnB);
                    }
#pragma unroll
                    for (int i = 0; i < MT; ++i)
#pragma unroll
                        for (int j = 0; j < NT; ++j) {
                            dmma(acc[i][j][0], acc[i][j][1], af[i].x, bf[j].x);
                            dmma(acc[i][j][0], acc[i][j][1], naf[i], bf[j].y);
                            dmma(acc[i][j][2], acc[i][j][3], af[i].x, bf[j].y);
                            dmma(acc[i][j][2], acc[i][j][3], af[i].y, bf[j].x);
                        }
                }
            }
        }
        cp_async_wait<0>();

        // epilogue: lane holds C[row][col], C[row][col+1] of every 8x8 tile
        const int64_t ldc = P.ldc;
        if constexpr (!CPLX) {
            double* Cp = reinterpret_cast<double*>(C) + P.offC;
            const bool vec = base_aligned && ((P.offC & 1) == 0) && ((ldc & 1) == 0);
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int row = T.m0 + wm0 + i * 8 + lx;
                if (row < P.M) {
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const int col = T.n0 + wn0 + j * 8 + 2 * lk;
                        double* p = Cp + row * ldc + col;
                        if (vec && col + 1 < P.N) {
                            *reinterpret_cast<double2*>(p) = make_double2(acc[i][j][0], acc[i][j][1]);
                        } else {
                            if (col < P.N) p[0] = acc[i][j][0];
                            if (col + 1 < P.N) p[1] = acc[i][j][1];
                        }
                    }
                }
            }
        } else {
            double2* Cp = reinterpret_cast<double2*>(C) + P.offC;
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int row = T.m0 + wm0 + i * 8 + lx;
                if (row < P.M) {
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const int col = T.n0 + wn0 + j * 8 + 2 * lk;
                        double2* p = Cp + row * ldc + col;
                        if (col < P.N) p[0] = make_double2(acc[i][j][0], acc[i][j][2]);
                        if (col + 1 < P.N) p[1] = make_double2(acc[i][j][1], acc[i][j][3]);
                    }
                }
            }
        }
    }
};

template <bool CPLX, int AL, int BL>
struct GroupKernel {
    using Big = TileKernel<CPLX, AL, BL, 64, CPLX ? 64 : 128>;
    using Small = TileKernel<CPLX, AL, BL, CPLX ? 32 : 64, CPLX ? 32 : 64>;
    static constexpr int SMEM_BYTES = Big::SMEM_BYTES > Small::SMEM_BYTES ? Big::SMEM_BYTES : Small::SMEM_BYTES;
};

template <bool CPLX, int AL, int BL>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_kernel(const GemmProblem* __restrict__ problems, const GemmSegment* __restrict__ segs,
            const GemmTile* __restrict__ tiles, const char* __restrict__ A, const char* __restrict__ B,
            char* __restrict__ C, int flags, int base_aligned) {
    extern __shared__ __align__(128) char smem_raw[];
    const uint32_t smem = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const GemmTile T = tiles[blockIdx.x];
    const GemmProblem P = problems[T.prob];
    using G = GroupKernel<CPLX, AL, BL>;
    if (T.cfg == 0)
        G::Big::run(P, segs, T, A, B, C, flags, base_aligned != 0, smem);
    else
        G::Small::run(P, segs, T, A, B, C, flags, base_aligned != 0, smem);
}

}  // namespace yb

using namespace yb;

struct yb_gemm_plan {
    int dtype = 0, device = 0;
    int al = KC, bl = XC;
    int ntiles = 0;
    int64_t macs = 0, nbig = 0, nsmall = 0;
    DeviceTable problems, segments, tiles;
};

namespace {

template <bool CPLX, int AL, int BL>
int launch(const yb_gemm_plan* p, const void* A, const void* B, void* C, int flags, int base_aligned, cudaStream_t st) {
    using G = GroupKernel<CPLX, AL, BL>;
    static bool configured[64] = {false};
    if (p->device >= 0 && p->device < 64 && !configured[p->device]) {
        YB_CUDA(cudaFuncSetAttribute(gemm_kernel<CPLX, AL, BL>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
        configured[p->device] = true;
    }
    gemm_kernel<CPLX, AL, BL><<<p->ntiles, kGemmThreads, G::SMEM_BYTES, st>>>(
        (const GemmProblem*)p->problems.ptr, (const GemmSegment*)p->segments.ptr, (const GemmTile*)p->tiles.ptr,
        (const char*)A, (const char*)B, (char*)C, flags, base_aligned);
    YB_CUDA(cudaGetLastError());
    return kOk;
}

template <bool CPLX>
int dispatch_layout(const yb_gemm_plan* p, const void* A, const void* B, void* C, int flags, int ba, cudaStream_t st) {
    if (p->al == KC && p->bl == XC) return launch<CPLX, KC, XC>(p, A, B, C, flags, ba, st);
    if (p->al == KC && p->bl == KC) return launch<CPLX, KC, KC>(p, A, B, C, flags, ba, st);
    if (p->al == XC && p->bl == XC) return launch<CPLX, XC, XC>(p, A, B, C, flags, ba, st);
    return launch<CPLX, XC, KC>(p, A, B, C, flags, ba, st);
}

}  // namespace

extern "C" int yb_gemm_plan_create(const int64_t* problems, int64_t nprob, const int64_t* segments, int64_t nseg,
                                   int dtype, int device, yb_gemm_plan** out) {
    if (!out) return fail(kErrArg, "yb_gemm_plan_create: out is null");
    *out = nullptr;
    if (dtype != YB_F64 && dtype != YB_C128) return fail(kErrUnsupported, "yb_gemm_plan_create: dtype %d", dtype);
    if (nprob < 0 || nseg < 0 || (nprob > 0 && !problems) || (nseg > 0 && !segments))
        return fail(kErrArg, "yb_gemm_plan_create: bad tables");
    const bool cplx = dtype == YB_C128;

    // operand layouts: decided by which stride is 1; must be consistent over the plan
    int al = KC, bl = XC;
    bool a_ok[2] = {true, true}, b_ok[2] = {true, true};
    std::vector<GemmSegment> hs((size_t)nseg);
    for (int64_t s = 0; s < nseg; ++s) {
        const int64_t* q = segments + s * 7;
        GemmSegment& g = hs[(size_t)s];
        if (q[0] < 0 || q[0] > 0x7fffffff) return fail(kErrArg, "yb_gemm_plan_create: segment %lld K out of range", (long long)s);
        g.K = (int32_t)q[0];
        g.offA = q[1];
        g.sAm = q[2];
        g.sAk = q[3];
        g.offB = q[4];
        g.sBk = q[5];
        g.sBn = q[6];
        g.align = 0;
    }
    std::vector<GemmProblem> hp((size_t)nprob);
    for (int64_t i = 0; i < nprob; ++i) {
        const int64_t* q = problems + i * 6;
        GemmProblem& g = hp[(size_t)i];
        if (q[0] < 0 || q[1] < 0 || q[0] > 0x7fffffff || q[1] > 0x7fffffff) return fail(kErrArg, "yb_gemm_plan_create: problem %lld shape out of range", (long long)i);
        g.M = (int32_t)q[0];
        g.N = (int32_t)q[1];
        g.offC = q[2];
        g.ldc = q[3];
        if (q[4] < 0 || q[5] < q[4] || q[5] > nseg) return fail(kErrArg, "yb_gemm_plan_create: problem %lld segment range", (long long)i);
        g.seg_begin = (int32_t)q[4];
        g.seg_end = (int32_t)q[5];
        // layout feasibility: a layout is usable when its contiguous index has unit stride (or extent <= 1)
        for (int s = g.seg_begin; s < g.seg_end; ++s) {
            const GemmSegment& sg = hs[(size_t)s];
            if (g.M == 0 || g.N == 0 || sg.K == 0) continue;
            a_ok[KC] = a_ok[KC] && (sg.sAk == 1 || sg.K <= 1);
            a_ok[XC] = a_ok[XC] && (sg.sAm == 1 || g.M <= 1);
            b_ok[XC] = b_ok[XC] && (sg.sBn == 1 || g.N <= 1);
            b_ok[KC] = b_ok[KC] && (sg.sBk == 1 || sg.K <= 1);
        }
    }
    if (!a_ok[KC] && !a_ok[XC]) return fail(kErrUnsupported, "yb_gemm_plan_create: operand A has no common unit-stride index");
    if (!b_ok[KC] && !b_ok[XC]) return fail(kErrUnsupported, "yb_gemm_plan_create: operand B has no common unit-stride index");
    al = a_ok[KC] ? KC : XC;
    bl = b_ok[XC] ? XC : KC;
    // 16-byte alignment of every row start (f64 only; complex elements are 16 bytes)
    for (auto& sg : hs) {
        const int64_t a_ld = (al == KC) ? sg.sAm : sg.sAk;
        const int64_t b_ld = (bl == XC) ? sg.sBk : sg.sBn;
        if (((sg.offA | a_ld) & 1) == 0) sg.align |= 1;
        if (((sg.offB | b_ld) & 1) == 0) sg.align |= 2;
    }

    // tiles, longest first (work ~ sum of K over the problem's segments)
    struct TW {
        GemmTile t;
        int64_t work;
    };
    std::vector<TW> tw;
    int64_t macs = 0, nbig = 0, nsmall = 0;
    const int bigM = 64, bigN = cplx ? 64 : 128, smM = cplx ? 32 : 64, smN = cplx ? 32 : 64;
    for (int64_t i = 0; i < nprob; ++i) {
        const GemmProblem& g = hp[(size_t)i];
        if (g.M == 0 || g.N == 0) continue;
        int64_t ksum = 0;
        for (int s = g.seg_begin; s < g.seg_end; ++s) ksum += hs[(size_t)s].K;
        macs += (int64_t)g.M * g.N * ksum;
        const bool big = g.N > smN && g.M > smM / 2;
        const int bm = big ? bigM : smM, bn = big ? bigN : smN;
        for (int m0 = 0; m0 < g.M; m0 += bm)
            for (int n0 = 0; n0 < g.N; n0 += bn) {
                TW x;
                x.t = {(int32_t)i, m0, n0, big ? 0 : 1};
                x.work = ksum * (big ? 2 : 1);
                tw.push_back(x);
                (big ? nbig : nsmall)++;
            }
    }
    std::stable_sort(tw.begin(), tw.end(), [](const TW& a, const TW& b) { return a.work > b.work; });
    std::vector<GemmTile> ht(tw.size());
    for (size_t i = 0; i < tw.size(); ++i) ht[i] = tw[i].t;

    yb_gemm_plan* plan = new yb_gemm_plan();
    plan->dtype = dtype;
    plan->device = device;
    plan->al = al;
    plan->bl = bl;
    plan->ntiles = (int)ht.size();
    plan->macs = macs;
    plan->nbig = nbig;
    plan->nsmall = nsmall;
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = kOk;
    if (cudaSetDevice(device) != cudaSuccess) rc = fail(kErrCuda, "yb_gemm_plan_create: cudaSetDevice(%d) failed", device);
    if (rc == kOk) rc = plan->problems.upload(hp.data(), hp.size() * sizeof(GemmProblem));
    if (rc == kOk) rc = plan->segments.upload(hs.data(), hs.size() * sizeof(GemmSegment));
    if (rc == kOk) rc = plan->tiles.upload(ht.data(), ht.size() * sizeof(GemmTile));
    cudaSetDevice(prev);
    if (rc != kOk) {
        yb_gemm_plan_destroy(plan);
        return rc;
    }
    *out = plan;
    return kOk;
}

extern "C" int yb_gemm_plan_info(const yb_gemm_plan* plan, int64_t info[4]) {
    if (!plan || !info) return fail(kErrArg, "yb_gemm_plan_info: null argument");
    info[0] = plan->ntiles;
    info[1] = plan->macs;
    info[2] = plan->nbig;
    info[3] = plan->nsmall;
    return kOk;
}

extern "C" int yb_gemm_run(const yb_gemm_plan* plan, const void* A, const void* B, void* C, int flags, void* stream) {
    if (!plan) return fail(kErrArg, "yb_gemm_run: plan is null");
    if (plan->ntiles == 0) return kOk;
    if (!A || !B || !C) return fail(kErrArg, "yb_gemm_run: null data pointer");
    const int ba = (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) == 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->dtype == YB_C128) {
        if (!ba) return fail(kErrArg, "yb_gemm_run: complex128 operands must be 16-byte aligned");
        return dispatch_layout<true>(plan, A, B, C, flags, ba, st);
    }
    return dispatch_layout<false>(plan, A, B, C, flags, ba, st);
}

extern "C" void yb_gemm_plan_destroy(yb_gemm_plan* plan) {
    if (!plan) return;
    plan->problems.release();
    plan->segments.release();
    plan->tiles.release();
    delete plan;
}
