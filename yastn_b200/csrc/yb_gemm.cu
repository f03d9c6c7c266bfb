// Grouped variable-shape block GEMM on the FP64 tensor pipe (DMMA.8x8x4) for float64 and complex128.
//
// Replaces the per-sector loop of backend.dot (yastn/backend/_backend_torch_backwards.py:100-109: one cuBLAS
// launch + one slice copy per charge sector), the per-pair loop of backend.transpose_dot_sum (:143-157) and —
// through the scatter epilogue — the per-block loop of backend.unmerge (:397-408) by ONE persistent launch.
//
// FP64 has no tcgen05/TMEM path on sm_100a (tcgen05.mma kinds are f16/tf32/f8f6f4/i8/mx*); the FP64 tensor
// instruction is the warp-level DMMA.8x8x4 (every wider mma.sync f64 shape lowers to it), measured at
// 37.1 TFLOP/s on B200 (tools/microbench/fp64_pipes.cu).  Block offsets and leading dimensions of YASTN's 1-D
// storage are arbitrary element counts, i.e. float64 operands are only 8-byte aligned in general, which rules
// out TMA (16-byte global alignment); tiles are staged with a multi-stage cp.async pipeline into XOR-swizzled
// shared memory (16-byte copies when a sector happens to be aligned, 8-byte copies otherwise).
//
// Scheduling is stream-K: the k-iterations of all output tiles of all sectors form one weighted work line that
// the host cuts into equal shares, one per resident CTA (grid = SMs x occupancy).  A CTA that starts in the
// middle of a tile writes its partial accumulators to a per-CTA workspace slot and raises a flag; the CTA that
// started the tile adds the partials in fixed CTA order (deterministic) and runs the epilogue.  This keeps
// every SM busy for few large sectors, many tiny sectors and the huge-K / tiny-output shape alike.
//
// complex128 uses the 4M scheme on the same pipe: Cr += Ar*Br - Ai*Bi ; Ci += Ar*Bi + Ai*Br, with
// conjugation of either operand folded into the fragment loads (torch's lazy conj bit).
#include <algorithm>
#include <cmath>
#include <type_traits>

#include "yb_gemm_types.h"

namespace yb {

constexpr int kGemmThreads = 128;   // 2 x 2 warps
constexpr int kStages = 4;
constexpr int kRowBytes = 128;      // bytes of one K-row (KC format): 16 doubles or 8 complex
constexpr int kTilesInFlight = 296;   // resident CTAs of a full launch (148 SMs x 2): sizes the L2 tile bands
constexpr int kWsDoubles = 64 * kGemmThreads;   // partial-accumulator slot per CTA (64 doubles per thread)

struct GemmTile {
    int32_t prob, m0, n0, cfg;   // cfg 0: big tile, 1: small tile
    int32_t iters;               // k-iterations of the whole tile (sum over segments of ceil(K / BK))
    int32_t pad_[3];
};

// Share of one CTA: tiles [tile_begin, tile_end], starting at k-iteration it_begin of the first tile and ending
// before k-iteration it_end of the last one.
struct CtaRange {
    int32_t tile_begin, it_begin, tile_end, it_end;
};

struct GemmArgs {
    const GemmProblem* problems;
    const GemmSegment* segs;
    const GemmTile* tiles;
    const CtaRange* ranges;
    const ScatterInfo* scat;
    const int2* rowinfo;
    const int4* colinfo;
    const int64_t* dstpool;
    double* ws;
    int* sync_flags;
    const char* A;
    const char* B;
    char* C;
    int flags;
    int base_aligned;
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
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
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

// Loader of one operand tile (BX non-contracted x BK contracted).  All 128 threads cooperate, 16-byte chunks,
// zero-filled outside the problem.  Thread t owns chunks t, t+128, ...: they share the coordinate along the
// contiguous dim and are RSTEP apart along the strided one, so the global address of chunk i is
// g0 + i * cstep and a k-iteration only advances g0 — no per-chunk index arithmetic in the main loop.
template <bool CPLX, int LAYOUT, int BX>
struct TileLoader {
    using TR = ElemTraits<CPLX>;
    static constexpr int ES = TR::ES, BK = TR::BK;
    static constexpr int CHUNKS = BX * BK * ES / 16;            // 16B chunks per tile
    static constexpr int NCH = CHUNKS / kGemmThreads;           // chunks per thread
    static constexpr int CPR = (LAYOUT == KC) ? (kRowBytes / 16) : (BX * ES / 16);  // chunks per smem row
    static constexpr int EPC = 16 / ES;                          // elements per chunk
    static constexpr int RSTEP = kGemmThreads / CPR;             // smem rows between consecutive chunks of a thread
    static constexpr int ROWB = (LAYOUT == KC) ? kRowBytes : BX * ES;   // bytes of one smem row
    static_assert(CPR <= kGemmThreads && kGemmThreads % CPR == 0 && NCH >= 1, "tile / thread mismatch");
    static_assert(LAYOUT == KC || RSTEP % 2 == 0, "XC swizzle update assumes an even row step");

    const char* g0;     // global address of chunk 0 at the current k-iteration
    uint32_t cstep;     // bytes between consecutive chunks of this thread
    int64_t kadv;       // bytes per k-iteration
    uint32_t soff0;     // smem offset of chunk 0 inside a stage
    int lim;            // KC: bit mask of chunks whose row is inside the problem; XC: bytes (0..16) of every chunk
    int c0;             // KC: first contracted index of this thread's chunks (j * EPC); XC: strided row r0
    bool al16;

    // base: operand base already offset to the segment; (sx, sk): strides of x and k; X: extent of x;
    // x0: tile origin; k0: first contracted index of the first iteration
    __device__ __forceinline__ void init(const char* base, int64_t sx, int64_t sk, int X, int x0, int k0, bool aligned16, int tid) {
        const int r0 = tid / CPR, j = tid % CPR;
        al16 = CPLX || aligned16;
        if (LAYOUT == KC) {
            const int jj = CPLX ? (j ^ ((r0 & 1) << 2)) : (j ^ ((r0 & 3) << 1));
            soff0 = r0 * ROWB + jj * 16;
            g0 = base + ((int64_t)(x0 + r0) * sx + (int64_t)(k0 + j * EPC) * sk) * ES;
            cstep = (uint32_t)(RSTEP * sx * ES);
            kadv = (int64_t)BK * sk * ES;
            lim = 0;
#pragma unroll
            for (int i = 0; i < NCH; ++i) lim |= (x0 + r0 + i * RSTEP < X) ? (1 << i) : 0;
            c0 = j * EPC;
        } else {
            const int jj = j ^ ((r0 & 3) << 1);
            soff0 = r0 * ROWB + jj * 16;
            g0 = base + ((int64_t)(x0 + j * EPC) * sx + (int64_t)(k0 + r0) * sk) * ES;
            cstep = (uint32_t)(RSTEP * sk * ES);
            kadv = (int64_t)BK * sk * ES;
            int rem = (X - (x0 + j * EPC)) * ES;
            lim = rem < 0 ? 0 : (rem > 16 ? 16 : rem);
            c0 = r0;
        }
    }

    // kleft = K - k0 of this iteration (>= 1)
    __device__ __forceinline__ void issue(uint32_t sbase, int kleft) {
        int cb = 16;
        if (LAYOUT == KC) {
            cb = (kleft - c0) * ES;
            cb = cb < 0 ? 0 : (cb > 16 ? 16 : cb);
        }
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            int bytes;
            uint32_t so;
            if (LAYOUT == KC) {
                bytes = ((lim >> i) & 1) ? cb : 0;
                so = soff0 + i * RSTEP * ROWB;
            } else {
                bytes = (c0 + i * RSTEP < kleft) ? lim : 0;
                // the swizzle term ((row & 3) << 1) changes with i when RSTEP is not a multiple of 4
                so = (soff0 ^ ((((i * RSTEP) & 3) << 1) * 16)) + i * RSTEP * ROWB;
            }
            const char* g = g0 + (uint64_t)i * cstep;
            if (al16) {
                cp_async16(sbase + so, g, bytes);
            } else {
                cp_async8(sbase + so, g, bytes >= 8 ? 8 : 0);
                cp_async8(sbase + so + 8, g + 8, bytes >= 16 ? 8 : 0);
            }
        }
        g0 += kadv;
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
    static constexpr int BM_ = BM, BN_ = BN;
    static constexpr int WM = BM / 2, WN = BN / 2;     // warp tile
    static constexpr int MT = WM / 8, NT = WN / 8;     // 8x8 DMMA tiles per warp
    static constexpr int NACC = MT * NT * (CPLX ? 4 : 2);
    static constexpr int A_BYTES = BM * kRowBytes, B_BYTES = BN * kRowBytes;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM_BYTES = STAGE_BYTES * kStages;
    static_assert(NACC <= 64, "workspace slot too small");

    // Run k-iterations [it_begin, it_end) of tile T; handles the stream-K fix-up and the epilogue.
    __device__ static void run(const GemmArgs& g, const GemmProblem& P, const GemmTile& T, int tile_index, int it_begin, int it_end,
                               uint32_t smem) {
        const GemmSegment* __restrict__ segs = g.segs;
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int wm0 = (warp >> 1) * WM, wn0 = (warp & 1) * WN;
        const int lx = lane >> 2, lk = lane & 3;
        const bool base_aligned = g.base_aligned != 0;

        double acc[MT][NT][CPLX ? 4 : 2];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int c = 0; c < (CPLX ? 4 : 2); ++c) acc[i][j][c] = 0.0;

        const int n_iters = min(it_end, T.iters) - it_begin;   // T.iters may be 0 (all K == 0): plain zero block

        // loader state: position (segment, k0) of the next iteration to issue
        TileLoader<CPLX, AL, BM> la;
        TileLoader<CPLX, BL, BN> lb;
        int ld_seg = P.seg_begin, ld_k0 = 0, ld_K = 0, ld_it = 0;
        bool fresh = true;
        if (n_iters > 0) {  // locate iteration it_begin
            int skip = it_begin;
            for (;; ++ld_seg) {
                const int K = segs[ld_seg].K;
                const int n = (K + BK - 1) / BK;
                if (skip < n) {
                    ld_k0 = skip * BK;
                    ld_K = K;
                    break;
                }
                skip -= n;
            }
        }
        auto issue_load = [&](int stage) {
            if (ld_it < n_iters) {
                if (ld_k0 >= ld_K) {  // next segment with K > 0
                    do {
                        ++ld_seg;
                        ld_K = segs[ld_seg].K;
                    } while (ld_K <= 0);
                    ld_k0 = 0;
                    fresh = true;
                }
                if (fresh) {
                    const GemmSegment& S = segs[ld_seg];
                    la.init(g.A + S.offA * ES, S.sAm, S.sAk, P.M, T.m0, ld_k0, base_aligned && (S.align & 1), tid);
                    lb.init(g.B + S.offB * ES, S.sBn, S.sBk, P.N, T.n0, ld_k0, base_aligned && (S.align & 2), tid);
                    fresh = false;
                }
                const uint32_t sa = smem + stage * STAGE_BYTES, sb = sa + A_BYTES;
                la.issue(sa, ld_K - ld_k0);
                lb.issue(sb, ld_K - ld_k0);
                ld_k0 += BK;
                ++ld_it;
            }
            cp_async_commit();
        };

#pragma unroll
        for (int s = 0; s < kStages - 1; ++s) issue_load(s);

        const uint32_t sgnA = (CPLX && (g.flags & YB_GEMM_CONJ_A)) ? 0x80000000u : 0u;
        const uint32_t sgnB = (CPLX && (g.flags & YB_GEMM_CONJ_B)) ? 0x80000000u : 0u;

        // per-thread fragment offsets inside a stage (k-step 0); later k-steps add a compile-time constant
        uint32_t aoff[MT], boff[NT];
#pragma unroll
        for (int i = 0; i < MT; ++i) aoff[i] = frag_offset<CPLX, AL, BM>(wm0 + i * 8 + lx, lk);
#pragma unroll
        for (int j = 0; j < NT; ++j) boff[j] = A_BYTES + frag_offset<CPLX, BL, BN>(wn0 + j * 8 + lx, lk);

        for (int it = 0; it < n_iters; ++it) {
            cp_async_wait<kStages - 2>();
            __syncthreads();
            issue_load((it + kStages - 1) % kStages);
            const uint32_t sa = smem + (it % kStages) * STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                // the swizzle only depends on the low bits of x (KC) or of k (XC); k-step ks moves k by 4*ks:
                //   KC: byte offset += 4*ks*ES inside the 128-byte row -> chunk index changes by a constant XOR-free add
                //   XC: row += 4*ks rows of BX*ES bytes, swizzle term (k & 3) unchanged
                if constexpr (!CPLX) {
                    double af[MT], bf[NT];
#pragma unroll
                    for (int i = 0; i < MT; ++i)
                        af[i] = lds64(sa + (AL == KC ? (aoff[i] ^ (ks * 32)) : (aoff[i] + ks * 4 * BM * ES)));
#pragma unroll
                    for (int j = 0; j < NT; ++j)
                        bf[j] = lds64(sa + (BL == KC ? (boff[j] ^ (ks * 32)) : (boff[j] + ks * 4 * BN * ES)));
#pragma unroll
                    for (int i = 0; i < MT; ++i)
#pragma unroll
                        for (int j = 0; j < NT; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                } else {
                    double2 af[MT], bf[NT];
                    double naf[MT];
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        af[i] = lds128(sa + (AL == KC ? (aoff[i] ^ (ks * 64)) : (aoff[i] + ks * 4 * BM * ES)));
                        af[i].y = flip_sign(af[i].y, sgnA);
                        naf[i] = flip_sign(af[i].y, 0x80000000u);
                    }
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        bf[j] = lds128(sa + (BL == KC ? (boff[j] ^ (ks * 64)) : (boff[j] + ks * 4 * BN * ES)));
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
        __syncthreads();   // all warps are done with the stages before the next tile's prologue overwrites them

        // ---- stream-K fix-up ---------------------------------------------------------------------
        const bool starts = it_begin == 0, ends = it_end >= T.iters;
        if (!starts) {
            // contributor: publish the partial tile in this CTA's slot and leave
            double* slot = g.ws + (size_t)blockIdx.x * kWsDoubles + tid;
            int r = 0;
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int c = 0; c < (CPLX ? 4 : 2); ++c) __stcg(slot + (r++) * kGemmThreads, acc[i][j][c]);
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release(g.sync_flags + blockIdx.x, 1);
            return;
        }
        if (!ends) {
            // finisher: add the partials of the following CTAs in CTA order
            for (int d = blockIdx.x + 1;; ++d) {
                const CtaRange Rd = g.ranges[d];
                if (tid == 0) {
                    while (ld_acquire(g.sync_flags + d) == 0) __nanosleep(64);
                    g.sync_flags[d] = 0;   // flags rest at 0 between launches (graph-replay safe)
                }
                __syncthreads();
                const double* slot = g.ws + (size_t)d * kWsDoubles + tid;
                int r = 0;
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j)
#pragma unroll
                        for (int c = 0; c < (CPLX ? 4 : 2); ++c) acc[i][j][c] += __ldcg(slot + (r++) * kGemmThreads);
                if (Rd.tile_end > tile_index || Rd.it_end >= T.iters) break;
            }
        }

        // ---- epilogue: lane holds C[row][col], C[row][col+1] of every 8x8 tile ---------------------
        using T2 = typename std::conditional<CPLX, double2, double>::type;
        auto value = [&](int i, int j, int e) -> T2 {
            if constexpr (CPLX) return make_double2(acc[i][j][e], acc[i][j][2 + e]);
            else return acc[i][j][e];
        };
        if (P.scat < 0) {
            T2* Cp = reinterpret_cast<T2*>(g.C) + P.offC;
            const bool vec = !CPLX && base_aligned && ((P.offC & 1) == 0) && ((P.ldc & 1) == 0);
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int row = T.m0 + wm0 + i * 8 + lx;
                if (row < P.M) {
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const int col = T.n0 + wn0 + j * 8 + 2 * lk;
                        T2* p = Cp + (int64_t)row * P.ldc + col;
                        if constexpr (!CPLX) {
                            if (vec && col + 1 < P.N) {
                                *reinterpret_cast<double2*>(p) = make_double2(acc[i][j][0], acc[i][j][1]);
                                continue;
                            }
                        }
                        if (col < P.N) p[0] = value(i, j, 0);
                        if (col + 1 < P.N) p[1] = value(i, j, 1);
                    }
                }
            }
        } else {
            // fused unmerge: the row lookups are shared by the NT column pairs of a row and the column lookups by the MT rows
            const ScatterInfo S = g.scat[P.scat];
            T2* out = reinterpret_cast<T2*>(g.C);
            int2 ri[MT];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int row = T.m0 + wm0 + i * 8 + lx;
                ri[i] = row < P.M ? g.rowinfo[S.row_off + row] : make_int2(-1, 0);
            }
            const int64_t* dst = g.dstpool + S.dst_off;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int col = T.n0 + wn0 + j * 8 + 2 * lk;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (col + e < P.N) {
                        const int4 ci = g.colinfo[S.col_off + col + e];
#pragma unroll
                        for (int i = 0; i < MT; ++i)
                            if (ri[i].x >= 0) out[dst[(int64_t)ri[i].x * S.ncs + ci.x] + (int64_t)ri[i].y * ci.z + ci.y] = value(i, j, e);
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
__global__ void __launch_bounds__(kGemmThreads, 2) gemm_kernel(const GemmArgs g) {
    extern __shared__ __align__(128) char smem_raw[];
    const uint32_t smem = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const CtaRange R = g.ranges[blockIdx.x];
    using G = GroupKernel<CPLX, AL, BL>;
    for (int t = R.tile_begin; t <= R.tile_end; ++t) {
        const GemmTile T = g.tiles[t];
        const GemmProblem P = g.problems[T.prob];
        const int ib = (t == R.tile_begin) ? R.it_begin : 0;
        const int ie = (t == R.tile_end) ? R.it_end : max(T.iters, 1);
        if (T.cfg == 0)
            G::Big::run(g, P, T, t, ib, ie, smem);
        else
            G::Small::run(g, P, T, t, ib, ie, smem);
    }
}

// =====================================================================================================================
// Warp-specialised variant: a producer warpgroup streams operand tiles through the shared-memory ring, a consumer warpgroup
// (the same 2 x 2 warps and fragment layout as gemm_kernel) issues the DMMAs.
//
// Why (ncu on gemm_kernel, profiles/ncu_r02_gemm_u1u1.txt): every k-iteration the math warps also issue 12 cp.async each
// with their address arithmetic and meet at a CTA barrier; a tile shorter than ~16 k-iterations (K <= 163 in U1xU1 sectors)
// pays the full pipeline fill and the epilogue with an idle tensor pipe (DMMA 71 % busy while active, 64 % of the launch),
// and at 254 registers the kernel has no room to skip the padded 8 x 8 slices of edge tiles.  Here
//   * the loaders live in the producer warps only (setmaxnreg moves their registers to the consumers: 40 / 216 per thread at
//     256 threads and 2 CTAs per SM),
//   * stages are handed over through mbarriers (full: cp.async.mbarrier.arrive of the 128 producer threads; empty: one arrive
//     per consumer warp), so there is no CTA-wide barrier in the main loop and consumer warps drift up to a ring apart,
//   * the producer runs ahead ACROSS tile boundaries: the first stages of the next tile are in flight while the consumers
//     finish the current tile and write its epilogue.
// =====================================================================================================================
constexpr int kWsThreads = 256;          // warpgroup 0: consumers (math), warpgroup 1: producers (loads)
constexpr int kFlagNoEdgeSkip = 0x100;   // internal run flag (YB_GEMM_NOSKIP=1): multiply the padding of edge tiles too (A/B measurements)
constexpr int kWsProducerRegs = 40;
constexpr int kWsConsumerRegs = 216;

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cp_async(uint32_t bar) {   // arrives once all cp.async of this thread have landed
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void consumer_bar_sync() { asm volatile("bar.sync 1, 128;\n" ::: "memory"); }

template <bool CPLX, int AL, int BL>
struct WsShape {
    using G = GroupKernel<CPLX, AL, BL>;
    static constexpr int STAGE_BYTES = G::Big::STAGE_BYTES > G::Small::STAGE_BYTES ? G::Big::STAGE_BYTES : G::Small::STAGE_BYTES;
    static constexpr int BAR_BYTES = 2 * kStages * 8;
    static constexpr int SMEM_BYTES = STAGE_BYTES * kStages + BAR_BYTES;
};

// iterations [ib, ie) of tile T that actually exist
__device__ __forceinline__ int ws_tile_iters(const GemmTile& T, int ib, int ie) {
    const int n = min(ie, T.iters) - ib;
    return n > 0 ? n : 0;
}

// ---- producer: all 128 threads of warpgroup 1 -------------------------------------------------------------------------
template <bool CPLX, int AL, int BL, int BM, int BN>
__device__ __forceinline__ void ws_produce_tile(const GemmArgs& g, const GemmProblem& P, const GemmTile& T, int it_begin, int n_iters,
                                                uint32_t smem, uint32_t bars, int stage_bytes, uint32_t& it_glob, int tid) {
    using TR = ElemTraits<CPLX>;
    constexpr int ES = TR::ES, BK = TR::BK;
    constexpr int A_BYTES = BM * kRowBytes;
    const GemmSegment* __restrict__ segs = g.segs;
    const bool base_aligned = g.base_aligned != 0;
    TileLoader<CPLX, AL, BM> la;
    TileLoader<CPLX, BL, BN> lb;
    int ld_seg = P.seg_begin, ld_k0 = 0, ld_K = 0;
    {   // locate iteration it_begin
        int skip = it_begin;
        for (;; ++ld_seg) {
            const int K = segs[ld_seg].K;
            const int n = (K + BK - 1) / BK;
            if (skip < n) {
                ld_k0 = skip * BK;
                ld_K = K;
                break;
            }
            skip -= n;
        }
    }
    bool fresh = true;
    for (int it = 0; it < n_iters; ++it) {
        if (ld_k0 >= ld_K) {   // next segment with K > 0
            do {
                ++ld_seg;
                ld_K = segs[ld_seg].K;
            } while (ld_K <= 0);
            ld_k0 = 0;
            fresh = true;
        }
        if (fresh) {
            const GemmSegment& S = segs[ld_seg];
            la.init(g.A + S.offA * ES, S.sAm, S.sAk, P.M, T.m0, ld_k0, base_aligned && (S.align & 1), tid);
            lb.init(g.B + S.offB * ES, S.sBn, S.sBk, P.N, T.n0, ld_k0, base_aligned && (S.align & 2), tid);
            fresh = false;
        }
        const uint32_t stage = it_glob % kStages, phase = (it_glob / kStages) & 1;
        mbar_wait(bars + (kStages + stage) * 8, phase ^ 1);            // the consumers have released this stage
        const uint32_t sa = smem + stage * stage_bytes, sb = sa + A_BYTES;
        la.issue(sa, ld_K - ld_k0);
        lb.issue(sb, ld_K - ld_k0);
        mbar_arrive_cp_async(bars + stage * 8);
        ld_k0 += BK;
        ++it_glob;
    }
}

// ---- consumer: warpgroup 0, the 2 x 2 warp layout of TileKernel -------------------------------------------------------
// EDGE: the tile sticks out of the problem; warps whose part of it holds no valid row or column skip the math (a sector of 68
// rows is one full row of tiles plus one whose lower two warps have nothing to do).  Finer skipping — guarding every 8 x 8
// slice — was measured and rejected: predicated DMMAs lose the issue cadence of the unguarded loop, an edge tile then costs
// MORE than multiplying its padding (U1 D=16384 P2 float64: 3.07 -> 4.63 ms; profiles/kernel_table_r02_edge_skip.json).
template <bool CPLX, int AL, int BL, int BM, int BN, bool EDGE>
__device__ __forceinline__ void ws_consume_tile(const GemmArgs& g, const GemmProblem& P, const GemmTile& T, int tile_index, int it_begin,
                                                int it_end, int n_iters, uint32_t smem, uint32_t bars, int stage_bytes, uint32_t& it_glob) {
    using TR = ElemTraits<CPLX>;
    constexpr int ES = TR::ES, KSTEPS = TR::KSTEPS;
    constexpr int WM = BM / 2, WN = BN / 2, MT = WM / 8, NT = WN / 8;
    constexpr int A_BYTES = BM * kRowBytes;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm0 = (warp >> 1) * WM, wn0 = (warp & 1) * WN;
    const int lx = lane >> 2, lk = lane & 3;
    const bool base_aligned = g.base_aligned != 0;

    double acc[MT][NT][CPLX ? 4 : 2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int c = 0; c < (CPLX ? 4 : 2); ++c) acc[i][j][c] = 0.0;

    const uint32_t sgnA = (CPLX && (g.flags & YB_GEMM_CONJ_A)) ? 0x80000000u : 0u;
    const uint32_t sgnB = (CPLX && (g.flags & YB_GEMM_CONJ_B)) ? 0x80000000u : 0u;
    uint32_t aoff[MT], boff[NT];
#pragma unroll
    for (int i = 0; i < MT; ++i) aoff[i] = frag_offset<CPLX, AL, BM>(wm0 + i * 8 + lx, lk);
#pragma unroll
    for (int j = 0; j < NT; ++j) boff[j] = A_BYTES + frag_offset<CPLX, BL, BN>(wn0 + j * 8 + lx, lk);
    // a warp whose 32 x 64 (16 x 16 ...) part of an edge tile lies completely outside the problem only keeps the ring moving
    const bool idle = EDGE && (T.m0 + wm0 >= P.M || T.n0 + wn0 >= P.N);

    for (int it = 0; it < n_iters; ++it) {
        const uint32_t stage = it_glob % kStages, phase = (it_glob / kStages) & 1;
        mbar_wait(bars + stage * 8, phase);                              // the stage has landed
        const uint32_t sa = smem + stage * stage_bytes;
        if (!idle) {
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            if constexpr (!CPLX) {
                double af[MT], bf[NT];
#pragma unroll
                for (int i = 0; i < MT; ++i) af[i] = lds64(sa + (AL == KC ? (aoff[i] ^ (ks * 32)) : (aoff[i] + ks * 4 * BM * ES)));
#pragma unroll
                for (int j = 0; j < NT; ++j) bf[j] = lds64(sa + (BL == KC ? (boff[j] ^ (ks * 32)) : (boff[j] + ks * 4 * BN * ES)));
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            } else {
                double2 af[MT], bf[NT];
                double naf[MT];
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    af[i] = lds128(sa + (AL == KC ? (aoff[i] ^ (ks * 64)) : (aoff[i] + ks * 4 * BM * ES)));
                    af[i].y = flip_sign(af[i].y, sgnA);
                    naf[i] = flip_sign(af[i].y, 0x80000000u);
                }
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    bf[j] = lds128(sa + (BL == KC ? (boff[j] ^ (ks * 64)) : (boff[j] + ks * 4 * BN * ES)));
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
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + (kStages + stage) * 8);        // this warp is done reading the stage
        ++it_glob;
    }

    // ---- stream-K fix-up (consumer threads only; same protocol as gemm_kernel) -------------------------------------------
    const bool starts = it_begin == 0, ends = it_end >= T.iters;
    if (!starts) {
        double* slot = g.ws + (size_t)blockIdx.x * kWsDoubles + tid;
        int r = 0;
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int c = 0; c < (CPLX ? 4 : 2); ++c) __stcg(slot + (r++) * kGemmThreads, acc[i][j][c]);
        __threadfence();
        consumer_bar_sync();
        if (tid == 0) st_release(g.sync_flags + blockIdx.x, 1);
        return;
    }
    if (!ends) {
        for (int d = blockIdx.x + 1;; ++d) {
            const CtaRange Rd = g.ranges[d];
            if (tid == 0) {
                while (ld_acquire(g.sync_flags + d) == 0) __nanosleep(64);
                g.sync_flags[d] = 0;
            }
            consumer_bar_sync();
            const double* slot = g.ws + (size_t)d * kWsDoubles + tid;
            int r = 0;
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int c = 0; c < (CPLX ? 4 : 2); ++c) acc[i][j][c] += __ldcg(slot + (r++) * kGemmThreads);
            if (Rd.tile_end > tile_index || Rd.it_end >= T.iters) break;
        }
    }

    // ---- epilogue ---------------------------------------------------------------------------------------------------------
    using T2 = typename std::conditional<CPLX, double2, double>::type;
    auto value = [&](int i, int j, int e) -> T2 {
        if constexpr (CPLX) return make_double2(acc[i][j][e], acc[i][j][2 + e]);
        else return acc[i][j][e];
    };
    if (P.scat < 0) {
        T2* Cp = reinterpret_cast<T2*>(g.C) + P.offC;
        const bool vec = !CPLX && base_aligned && ((P.offC & 1) == 0) && ((P.ldc & 1) == 0);
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int row = T.m0 + wm0 + i * 8 + lx;
            if (row < P.M) {
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const int col = T.n0 + wn0 + j * 8 + 2 * lk;
                    T2* p = Cp + (int64_t)row * P.ldc + col;
                    if constexpr (!CPLX) {
                        if (vec && col + 1 < P.N) {
                            *reinterpret_cast<double2*>(p) = make_double2(acc[i][j][0], acc[i][j][1]);
                            continue;
                        }
                    }
                    if (col < P.N) p[0] = value(i, j, 0);
                    if (col + 1 < P.N) p[1] = value(i, j, 1);
                }
            }
        }
    } else {
        const ScatterInfo S = g.scat[P.scat];
        T2* out = reinterpret_cast<T2*>(g.C);
        int2 ri[MT];
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            const int row = T.m0 + wm0 + i * 8 + lx;
            ri[i] = row < P.M ? g.rowinfo[S.row_off + row] : make_int2(-1, 0);
        }
        const int64_t* dst = g.dstpool + S.dst_off;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int col = T.n0 + wn0 + j * 8 + 2 * lk;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (col + e < P.N) {
                    const int4 ci = g.colinfo[S.col_off + col + e];
#pragma unroll
                    for (int i = 0; i < MT; ++i)
                        if (ri[i].x >= 0) out[dst[(int64_t)ri[i].x * S.ncs + ci.x] + (int64_t)ri[i].y * ci.z + ci.y] = value(i, j, e);
                }
            }
        }
    }
}

template <bool CPLX, int AL, int BL>
__global__ void __launch_bounds__(kWsThreads, 2) gemm_ws_kernel(const GemmArgs g) {
    extern __shared__ __align__(128) char smem_raw[];
    using W = WsShape<CPLX, AL, BL>;
    using G = GroupKernel<CPLX, AL, BL>;
    const uint32_t smem = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t bars = smem + W::STAGE_BYTES * kStages;      // full[0..kStages), empty[0..kStages)
    const int tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bars + s * 8, kGemmThreads);               // one cp.async arrive per producer thread
            mbar_init(bars + (kStages + s) * 8, kGemmThreads / 32);   // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const CtaRange R = g.ranges[blockIdx.x];
    uint32_t it_glob = 0;
    if (tid >= kGemmThreads) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(kWsProducerRegs));
        const int ptid = tid - kGemmThreads;
        for (int t = R.tile_begin; t <= R.tile_end; ++t) {
            const GemmTile T = g.tiles[t];
            const GemmProblem P = g.problems[T.prob];
            const int ib = (t == R.tile_begin) ? R.it_begin : 0;
            const int ie = (t == R.tile_end) ? R.it_end : max(T.iters, 1);
            const int n = ws_tile_iters(T, ib, ie);
            if (n == 0) continue;
            if (T.cfg == 0)
                ws_produce_tile<CPLX, AL, BL, G::Big::BM_, G::Big::BN_>(g, P, T, ib, n, smem, bars, W::STAGE_BYTES, it_glob, ptid);
            else
                ws_produce_tile<CPLX, AL, BL, G::Small::BM_, G::Small::BN_>(g, P, T, ib, n, smem, bars, W::STAGE_BYTES, it_glob, ptid);
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(kWsConsumerRegs));
        for (int t = R.tile_begin; t <= R.tile_end; ++t) {
            const GemmTile T = g.tiles[t];
            const GemmProblem P = g.problems[T.prob];
            const int ib = (t == R.tile_begin) ? R.it_begin : 0;
            const int ie = (t == R.tile_end) ? R.it_end : max(T.iters, 1);
            const int n = ws_tile_iters(T, ib, ie);
            const int bm = T.cfg == 0 ? G::Big::BM_ : G::Small::BM_, bn = T.cfg == 0 ? G::Big::BN_ : G::Small::BN_;
            const bool edge = (T.m0 + bm > P.M || T.n0 + bn > P.N) && !(g.flags & kFlagNoEdgeSkip);
            if (T.cfg == 0) {
                if (edge) ws_consume_tile<CPLX, AL, BL, G::Big::BM_, G::Big::BN_, true>(g, P, T, t, ib, ie, n, smem, bars, W::STAGE_BYTES, it_glob);
                else ws_consume_tile<CPLX, AL, BL, G::Big::BM_, G::Big::BN_, false>(g, P, T, t, ib, ie, n, smem, bars, W::STAGE_BYTES, it_glob);
            } else {
                if (edge) ws_consume_tile<CPLX, AL, BL, G::Small::BM_, G::Small::BN_, true>(g, P, T, t, ib, ie, n, smem, bars, W::STAGE_BYTES, it_glob);
                else ws_consume_tile<CPLX, AL, BL, G::Small::BM_, G::Small::BN_, false>(g, P, T, t, ib, ie, n, smem, bars, W::STAGE_BYTES, it_glob);
            }
        }
    }
}

}  // namespace yb

using namespace yb;

struct yb_gemm_plan {
    int dtype = 0, device = 0;
    int al = KC, bl = XC;
    int ntiles = 0, grid = 0, nsplit = 0;
    int64_t macs = 0, nbig = 0, nsmall = 0;
    DeviceTable problems, segments, tiles, ranges, scat, rowinfo, colinfo, dstpool;
    int ws_slots = 0;       // stream-K partial-tile slots this plan needs (0: no CTA starts inside a tile)
    bool cooperative = false;
    SkinnyPlan* skinny = nullptr;   // problems with a tiny result block and a long contraction index (yb_skinny.cu)
    PanelPlan* panel = nullptr;     // a tiny matrix times a very long one (yb_panel.cu)
};

namespace {

// Kernel variant: the warp-specialised kernel unless YB_GEMM_CLASSIC=1 is set in the environment (A/B measurements).
bool use_ws_kernel() {
    static const bool ws = [] {
        const char* e = getenv("YB_GEMM_CLASSIC");
        return !(e && e[0] == '1');
    }();
    return ws;
}

template <bool CPLX, int AL, int BL>
int occupancy(int device, int* blocks_per_sm) {
    using G = GroupKernel<CPLX, AL, BL>;
    using W = WsShape<CPLX, AL, BL>;
    static int cached[64] = {0};   // per device: the attribute has to be set once per context, the answer never changes
    if (device >= 0 && device < 64 && cached[device] > 0) {
        *blocks_per_sm = cached[device];
        return kOk;
    }
    if (use_ws_kernel()) {
        YB_CUDA(cudaFuncSetAttribute(gemm_ws_kernel<CPLX, AL, BL>, cudaFuncAttributeMaxDynamicSharedMemorySize, W::SMEM_BYTES));
        YB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, gemm_ws_kernel<CPLX, AL, BL>, kWsThreads, W::SMEM_BYTES));
    } else {
        YB_CUDA(cudaFuncSetAttribute(gemm_kernel<CPLX, AL, BL>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
        YB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, gemm_kernel<CPLX, AL, BL>, kGemmThreads, G::SMEM_BYTES));
    }
    if (device >= 0 && device < 64) cached[device] = *blocks_per_sm;
    return kOk;
}

template <bool CPLX>
int occupancy_layout(int al, int bl, int device, int* b) {
    if (al == KC && bl == XC) return occupancy<CPLX, KC, XC>(device, b);
    if (al == KC && bl == KC) return occupancy<CPLX, KC, KC>(device, b);
    if (al == XC && bl == XC) return occupancy<CPLX, XC, XC>(device, b);
    return occupancy<CPLX, XC, KC>(device, b);
}

template <bool CPLX, int AL, int BL>
int launch(const yb_gemm_plan* p, const GemmArgs& args, cudaStream_t st) {
    using G = GroupKernel<CPLX, AL, BL>;
    using W = WsShape<CPLX, AL, BL>;
    const bool ws = use_ws_kernel();
    const void* fn = ws ? (const void*)gemm_ws_kernel<CPLX, AL, BL> : (const void*)gemm_kernel<CPLX, AL, BL>;
    const int threads = ws ? kWsThreads : kGemmThreads;
    const size_t smem = ws ? (size_t)W::SMEM_BYTES : (size_t)G::SMEM_BYTES;
    if (p->cooperative) {
        // CTAs that finish a split tile wait for the partials of other CTAs: a cooperative launch makes the driver
        // guarantee that the whole grid (<= SMs x occupancy) is resident at once, whatever else runs on the device
        void* kargs[] = {(void*)&args};
        YB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(p->grid), dim3(threads), kargs, smem, st));
    } else if (ws) {
        gemm_ws_kernel<CPLX, AL, BL><<<p->grid, kWsThreads, W::SMEM_BYTES, st>>>(args);
    } else {
        gemm_kernel<CPLX, AL, BL><<<p->grid, kGemmThreads, G::SMEM_BYTES, st>>>(args);
    }
    YB_CUDA(cudaGetLastError());
    return kOk;
}

template <bool CPLX>
int dispatch_layout(const yb_gemm_plan* p, const GemmArgs& args, cudaStream_t st) {
    if (p->al == KC && p->bl == XC) return launch<CPLX, KC, XC>(p, args, st);
    if (p->al == KC && p->bl == KC) return launch<CPLX, KC, KC>(p, args, st);
    if (p->al == XC && p->bl == XC) return launch<CPLX, XC, XC>(p, args, st);
    return launch<CPLX, XC, KC>(p, args, st);
}

int create_plan(const int64_t* problems, int64_t nprob, const int64_t* segments, int64_t nseg, const int64_t* scat_index,
                int64_t nscat, const int64_t* row_ptr, const int64_t* row_cuts, const int64_t* col_ptr, const int64_t* col_cuts,
                const int64_t* dst_ptr, const int64_t* dst, int dtype, int device, yb_gemm_plan** out) {
    if (!out) return fail(kErrArg, "yb_gemm_plan_create: out is null");
    *out = nullptr;
    if (dtype != YB_F64 && dtype != YB_C128) return fail(kErrUnsupported, "yb_gemm_plan_create: dtype %d", dtype);
    if (nprob < 0 || nseg < 0 || (nprob > 0 && !problems) || (nseg > 0 && !segments))
        return fail(kErrArg, "yb_gemm_plan_create: bad tables");
    const bool cplx = dtype == YB_C128;
    const int ES = cplx ? 16 : 8, BK = cplx ? 8 : 16;

    // operand layouts: decided by which stride is 1; must be consistent over the plan
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
        g.scat = -1;
        g.pad_ = 0;
        if (scat_index) {
            if (scat_index[i] < -1 || scat_index[i] >= nscat) return fail(kErrArg, "yb_gemm_plan_create: problem %lld scatter index", (long long)i);
            g.scat = (int32_t)scat_index[i];
        }
    }

    // ---- routing: tiny result blocks with a long contraction index are reductions, not tiles (yb_skinny.cu) -----------
    std::vector<char> is_skinny((size_t)nprob, 0);
    std::vector<int> skinny_set;
    {
        auto p2 = [](int v) { int p = 1; while (p < v) p <<= 1; return p; };
        auto ksum = [&](const GemmProblem& g) {
            int64_t k = 0;
            for (int s = g.seg_begin; s < g.seg_end; ++s) k += hs[(size_t)s].K;
            return k;
        };
        int lim = kSkinnyMax;
        for (int pass = 0; pass < 2; ++pass) {
            skinny_set.clear();
            bool all = true;
            int mx = 1, nx = 1;
            for (int64_t i = 0; i < nprob; ++i) {
                const GemmProblem& g = hp[(size_t)i];
                if (g.M == 0 || g.N == 0) continue;
                if (g.M <= lim && g.N <= lim && p2(g.M) * p2(g.N) <= skinny_entries(cplx)) {
                    skinny_set.push_back((int)i);
                    mx = std::max(mx, (int)g.M);
                    nx = std::max(nx, (int)g.N);
                } else {
                    all = false;
                }
            }
            if (!all) {   // mixed plan: a second launch only pays for long contractions
                std::vector<int> keep;
                for (int i : skinny_set)
                    if (ksum(hp[(size_t)i]) >= 2048) keep.push_back(i);
                skinny_set.swap(keep);
            }
            if (p2(mx) * p2(nx) <= skinny_entries(cplx)) break;
            lim = 4;      // complex128: 8 x 2 and 2 x 8 blocks in one plan would need 8 x 8 accumulators per lane
        }
        for (int i : skinny_set) is_skinny[(size_t)i] = 1;
    }
    // ... and a tiny matrix times a very long one is a streaming product (yb_panel.cu).  `is_skinny` marks both kinds: neither
    // is tiled.
    std::vector<int> panel_set;
    for (int64_t i = 0; i < nprob; ++i)
        if (!is_skinny[(size_t)i] && panel_eligible(hp[(size_t)i], hs)) {
            panel_set.push_back((int)i);
            is_skinny[(size_t)i] = 1;
        }

    for (int64_t i = 0; i < nprob; ++i) {
        const GemmProblem& g = hp[(size_t)i];
        if (is_skinny[(size_t)i]) continue;    // the skinny kernel takes any strides
        // layout feasibility: a layout is usable when its contiguous index has unit stride (or extent <= 1)
        for (int s = g.seg_begin; s < g.seg_end; ++s) {
            const GemmSegment& sg = hs[(size_t)s];
            if (g.M == 0 || g.N == 0 || sg.K == 0) continue;
            a_ok[KC] = a_ok[KC] && (sg.sAk == 1 || sg.K <= 1);
            a_ok[XC] = a_ok[XC] && (sg.sAm == 1 || g.M <= 1);
            b_ok[XC] = b_ok[XC] && (sg.sBn == 1 || g.N <= 1);
            b_ok[KC] = b_ok[KC] && (sg.sBk == 1 || sg.K <= 1);
            // the loaders step through the strided dim with 32-bit byte strides
            const int64_t lim = (1ll << 32) / (16 * ES) - 1;
            if (sg.sAm > lim || sg.sAk > lim || sg.sBk > lim || sg.sBn > lim)
                return fail(kErrUnsupported, "yb_gemm_plan_create: leading dimension above %lld elements", (long long)lim);
        }
    }
    if (!a_ok[KC] && !a_ok[XC]) return fail(kErrUnsupported, "yb_gemm_plan_create: operand A has no common unit-stride index");
    if (!b_ok[KC] && !b_ok[XC]) return fail(kErrUnsupported, "yb_gemm_plan_create: operand B has no common unit-stride index");
    const int al = a_ok[KC] ? KC : XC;
    const int bl = b_ok[XC] ? XC : KC;
    // 16-byte alignment of every row start (f64 only; complex elements are 16 bytes)
    for (auto& sg : hs) {
        const int64_t a_ld = (al == KC) ? sg.sAm : sg.sAk;
        const int64_t b_ld = (bl == XC) ? sg.sBk : sg.sBn;
        if (((sg.offA | a_ld) & 1) == 0) sg.align |= 1;
        if (((sg.offB | b_ld) & 1) == 0) sg.align |= 2;
    }

    // scatter tables (fused unmerge)
    std::vector<ScatterInfo> hsc((size_t)nscat);
    std::vector<int2> hrow;
    std::vector<int4> hcol;
    std::vector<int64_t> hdst;
    for (int64_t s = 0; s < nscat; ++s) {
        const int64_t nrs = row_ptr[s + 1] - row_ptr[s] - 1, ncs = col_ptr[s + 1] - col_ptr[s] - 1;
        if (nrs < 1 || ncs < 1 || dst_ptr[s + 1] - dst_ptr[s] != nrs * ncs)
            return fail(kErrArg, "yb_gemm_plan_create: scatter %lld has inconsistent cut / destination tables", (long long)s);
        ScatterInfo& si = hsc[(size_t)s];
        si.dst_off = (int64_t)hdst.size();
        si.row_off = (int32_t)hrow.size();
        si.col_off = (int32_t)hcol.size();
        si.ncs = (int32_t)ncs;
        si.pad_ = 0;
        const int64_t* rc = row_cuts + row_ptr[s];
        const int64_t* cc = col_cuts + col_ptr[s];
        if (rc[0] != 0 || cc[0] != 0) return fail(kErrArg, "yb_gemm_plan_create: scatter %lld cuts must start at 0", (long long)s);
        for (int64_t i = 0; i < nrs; ++i) {
            if (rc[i + 1] <= rc[i]) return fail(kErrArg, "yb_gemm_plan_create: scatter %lld row cuts not increasing", (long long)s);
            for (int64_t r = rc[i]; r < rc[i + 1]; ++r) hrow.push_back(make_int2((int)i, (int)(r - rc[i])));
        }
        for (int64_t j = 0; j < ncs; ++j) {
            if (cc[j + 1] <= cc[j]) return fail(kErrArg, "yb_gemm_plan_create: scatter %lld col cuts not increasing", (long long)s);
            for (int64_t c = cc[j]; c < cc[j + 1]; ++c) hcol.push_back(make_int4((int)j, (int)(c - cc[j]), (int)(cc[j + 1] - cc[j]), 0));
        }
        hdst.insert(hdst.end(), dst + dst_ptr[s], dst + dst_ptr[s + 1]);
    }
    for (int64_t i = 0; i < nprob; ++i) {
        const GemmProblem& g = hp[(size_t)i];
        if (g.scat < 0) continue;
        const int64_t s = g.scat;
        if (row_cuts[row_ptr[s + 1] - 1] != g.M || col_cuts[col_ptr[s + 1] - 1] != g.N)
            return fail(kErrArg, "yb_gemm_plan_create: scatter cuts of problem %lld do not cover its %d x %d block", (long long)i, g.M, g.N);
    }

    // tiles in natural order: problem by problem, row-major inside a problem
    // Scratch that survives between plan creations of a thread: a tall-and-skinny problem (an MPO tensor applied to a D=4096
    // two-site tensor) has ~10^6 tiles, and building 29 MB vectors from fresh pages cost more (page faults, regrowth) than
    // filling them: 26 ms per plan, 10 % of a DMRG sweep.
    struct PlanScratch {
        std::vector<GemmTile> tiles, line;
        std::vector<int64_t> wbegin;
    };
    static thread_local PlanScratch scratch;
    std::vector<GemmTile>& ht = scratch.tiles;
    ht.clear();
    int64_t macs = 0, nbig = 0, nsmall = 0, W = 0;
    const int bigM = 64, bigN = cplx ? 64 : 128, smM = cplx ? 32 : 64, smN = cplx ? 32 : 64;
    auto tile_weight = [](const GemmTile& t) { return std::max<int64_t>(t.iters, 1) * (t.cfg == 0 ? 2 : 1); };
    for (int64_t i = 0; i < nprob; ++i) {
        const GemmProblem& g = hp[(size_t)i];
        if (g.M == 0 || g.N == 0 || is_skinny[(size_t)i]) continue;
        int64_t ksum = 0, iters = 0;
        for (int s = g.seg_begin; s < g.seg_end; ++s) {
            ksum += hs[(size_t)s].K;
            iters += (hs[(size_t)s].K + BK - 1) / BK;
        }
        if (iters > 0x7fffffff) return fail(kErrUnsupported, "yb_gemm_plan_create: problem %lld contraction too long", (long long)i);
        macs += (int64_t)g.M * g.N * ksum;
        const bool big = g.N > smN && g.M > smM / 2;
        const int bm = big ? bigM : smM, bn = big ? bigN : smN;
        // Column bands: the ~300 tiles in flight at any time (consecutive tiles of this order, see the work line below)
        // then cover a near-square region of C, so a B band stays in L2 while the A row panels stream past it once per
        // band.  Plain row-major order re-read B from DRAM for every wave of tiles (ncu: 11.2 GB read for a launch whose
        // operands are 0.9 GB).
        const int ntn = (g.N + bn - 1) / bn;
        int bw = (int)std::lround(std::sqrt((double)kTilesInFlight * bm / bn));
        bw = std::max(1, std::min(bw, ntn));
        const int nbands = (ntn + bw - 1) / bw;
        bw = (ntn + nbands - 1) / nbands;
        const int64_t ntm = (g.M + bm - 1) / bm, count = ntm * ntn;
        const size_t base = ht.size();
        ht.resize(base + (size_t)count);
        GemmTile* out = ht.data() + base;
        const GemmTile proto = {(int32_t)i, 0, 0, big ? 0 : 1, (int32_t)iters, {0, 0, 0}};
        for (int band = 0; band < nbands; ++band) {
            const int tn0 = band * bw, tn1 = std::min(ntn, (band + 1) * bw);
            for (int m0 = 0; m0 < g.M; m0 += bm)
                for (int tn = tn0; tn < tn1; ++tn) {
                    GemmTile t = proto;
                    t.m0 = m0;
                    t.n0 = tn * bn;
                    *out++ = t;
                }
        }
        W += count * tile_weight(proto);
        (big ? nbig : nsmall) += count;
    }

    for (const std::vector<int>* set : {&skinny_set, &panel_set})
        for (int i : *set) {
            const GemmProblem& g = hp[(size_t)i];
            for (int s = g.seg_begin; s < g.seg_end; ++s) macs += (int64_t)g.M * g.N * hs[(size_t)s].K;
        }

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

    // stream-K shares: one per resident CTA
    std::vector<CtaRange> hr;
    int max_grid = 0;
    if (rc == kOk && !ht.empty()) {
        int sms = 148, occ = 2;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        rc = cplx ? occupancy_layout<true>(al, bl, device, &occ) : occupancy_layout<false>(al, bl, device, &occ);
        if (rc == kOk && occ < 1) rc = fail(kErrCuda, "yb_gemm_plan_create: kernel does not fit on an SM");
        if (rc == kOk) {
            const int64_t maxG = (int64_t)sms * std::min(occ, 2);
            max_grid = (int)maxG;
            const int64_t minShare = 16;   // weighted iterations: a share below ~8 big-tile iterations is dominated by fix-up traffic
            const int64_t G = std::max<int64_t>(1, std::min(maxG, W / minShare));
            const int64_t share = (W + G - 1) / G;
            // Work line: CTA c walks its share sequentially, so at any time the G CTAs sit at the same depth of their
            // shares.  Dealing the natural order round-robin over G bins and concatenating the bins makes those
            // simultaneously processed tiles neighbours in the natural order: they share operand panels in L2.
            {
                std::vector<GemmTile>& line = scratch.line;
                line.resize(ht.size());
                GemmTile* out = line.data();
                const GemmTile* in = ht.data();
                const size_t n = ht.size();
                for (int64_t c = 0; c < G; ++c)
                    for (size_t i = (size_t)c; i < n; i += (size_t)G) *out++ = in[i];
                ht.swap(line);
            }
            std::vector<int64_t>& wbegin = scratch.wbegin;   // weighted start of every tile on the work line
            wbegin.assign(ht.size() + 1, 0);
            for (size_t i = 0; i < ht.size(); ++i) wbegin[i + 1] = wbegin[i] + tile_weight(ht[i]);
            // position of a cut on the work line: (tile, iteration); cuts inside tiles much smaller than a share
            // snap to the tile start (no fix-up traffic for an imbalance below 1/16 share)
            auto locate = [&](int64_t w, int& tile, int& it) {
                if (w >= W) {
                    tile = (int)ht.size();
                    it = 0;
                    return;
                }
                const size_t t = (size_t)(std::upper_bound(wbegin.begin(), wbegin.end(), w) - wbegin.begin()) - 1;
                const int64_t tw = wbegin[t + 1] - wbegin[t];
                tile = (int)t;
                it = (int)((w - wbegin[t]) / (ht[t].cfg == 0 ? 2 : 1));
                if (tw * 16 <= share || ht[t].iters <= 1) it = 0;
            };
            int pt = 0, pi = 0;
            for (int64_t c = 1; c <= G; ++c) {
                int t, i;
                locate(std::min(W, c * share), t, i);
                if (t == pt && i == pi) continue;   // empty share
                CtaRange r;
                r.tile_begin = pt;
                r.it_begin = pi;
                if (i == 0) {
                    r.tile_end = t - 1;
                    r.it_end = std::max(ht[(size_t)t - 1].iters, 1);
                } else {
                    r.tile_end = t;
                    r.it_end = i;
                }
                hr.push_back(r);
                if (r.it_begin > 0) plan->nsplit++;
                pt = t;
                pi = i;
            }
            plan->grid = (int)hr.size();
        }
    }
    if (rc == kOk) {
        TableBatch up;     // one pool block, one copy
        up.add(plan->problems, hp.data(), hp.size() * sizeof(GemmProblem));
        up.add(plan->segments, hs.data(), hs.size() * sizeof(GemmSegment));
        up.add(plan->tiles, ht.data(), ht.size() * sizeof(GemmTile));
        up.add(plan->ranges, hr.data(), hr.size() * sizeof(CtaRange));
        up.add(plan->scat, hsc.data(), hsc.size() * sizeof(ScatterInfo));
        up.add(plan->rowinfo, hrow.data(), hrow.size() * sizeof(int2));
        up.add(plan->colinfo, hcol.data(), hcol.size() * sizeof(int4));
        up.add(plan->dstpool, hdst.data(), hdst.size() * sizeof(int64_t));
        rc = up.commit();
    }
    if (rc == kOk && plan->nsplit > 0) {
        plan->ws_slots = max_grid;
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
        plan->cooperative = coop != 0;
    }
    if (rc == kOk && !skinny_set.empty()) rc = skinny_create(hp, hs, skinny_set, cplx, device, &plan->skinny);
    if (rc == kOk && !panel_set.empty()) rc = panel_create(hp, hs, panel_set, cplx, device, &plan->panel);
    if (rc == kOk && (plan->ws_slots > 0 || plan->skinny)) rc = workspace_reserve(device);
    cudaSetDevice(prev);
    if (rc != kOk) {
        yb_gemm_plan_destroy(plan);
        return rc;
    }
    *out = plan;
    return kOk;
}

}  // namespace

extern "C" int yb_gemm_plan_create(const int64_t* problems, int64_t nprob, const int64_t* segments, int64_t nseg,
                                   int dtype, int device, yb_gemm_plan** out) {
    return create_plan(problems, nprob, segments, nseg, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, dtype, device, out);
}

extern "C" int yb_gemm_plan_create_scatter(const int64_t* problems, int64_t nprob, const int64_t* segments, int64_t nseg,
                                           const int64_t* scat_index, int64_t nscat, const int64_t* row_ptr,
                                           const int64_t* row_cuts, const int64_t* col_ptr, const int64_t* col_cuts,
                                           const int64_t* dst_ptr, const int64_t* dst, int dtype, int device,
                                           yb_gemm_plan** out) {
    if (nscat < 0 || (nscat > 0 && (!scat_index || !row_ptr || !row_cuts || !col_ptr || !col_cuts || !dst_ptr || !dst)))
        return fail(kErrArg, "yb_gemm_plan_create_scatter: bad scatter tables");
    return create_plan(problems, nprob, segments, nseg, scat_index, nscat, row_ptr, row_cuts, col_ptr, col_cuts, dst_ptr, dst, dtype, device, out);
}

extern "C" int yb_gemm_plan_info(const yb_gemm_plan* plan, int64_t info[10]) {
    if (!plan || !info) return fail(kErrArg, "yb_gemm_plan_info: null argument");
    info[0] = plan->ntiles;
    info[1] = plan->macs;
    info[2] = plan->nbig;
    info[3] = plan->nsmall;
    info[4] = plan->grid;
    info[5] = plan->nsplit;
    skinny_info(plan->skinny, &info[6], &info[7]);
    info[8] = panel_parts(plan->panel);
    info[9] = 0;
    return kOk;
}

extern "C" int yb_gemm_run(const yb_gemm_plan* plan, const void* A, const void* B, void* C, int flags, void* stream) {
    if (!plan) return fail(kErrArg, "yb_gemm_run: plan is null");
    if (plan->ntiles == 0 && !plan->skinny && !plan->panel) return kOk;
    if (!A || !B || !C) return fail(kErrArg, "yb_gemm_run: null data pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->dtype == YB_C128 && ((((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) != 0))
        return fail(kErrArg, "yb_gemm_run: complex128 operands must be 16-byte aligned");
    if (plan->skinny || plan->panel) {
        ScatterTables sc = {(const ScatterInfo*)plan->scat.ptr, (const int2*)plan->rowinfo.ptr, (const int4*)plan->colinfo.ptr,
                            (const int64_t*)plan->dstpool.ptr};
        int rc = kOk;
        if (plan->skinny)
            rc = skinny_run(plan->skinny, (const GemmProblem*)plan->problems.ptr, (const GemmSegment*)plan->segments.ptr, sc, A, B, C, flags, st);
        if (rc == kOk && plan->panel)
            rc = panel_run(plan->panel, (const GemmProblem*)plan->problems.ptr, (const GemmSegment*)plan->segments.ptr, sc, A, B, C, flags, st);
        if (rc != kOk) return rc;
    }
    if (plan->ntiles == 0) return kOk;
    GemmArgs args;
    args.problems = (const GemmProblem*)plan->problems.ptr;
    args.segs = (const GemmSegment*)plan->segments.ptr;
    args.tiles = (const GemmTile*)plan->tiles.ptr;
    args.ranges = (const CtaRange*)plan->ranges.ptr;
    args.scat = (const ScatterInfo*)plan->scat.ptr;
    args.rowinfo = (const int2*)plan->rowinfo.ptr;
    args.colinfo = (const int4*)plan->colinfo.ptr;
    args.dstpool = (const int64_t*)plan->dstpool.ptr;
    args.ws = nullptr;
    args.sync_flags = nullptr;
    if (plan->ws_slots > 0) {   // looked up per launch: one workspace per (device, stream), never shared between streams
        void* ws = nullptr;
        int rc = stream_workspace(plan->device, st, (size_t)plan->ws_slots * kWsDoubles * sizeof(double), (size_t)plan->ws_slots, &ws,
                                  &args.sync_flags);
        if (rc != kOk) return rc;
        args.ws = (double*)ws;
    }
    args.A = (const char*)A;
    args.B = (const char*)B;
    args.C = (char*)C;
    static const bool no_skip = [] { const char* e = getenv("YB_GEMM_NOSKIP"); return e && e[0] == '1'; }();
    args.flags = (flags & 0xff) | (no_skip ? kFlagNoEdgeSkip : 0);
    args.base_aligned = (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) == 0;
    if (plan->dtype == YB_C128) return dispatch_layout<true>(plan, args, st);
    return dispatch_layout<false>(plan, args, st);
}

extern "C" void yb_gemm_plan_destroy(yb_gemm_plan* plan) {
    if (!plan) return;
    plan->problems.release();
    plan->segments.release();
    plan->tiles.release();
    plan->ranges.release();
    plan->scat.release();
    plan->rowinfo.release();
    plan->colinfo.release();
    plan->dstpool.release();
    skinny_destroy(plan->skinny);
    panel_destroy(plan->panel);
    delete plan;
}
