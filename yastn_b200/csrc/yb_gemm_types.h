// Tables shared by the grouped-GEMM kernels (yb_gemm.cu: DMMA tiles; yb_skinny.cu: huge-K / tiny-output reductions).
#pragma once
#include "yb_common.h"

namespace yb {

struct GemmProblem {
    int32_t M, N;
    int32_t seg_begin, seg_end;
    int64_t offC, ldc;
    int32_t scat;      // index into the scatter table, -1: plain row-major store at offC / ldc
    int32_t pad_;
};

struct GemmSegment {
    int64_t offA, offB;
    int64_t sAm, sAk, sBk, sBn;
    int32_t K;
    int32_t align;  // bit0: A rows 16B-aligned, bit1: B rows 16B-aligned (relative to a 16B-aligned base pointer)
};

// Fused unmerge: element (r, c) of the merged block goes to  dst[rowinfo[r].x * ncs + colinfo[c].x]
//                                                            + rowinfo[r].y * colinfo[c].z + colinfo[c].y
struct ScatterInfo {
    int64_t dst_off;             // into the int64 pool of destination block offsets (nrs x ncs)
    int32_t row_off, col_off;    // into the int2 (row) / int4 (col) pools
    int32_t ncs, pad_;
};

struct ScatterTables {
    const ScatterInfo* scat;
    const int2* rowinfo;
    const int4* colinfo;
    const int64_t* dstpool;
};

// Element offset inside C of entry (r, c) of problem P (plain row-major store or fused-unmerge scatter).
__device__ __forceinline__ int64_t c_offset(const GemmProblem& P, const ScatterTables& S, int r, int c) {
    if (P.scat < 0) return P.offC + (int64_t)r * P.ldc + c;
    const ScatterInfo si = S.scat[P.scat];
    const int2 ri = S.rowinfo[si.row_off + r];
    const int4 ci = S.colinfo[si.col_off + c];
    return S.dstpool[si.dst_off + (int64_t)ri.x * si.ncs + ci.x] + (int64_t)ri.y * ci.z + ci.y;
}

// ---- skinny path (yb_skinny.cu) ------------------------------------------------------------------------------------
constexpr int kSkinnyMax = 8;      // problems with M <= 8 and N <= 8 are reductions over the contraction index, not tiles
// accumulators (padded to powers of two) one lane can hold: 8 x 8 float64, 32 complex128 entries (8 x 4, 4 x 4, ...)
constexpr int skinny_entries(bool cplx) { return cplx ? 32 : 64; }

struct SkinnyPlan;
// `which` lists the problems (indices into hp) that the skinny kernel computes.
int skinny_create(const std::vector<GemmProblem>& hp, const std::vector<GemmSegment>& hs, const std::vector<int>& which,
                  bool cplx, int device, SkinnyPlan** out);
int skinny_run(const SkinnyPlan* plan, const GemmProblem* problems, const GemmSegment* segs, const ScatterTables& scat,
               const void* A, const void* B, void* C, int flags, cudaStream_t st);
void skinny_destroy(SkinnyPlan* plan);
void skinny_info(const SkinnyPlan* plan, int64_t* warps, int64_t* runs);

// ---- panel path (yb_panel.cu): K <= 8 and exactly one of M, N <= 8 while the other is long ---------------------------
constexpr int kPanelMax = 8;
constexpr int kPanelMinStream = 256;   // shorter ones stay tiles (a second launch would cost more than the padding)

struct PanelPlan;
bool panel_eligible(const GemmProblem& P, const std::vector<GemmSegment>& hs);
int panel_create(const std::vector<GemmProblem>& hp, const std::vector<GemmSegment>& hs, const std::vector<int>& which, bool cplx,
                 int device, PanelPlan** out);
int panel_run(const PanelPlan* plan, const GemmProblem* problems, const GemmSegment* segs, const ScatterTables& scat, const void* A,
              const void* B, void* C, int flags, cudaStream_t st);
void panel_destroy(PanelPlan* plan);
int64_t panel_parts(const PanelPlan* plan);

}  // namespace yb
