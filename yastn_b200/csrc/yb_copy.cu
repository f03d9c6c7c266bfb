// Block copy engine: one launch moves every block of a symmetric tensor between two layouts.
//
// Serves backend.transpose_and_merge / unmerge / transpose and their adjoints (reference loops:
// yastn/backend/_backend_torch_backwards.py:315-319, 340-364, 397-408).  The reference issues one
// strided-copy launch per block; here the host normalises every block move into a "record"
// (dims sorted by destination stride, unit dims dropped, mergeable dims coalesced) and the device
// walks a work-item table, ONE WARP PER ITEM (up to 4096 elements), warps never synchronising with each other:
//   * flat items   — a slice of one record, destination-linear, coalesced writes, full index decomposition per element;
//   * row items    — records whose innermost dim is a run contiguous on both sides: long runs move with 16-byte loads,
//                    short runs (>= 16) resolve the outer indices once per row and share them through warp shuffles
//                    (the per-element decomposition made the flat path integer-bound for float64);
//   * pack items   — many small records, one after the other, the next record prefetched while the current one moves;
//   * tiled items  — (src-fast dim != dst-fast dim) a 64-wide 2-d slab staged through shared memory by the whole CTA (second
//                    phase of the kernel) so that both the global reads and the global writes are 512-byte windows.
// HBM-bound: algorithmic bytes = itemsize * (elements read + elements written).
#include <algorithm>
#include <numeric>

#include "yb_common.h"

namespace yb {

constexpr int kMaxDims = 8;
constexpr int kCopyThreads = 256;
constexpr int kCopyWarps = kCopyThreads / 32;
constexpr uint32_t kItemElems = 4096;      // largest work item (one warp); plans with little work use smaller items (item_size)
constexpr uint32_t kSmallRec = 512;        // records below this many elements are packed
constexpr int kPackMaxRecs = 64;
constexpr int kTileA = 64;                 // tiled path: slab extent along the destination-fast dim
constexpr int kRunUnroll = 4;              // long-row path: 16-byte loads in flight per lane
constexpr uint32_t kRowMin = 2048;         // long-row path (16-byte loads) from this run length on
constexpr uint32_t kRowShortMin = 16;      // short-row path: a 256-element trip then spans at most 17 rows (one lane each)

// Slab of the tiled (transposing) path: kTileA (destination-fast) x kTileB (source-fast) elements, one CTA per slab.  Both
// the read windows (kTileB elements) and the write windows (kTileA elements) are 512 bytes or more: block offsets are
// arbitrary multiples of 8 bytes and DRAM is fetched in 64-byte pieces, so a window of n bytes costs n + 64 on average
// (ncu on 128-byte windows, the warp-sized 32x16 slab tried first: 1.6x the algorithmic DRAM reads, 3.6 TB/s).
template <typename T>
struct TileShape {
    static constexpr int kTileB = sizeof(T) == 8 ? 64 : 32;
};

struct alignas(16) CopyRec {
    int64_t src_base, dst_base;
    uint32_t total;
    int32_t nd;
    uint32_t ext[kMaxDims], mul[kMaxDims], shr[kMaxDims];
    uint32_t sstr[kMaxDims], dstr[kMaxDims];
    // tiled path (valid when tile_a >= 0): dims tile_a (dst-fast) and tile_b (src-fast) span the slab
    int32_t tile_a, tile_b;
    uint32_t tiles_a, tiles_b;       // number of slabs along each tiled dim
    uint32_t outer_total;            // product of the other extents
    uint32_t pad_[3];
};
static_assert(sizeof(CopyRec) % 16 == 0 && sizeof(CopyRec) / 4 <= 64, "CopyRec is staged by one warp with two words per lane");
constexpr int kRecWords = sizeof(CopyRec) / 4;

struct CopyItem {
    int32_t rec_begin, rec_end;  // [rec_begin, rec_end) ; a single record unless this is a pack; kind 5: range of the run table
    uint32_t e0, ne;             // flat / rows: element range; tiled: slab range
    int32_t kind;                // 0 flat, 1 pack, 2 tiled, 3 long rows, 4 short rows (3, 4: innermost dim contiguous on both sides),
                                 // 5 batch of up to 32 contiguous runs
    int32_t pad_[3];
};

// A contiguous run of n elements (one row, or a piece of a row, of a small record): the whole description of the move, so a
// warp fetches the metadata of 32 runs with ONE coalesced load.  Tensors made of thousands of sub-kilobyte blocks (U1xU1:
// 2030 blocks, median 132 elements) were metadata-latency bound on the record path: item -> 192-byte record -> data is three
// dependent DRAM round trips per kilobyte moved (ncu: 1.3 TB/s, 20 stall cycles per issue on the long scoreboard).
struct alignas(16) CopyRun {
    int64_t src, dst;
    uint32_t n;
    uint32_t flags;              // bit0: no source, store zeros (cells of a merged block that no source block covers)
    uint32_t pad_[2];
};
static_assert(sizeof(CopyRun) == 32, "CopyRun is fetched as two 16-byte words per lane");
constexpr int kRunBatch = 32;              // runs per batch item (one per lane)
constexpr int64_t kRunRecMax = 1 << 15;    // records up to this many elements go to the run table ...
constexpr int64_t kRunRowMin = 8;          // ... if their rows have at least this many elements
constexpr int64_t kRunPiece = 4096;        // longest run (longer rows are cut)

struct RunSmem {                           // per warp
    int64_t src[kRunBatch], dst[kRunBatch];
    uint32_t start[kRunBatch + 1];
    uint32_t flags[kRunBatch];
};

constexpr int kHostDims = 24;   // rank limit of an input record (YASTN tensors have at most ~12 legs)
struct HostRec {                 // fixed arrays: plan construction handles thousands of records, no per-record heap traffic
    int64_t src_base, dst_base;
    bool zero;                   // no source: fill the destination box with zeros
    int nd;
    int64_t ext[kHostDims], sstr[kHostDims], dstr[kHostDims];
};

__device__ __forceinline__ uint32_t fdiv(uint32_t n, uint32_t d, uint32_t mul, uint32_t shr) {
    return d == 1 ? n : (__umulhi(n, mul) >> shr);
}

template <typename T, bool CONJ>
__device__ __forceinline__ T load_elem(const T* p) {
    T v = *p;
    if constexpr (CONJ) v.y = -v.y;
    return v;
}

// Guarded load: the value is defined on both branches (an array filled by predicated loads only is otherwise kept in
// local memory by the compiler — seen in the SASS as one STL + LDL per element).
template <typename T, bool CONJ>
__device__ __forceinline__ T load_if(bool ok, const T* p) {
    T v = T{};
    if (ok) v = load_elem<T, CONJ>(p);
    return v;
}

// Decompose a destination-linear index of `rec` into source / destination offsets.
template <typename REC>
__device__ __forceinline__ void offsets_of(const REC& rec, uint32_t e, uint32_t& so, uint32_t& dof) {
    so = 0;
    dof = 0;
#pragma unroll
    for (int k = kMaxDims - 1; k >= 1; --k) {
        if (k < rec.nd) {
            uint32_t q = fdiv(e, rec.ext[k], rec.mul[k], rec.shr[k]);
            uint32_t i = e - q * rec.ext[k];
            so += i * rec.sstr[k];
            dof += i * rec.dstr[k];
            e = q;
        }
    }
    so += e * rec.sstr[0];
    dof += e * rec.dstr[0];
}

// Offsets of row `row` of a record whose innermost dim (index nd-1) is the run: dims [0, nd-1) only.
template <typename REC>
__device__ __forceinline__ void row_offsets_of(const REC& rec, uint32_t row, uint32_t& so, uint32_t& dof) {
    so = 0;
    dof = 0;
    const int last = rec.nd - 1;
#pragma unroll 1
    for (int k = last - 1; k >= 1; --k) {
        const uint32_t ext = rec.ext[k];
        uint32_t q = fdiv(row, ext, rec.mul[k], rec.shr[k]);
        uint32_t i = row - q * ext;
        so += i * rec.sstr[k];
        dof += i * rec.dstr[k];
        row = q;
    }
    if (last >= 1) {
        so += row * rec.sstr[0];
        dof += row * rec.dstr[0];
    }
}

// Warp-cooperative copy of n contiguous elements.  float64: the source is aligned to 16 bytes by peeling one element,
// the body moves as double2 loads (stores are double2 too when the destination has the same parity, two scalar stores
// otherwise); every lane keeps kRunUnroll 16-byte loads in flight.
template <typename T, bool CONJ>
__device__ __forceinline__ void copy_run(const T* __restrict__ s, T* __restrict__ d, uint32_t n, int lane) {
    if constexpr (sizeof(T) == 8) {
        const uint32_t head = min((uint32_t)((reinterpret_cast<uintptr_t>(s) >> 3) & 1), n);
        if (head && lane == 0) d[0] = s[0];
        s += head;
        d += head;
        n -= head;
        const uint32_t npair = n >> 1;
        const double2* __restrict__ s2 = reinterpret_cast<const double2*>(s);
        const bool dvec = (reinterpret_cast<uintptr_t>(d) & 15) == 0;
        for (uint32_t p0 = lane; p0 < npair; p0 += 32 * kRunUnroll) {
            double2 v[kRunUnroll];
#pragma unroll
            for (int u = 0; u < kRunUnroll; ++u) v[u] = load_if<double2, false>(p0 + u * 32 < npair, s2 + p0 + u * 32);
            if (dvec) {
#pragma unroll
                for (int u = 0; u < kRunUnroll; ++u)
                    if (p0 + u * 32 < npair) reinterpret_cast<double2*>(d)[p0 + u * 32] = v[u];
            } else {
#pragma unroll
                for (int u = 0; u < kRunUnroll; ++u)
                    if (p0 + u * 32 < npair) {
                        d[2 * (p0 + u * 32)] = v[u].x;
                        d[2 * (p0 + u * 32) + 1] = v[u].y;
                    }
            }
        }
        if ((n & 1) && lane == 0) d[n - 1] = s[n - 1];
    } else {
        for (uint32_t p0 = lane; p0 < n; p0 += 32 * kRunUnroll) {
            T v[kRunUnroll];
#pragma unroll
            for (int u = 0; u < kRunUnroll; ++u) v[u] = load_if<T, CONJ>(p0 + u * 32 < n, s + p0 + u * 32);
#pragma unroll
            for (int u = 0; u < kRunUnroll; ++u)
                if (p0 + u * 32 < n) d[p0 + u * 32] = v[u];
        }
    }
}

// Flat copy of the destination-linear range [e0, end) of one record: every lane keeps 64 bytes of independent loads
// outstanding before its first store (HBM latency x bandwidth asks for ~35 KB in flight per SM).
template <typename T, bool CONJ, typename REC>
__device__ __forceinline__ void copy_flat(const REC& rec, const T* __restrict__ s, T* __restrict__ d, uint32_t e0, uint32_t end, int lane) {
    constexpr int kUnroll = 64 / sizeof(T);
    const int nd = rec.nd;
    for (uint32_t e = e0 + lane; e < end; e += kUnroll * 32) {
        // index decomposition of kUnroll elements at once, dim by dim: the divisor of a dim is fetched once per trip
        // (keeping all 8 dims x 5 fields live across the unrolled element loop costs 40 registers and spills)
        uint32_t rem[kUnroll], so[kUnroll], dof[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            rem[u] = e + u * 32;
            so[u] = 0;
            dof[u] = 0;
        }
#pragma unroll 1
        for (int k = nd - 1; k >= 1; --k) {
            const uint32_t ext = rec.ext[k], mul = rec.mul[k], shr = rec.shr[k], ss = rec.sstr[k], ds = rec.dstr[k];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const uint32_t q = fdiv(rem[u], ext, mul, shr);
                const uint32_t i = rem[u] - q * ext;
                so[u] += i * ss;
                dof[u] += i * ds;
                rem[u] = q;
            }
        }
        const uint32_t ss0 = rec.sstr[0], ds0 = rec.dstr[0];
        T v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            so[u] += rem[u] * ss0;
            dof[u] += rem[u] * ds0;
            v[u] = load_if<T, CONJ>(e + u * 32 < end, s + so[u]);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (e + u * 32 < end) d[dof[u]] = v[u];
    }
}

// ---- the item kinds; each is its own (non-inlined) function so that the register allocation of one path does not
// ---- spill the others: inlined into one loop body the five paths needed 80 registers plus 0.4 KB of local memory

// short rows: the innermost dim is a run of L >= kRowShortMin elements contiguous on both sides.  A trip moves 256 (128
// for complex128) consecutive elements, i.e. at most 256 / L + 2 <= 18 rows: lane j resolves the outer indices of row
// (first + j) once, every element then costs one division by L and two shuffles instead of a full index decomposition.
template <typename T, bool CONJ>
__device__ __noinline__ void item_rows_short(const CopyRec& rec, const T* __restrict__ s, T* __restrict__ d, uint32_t e0, uint32_t end, int lane) {
    constexpr int kUnroll = 64 / sizeof(T);
    const int last = rec.nd - 1;
    const uint32_t L = rec.ext[last], lmul = rec.mul[last], lshr = rec.shr[last];
    for (uint32_t base = e0; base < end; base += kUnroll * 32) {
        const uint32_t row0 = __umulhi(base, lmul) >> lshr;
        uint32_t so_r, dof_r;
        row_offsets_of(rec, row0 + lane, so_r, dof_r);     // lanes beyond the rows of the trip compute unused values
        T v[kUnroll];
        uint32_t dofs[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint32_t ee = base + lane + u * 32;
            const uint32_t row = __umulhi(ee, lmul) >> lshr;
            const uint32_t col = ee - row * L;
            const int j = (int)(row - row0) & 31;
            const uint32_t so = __shfl_sync(0xffffffffu, so_r, j) + col;
            dofs[u] = __shfl_sync(0xffffffffu, dof_r, j) + col;
            v[u] = load_if<T, CONJ>(ee < end, s + so);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (base + lane + u * 32 < end) d[dofs[u]] = v[u];
    }
}

// long rows: the warp resolves the outer indices once per row and moves the run with 16-byte loads (copy_run)
template <typename T, bool CONJ>
__device__ __noinline__ void item_rows_long(const CopyRec& rec, const T* __restrict__ s, T* __restrict__ d, uint32_t e0, uint32_t end, int lane) {
    const int last = rec.nd - 1;
    const uint32_t L = rec.ext[last];
    uint32_t e = e0;
    while (e < end) {
        const uint32_t row = fdiv(e, L, rec.mul[last], rec.shr[last]);
        const uint32_t col = e - row * L;
        const uint32_t run = min(L - col, end - e);
        uint32_t so, dof;
        row_offsets_of(rec, row, so, dof);
        copy_run<T, CONJ>(s + so + col, d + dof + col, run, lane);
        e += run;
    }
}

// ---- tiled transpose (its own kernel, one CTA per item): slab index -> (outer index, slab coordinates along a and b); the
// slab goes through shared memory so that global reads run along b (source-fast) and global writes along a
// (destination-fast).  Every thread keeps 128 bytes of loads in flight; the four resident CTAs of an SM run out of phase,
// which overlaps the read and write halves well enough: two explicit pipelines were measured SLOWER on the D=16384
// float64 transposes (4.56 TB/s as written; 3.87 with the next slab prefetched into registers across the barriers; 3.40
// with 8/16-byte cp.async into a double-buffered tile at 3 CTAs per SM).
struct SlabPos {
    uint32_t so, dof, a0, b0;
};

template <typename T>
__device__ __forceinline__ SlabPos slab_position(const CopyRec& rec, uint32_t slab) {
    constexpr int kTileB = TileShape<T>::kTileB;
    const int da = rec.tile_a, db = rec.tile_b;
    uint32_t ta = slab % rec.tiles_a;
    uint32_t rest = slab / rec.tiles_a;
    uint32_t tb = rest % rec.tiles_b;
    uint32_t outer = rest / rec.tiles_b;
    SlabPos p;
    p.so = 0;
    p.dof = 0;
#pragma unroll 1
    for (int k = rec.nd - 1; k >= 0; --k) {
        if (k != da && k != db) {
            const uint32_t ext = rec.ext[k];
            uint32_t q = fdiv(outer, ext, rec.mul[k], rec.shr[k]);
            uint32_t i = outer - q * ext;
            p.so += i * rec.sstr[k];
            p.dof += i * rec.dstr[k];
            outer = q;
        }
    }
    p.a0 = ta * kTileA;
    p.b0 = tb * kTileB;
    return p;
}

template <typename T>
struct SlabRegs {
    T v[kTileA / kCopyWarps][TileShape<T>::kTileB / 32];
};

template <typename T, bool CONJ>
__device__ __forceinline__ void slab_load(const CopyRec& rec, const SlabPos& p, const T* __restrict__ s, SlabRegs<T>& r, int tid) {
    constexpr int kTileB = TileShape<T>::kTileB;
    const int lane = tid & 31, warp = tid >> 5;
    const int da = rec.tile_a, db = rec.tile_b;
    const uint32_t ea = rec.ext[da], eb = rec.ext[db], sa_s = rec.sstr[da], sb_s = rec.sstr[db];
#pragma unroll
    for (int j = 0; j < kTileA / kCopyWarps; ++j) {
        const uint32_t ia = p.a0 + j * kCopyWarps + warp;
#pragma unroll
        for (int h = 0; h < kTileB / 32; ++h) {
            const uint32_t ib = p.b0 + h * 32 + lane;
            r.v[j][h] = load_if<T, CONJ>(ia < ea && ib < eb, s + p.so + ia * sa_s + ib * sb_s);
        }
    }
}

template <typename T>
__device__ __forceinline__ void slab_to_tile(const SlabRegs<T>& r, T (*tile)[TileShape<T>::kTileB + 1], int tid) {
    constexpr int kTileB = TileShape<T>::kTileB;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int j = 0; j < kTileA / kCopyWarps; ++j)
#pragma unroll
        for (int h = 0; h < kTileB / 32; ++h) tile[j * kCopyWarps + warp][h * 32 + lane] = r.v[j][h];
}

template <typename T>
__device__ __forceinline__ void slab_store(const CopyRec& rec, const SlabPos& p, T* __restrict__ d, T (*tile)[TileShape<T>::kTileB + 1], int tid) {
    constexpr int kTileB = TileShape<T>::kTileB;
    const int lane = tid & 31, warp = tid >> 5;
    const int da = rec.tile_a, db = rec.tile_b;
    const uint32_t ea = rec.ext[da], eb = rec.ext[db], sa_d = rec.dstr[da], sb_d = rec.dstr[db];
#pragma unroll
    for (int j = 0; j < kTileB / kCopyWarps; ++j) {
        const uint32_t bc = j * kCopyWarps + warp, ib = p.b0 + bc;
#pragma unroll
        for (int h = 0; h < kTileA / 32; ++h) {
            const uint32_t ac = h * 32 + lane, ia = p.a0 + ac;
            if (ia < ea && ib < eb) d[p.dof + ia * sa_d + ib * sb_d] = tile[ac][bc];
        }
    }
}

template <typename T, bool CONJ>
__global__ void __launch_bounds__(kCopyThreads, 4)
tiled_kernel(const CopyRec* __restrict__ recs, const CopyItem* __restrict__ items, int nitems,
             const T* __restrict__ src, T* __restrict__ dst) {
    __shared__ T tile[kTileA][TileShape<T>::kTileB + 1];
    __shared__ CopyRec crec;
    const int tid = threadIdx.x;
    for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
        const CopyItem item = items[it];
        __syncthreads();   // the previous item is done with crec and the tile
        if (tid < kRecWords) reinterpret_cast<uint32_t*>(&crec)[tid] = reinterpret_cast<const uint32_t*>(recs + item.rec_begin)[tid];
        __syncthreads();
        const T* s = src + crec.src_base;
        T* d = dst + crec.dst_base;
        const uint32_t end = item.e0 + item.ne;
        for (uint32_t slab = item.e0; slab < end; ++slab) {
            const SlabPos cur = slab_position<T>(crec, slab);
            SlabRegs<T> regs;
            slab_load<T, CONJ>(crec, cur, s, regs, tid);
            slab_to_tile<T>(regs, tile, tid);
            __syncthreads();
            slab_store<T>(crec, cur, d, tile, tid);
            __syncthreads();
        }
    }
}

template <typename T, bool CONJ>
__device__ __noinline__ void item_flat(const CopyRec& rec, const T* __restrict__ s, T* __restrict__ d, uint32_t e0, uint32_t end, int lane) {
    copy_flat<T, CONJ>(rec, s, d, e0, end, lane);
}

// pack of small records, one after the other; the next record is fetched while the current one moves
template <typename T, bool CONJ>
__device__ __noinline__ void item_pack(const CopyRec* __restrict__ recs, int rec_begin, int rec_end, CopyRec& srec,
                                       const T* __restrict__ src, T* __restrict__ dst, int lane) {
    uint32_t w0, w1 = 0;
    {
        const uint32_t* g = reinterpret_cast<const uint32_t*>(recs + rec_begin);
        w0 = g[lane];
        if (lane + 32 < kRecWords) w1 = g[lane + 32];
    }
    for (int r = rec_begin; r < rec_end; ++r) {
        __syncwarp();
        uint32_t* sm = reinterpret_cast<uint32_t*>(&srec);
        sm[lane] = w0;
        if (lane + 32 < kRecWords) sm[lane + 32] = w1;
        __syncwarp();
        if (r + 1 < rec_end) {
            const uint32_t* g = reinterpret_cast<const uint32_t*>(recs + r + 1);
            w0 = g[lane];
            if (lane + 32 < kRecWords) w1 = g[lane + 32];
        }
        copy_flat<T, CONJ>(srec, src + srec.src_base, dst + srec.dst_base, 0u, srec.total, lane);
    }
}

// batch of up to 32 contiguous runs: lane j fetches run j (one coalesced 1 KB load for the warp), an inclusive scan of the
// lengths lays the runs on one line of `total` elements, and the warp walks that line 32 elements per load instruction —
// every lane keeps a cursor (run index) that only moves forward, so runs of any length keep all lanes and 64 bytes of loads
// per lane busy.  Addresses of a trip are resolved first (branches, shared memory), then all loads issue back to back.
template <typename T, bool CONJ>
__device__ __noinline__ void item_runs(const CopyRun* __restrict__ runs, int begin, int end, RunSmem& sm, const T* __restrict__ src,
                                       T* __restrict__ dst, int lane) {
    constexpr int kUnroll = 64 / sizeof(T);
    const int nrun = end - begin;
    int64_t rs = 0, rd = 0;
    uint32_t rn = 0, rf = 0;
    if (lane < nrun) {
        const uint4* g = reinterpret_cast<const uint4*>(runs + begin + lane);
        const uint4 w0 = g[0], w1 = g[1];
        rs = (int64_t)(((uint64_t)w0.y << 32) | w0.x);
        rd = (int64_t)(((uint64_t)w0.w << 32) | w0.z);
        rn = w1.x;
        rf = w1.y;
    }
    uint32_t e = rn;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, e, d);
        if (lane >= d) e += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, e, 31);
    __syncwarp();      // the previous item of this warp is done with sm
    sm.src[lane] = rs;
    sm.dst[lane] = rd;
    sm.flags[lane] = rf;
    sm.start[lane] = e - rn;
    if (lane == 31) sm.start[kRunBatch] = total;
    __syncwarp();
    int j = 0;
    uint32_t cstart = 0, cend = sm.start[1];
    int64_t csrc = sm.src[0], cdst = sm.dst[0];
    bool czero = sm.flags[0] & 1;
    for (uint32_t base = 0; base < total; base += 32 * kUnroll) {
        int64_t so[kUnroll], dof[kUnroll];
        uint32_t live = 0, fill = 0;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint32_t p = base + 32 * u + lane;
            so[u] = 0;
            dof[u] = 0;
            if (p < total) {
                if (p >= cend) {
                    do {
                        ++j;
                        cend = sm.start[j + 1];
                    } while (p >= cend);
                    cstart = sm.start[j];
                    csrc = sm.src[j];
                    cdst = sm.dst[j];
                    czero = sm.flags[j] & 1;
                }
                so[u] = csrc + (p - cstart);
                dof[u] = cdst + (p - cstart);
                live |= 1u << u;
                if (czero) fill |= 1u << u;
            }
        }
        T v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) v[u] = load_if<T, CONJ>(((live & ~fill) >> u) & 1, src + so[u]);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if ((live >> u) & 1) dst[dof[u]] = v[u];
    }
}

// One warp = one work item at a time; warps never synchronise with each other, so an SM keeps 32 independent
// item -> record -> loads -> stores chains in flight (a CTA-wide item pipeline kept 3 and starved on tensors made of
// thousands of small blocks).  Transposing slabs are a separate launch (tiled_kernel).
template <typename T, bool CONJ>
__global__ void __launch_bounds__(kCopyThreads, 4)
copy_kernel(const CopyRec* __restrict__ recs, const CopyItem* __restrict__ items, int nitems, const CopyRun* __restrict__ runs,
            const T* __restrict__ src, T* __restrict__ dst) {
    __shared__ CopyRec srecs[kCopyWarps];
    __shared__ RunSmem sruns[kCopyWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    CopyRec& srec = srecs[warp];
    const int nwarps = gridDim.x * kCopyWarps;
    for (int it = blockIdx.x * kCopyWarps + warp; it < nitems; it += nwarps) {
        const CopyItem item = items[it];
        if (item.kind == 5) {
            item_runs<T, CONJ>(runs, item.rec_begin, item.rec_end, sruns[warp], src, dst, lane);
            continue;
        }
        if (item.kind == 1) {
            item_pack<T, CONJ>(recs, item.rec_begin, item.rec_end, srec, src, dst, lane);
            continue;
        }
        __syncwarp();   // the previous item is done with srec
        {
            const uint32_t* g = reinterpret_cast<const uint32_t*>(recs + item.rec_begin);
            uint32_t* sm = reinterpret_cast<uint32_t*>(&srec);
            sm[lane] = g[lane];
            if (lane + 32 < kRecWords) sm[lane + 32] = g[lane + 32];
        }
        __syncwarp();
        const T* s = src + srec.src_base;
        T* d = dst + srec.dst_base;
        const uint32_t end = item.e0 + item.ne;
        if (item.kind == 0)
            item_flat<T, CONJ>(srec, s, d, item.e0, end, lane);
        else if (item.kind == 4)
            item_rows_short<T, CONJ>(srec, s, d, item.e0, end, lane);
        else
            item_rows_long<T, CONJ>(srec, s, d, item.e0, end, lane);
    }
}

}  // namespace yb

using namespace yb;

struct yb_copy_plan {
    int itemsize = 0, device = 0;
    int nitems = 0, nwarp_items = 0;
    int64_t elems = 0, nrecs = 0, ntiled = 0;
    DeviceTable recs, items, runs;
    int64_t nruns = 0;
    int grid = 0, grid_tiled = 0;
};

namespace {

// Drop unit dims, order by destination stride (descending), merge dims that are contiguous on both sides.
bool normalise(HostRec& r) {
    int keep[kHostDims], nk = 0;
    for (int k = 0; k < r.nd; ++k) {
        if (r.ext[k] == 0) return false;
        if (r.ext[k] != 1) keep[nk++] = k;
    }
    // stable insertion sort by destination stride, descending (nk is tiny)
    for (int i = 1; i < nk; ++i) {
        const int v = keep[i];
        int j = i - 1;
        while (j >= 0 && r.dstr[keep[j]] < r.dstr[v]) {
            keep[j + 1] = keep[j];
            --j;
        }
        keep[j + 1] = v;
    }
    int64_t e[kHostDims], s[kHostDims], d[kHostDims];
    int n = 0;
    for (int i = 0; i < nk; ++i) {
        const int k = keep[i];
        if (n > 0 && s[n - 1] == r.sstr[k] * r.ext[k] && d[n - 1] == r.dstr[k] * r.ext[k]) {
            e[n - 1] *= r.ext[k];
            s[n - 1] = r.sstr[k];
            d[n - 1] = r.dstr[k];
        } else {
            e[n] = r.ext[k];
            s[n] = r.sstr[k];
            d[n] = r.dstr[k];
            ++n;
        }
    }
    if (n == 0) {  // single element
        e[0] = s[0] = d[0] = 1;
        n = 1;
    }
    r.nd = n;
    for (int k = 0; k < n; ++k) {
        r.ext[k] = e[k];
        r.sstr[k] = s[k];
        r.dstr[k] = d[k];
    }
    return true;
}

void drop_front(HostRec& r) {
    for (int k = 1; k < r.nd; ++k) {
        r.ext[k - 1] = r.ext[k];
        r.sstr[k - 1] = r.sstr[k];
        r.dstr[k - 1] = r.dstr[k];
    }
    if (--r.nd == 0) {
        r.ext[0] = r.sstr[0] = r.dstr[0] = 1;
        r.nd = 1;
    }
}

// Split records that violate device limits (rank, 31-bit element counts / relative offsets).
void split_to_limits(const HostRec& r, std::vector<HostRec>& out) {
    int64_t total = 1, smax = 0, dmax = 0;
    for (int k = 0; k < r.nd; ++k) {
        total *= r.ext[k];
        smax += (r.ext[k] - 1) * r.sstr[k];
        dmax += (r.ext[k] - 1) * r.dstr[k];
    }
    const int64_t lim = (1ll << 31) - 1;
    if (r.nd <= kMaxDims && total <= lim && smax <= lim && dmax <= lim) {
        out.push_back(r);
        return;
    }
    // peel the outermost dim: either one index at a time (rank too high) or in two halves (too large)
    const int64_t e0 = r.ext[0];
    if (r.nd > kMaxDims || e0 == 1) {
        for (int64_t i = 0; i < e0; ++i) {
            HostRec sub = r;
            sub.src_base += i * r.sstr[0];
            sub.dst_base += i * r.dstr[0];
            drop_front(sub);
            split_to_limits(sub, out);
        }
        return;
    }
    HostRec lo = r, hi = r;
    lo.ext[0] = e0 / 2;
    hi.ext[0] = e0 - e0 / 2;
    hi.src_base += lo.ext[0] * r.sstr[0];
    hi.dst_base += lo.ext[0] * r.dstr[0];
    split_to_limits(lo, out);
    split_to_limits(hi, out);
}

int sm_count_of(int device) {
    static int sm_count[64] = {0};
    int sms = (device >= 0 && device < 64) ? sm_count[device] : 0;
    if (sms == 0) {
        sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (device >= 0 && device < 64) sm_count[device] = sms;
    }
    return sms;
}

// Elements per work item.  A warp walks its item trip by trip (256 float64 / 128 complex128 elements per trip, each trip
// paying the full HBM latency), so a plan needs about two items per resident warp (4 CTAs x 8 warps per SM) before longer
// items pay off: with fixed 4096-element items a 6 MB merge occupied 25 SMs for 16 dependent trips each.
uint32_t item_size(int64_t elems, int sms, int itemsize) {
    const int64_t trip = 32 * (64 / itemsize);
    const int64_t want = elems / (2ll * sms * 4 * kCopyWarps);
    const int64_t sz = std::max<int64_t>(trip, std::min<int64_t>(kItemElems, (want + trip - 1) / trip * trip));
    return (uint32_t)sz;
}

}  // namespace

extern "C" int yb_copy_plan_create(const int64_t* recs, int64_t nrec, int rank, int itemsize, int device,
                                   yb_copy_plan** out) {
    if (!out) return fail(kErrArg, "yb_copy_plan_create: out is null");
    *out = nullptr;
    if (nrec < 0 || rank < 0 || (nrec > 0 && !recs)) return fail(kErrArg, "yb_copy_plan_create: bad table");
    if (rank > kHostDims) return fail(kErrUnsupported, "yb_copy_plan_create: rank %d above %d", rank, kHostDims);
    if (itemsize != 8 && itemsize != 16) return fail(kErrUnsupported, "yb_copy_plan_create: itemsize %d (8 or 16)", itemsize);

    std::vector<HostRec> host;
    host.reserve((size_t)nrec);
    std::vector<CopyRun> runs;
    int64_t zero_elems = 0;
    const int64_t w = 2 + 3 * (int64_t)rank;
    for (int64_t i = 0; i < nrec; ++i) {
        const int64_t* p = recs + i * w;
        HostRec r;
        r.src_base = p[0];
        r.dst_base = p[1];
        r.zero = p[0] == YB_COPY_SRC_ZERO;
        r.nd = rank;
        for (int k = 0; k < rank; ++k) {
            r.ext[k] = p[2 + k];
            r.sstr[k] = p[2 + rank + k];
            r.dstr[k] = p[2 + 2 * rank + k];
        }
        if (r.zero) {   // no source: give the box the strides of its destination so that contiguous dims merge
            r.src_base = 0;
            for (int k = 0; k < rank; ++k) r.sstr[k] = r.dstr[k];
        }
        for (int k = 0; k < rank; ++k)
            if (r.ext[k] < 0 || r.sstr[k] < 0 || r.dstr[k] < 0) return fail(kErrArg, "yb_copy_plan_create: negative extent/stride in record %lld", (long long)i);
        if (!normalise(r)) continue;
        if (r.zero) {   // straight to the run table: one run per contiguous row (piece) of the box
            const bool inner = r.dstr[r.nd - 1] == 1;
            const int outer = inner ? r.nd - 1 : r.nd;
            const int64_t L = inner ? r.ext[r.nd - 1] : 1;
            int64_t idx[kHostDims] = {0}, rows = 1;
            for (int k = 0; k < outer; ++k) rows *= r.ext[k];
            for (int64_t q = 0; q < rows; ++q) {
                int64_t dof = r.dst_base;
                for (int k = 0; k < outer; ++k) dof += idx[k] * r.dstr[k];
                for (int64_t c = 0; c < L; c += kRunPiece) {
                    CopyRun run = {0, dof + c, (uint32_t)std::min<int64_t>(kRunPiece, L - c), 1u, {0, 0}};
                    runs.push_back(run);
                }
                for (int k = outer - 1; k >= 0 && ++idx[k] == r.ext[k]; --k) idx[k] = 0;
            }
            zero_elems += rows * L;
            continue;
        }
        split_to_limits(r, host);
    }

    const int tile_b = itemsize == 8 ? TileShape<double>::kTileB : TileShape<double2>::kTileB;
    const int sms = sm_count_of(device);
    int64_t all_elems = 0;
    for (const HostRec& h : host) {
        int64_t t = 1;
        for (int k = 0; k < h.nd; ++k) t *= h.ext[k];
        all_elems += t;
    }
    const uint32_t item_elems = item_size(all_elems, sms, itemsize);
    std::vector<CopyRec> drecs;
    std::vector<CopyItem> items, cta_items;
    std::vector<char> on_runs(host.size(), 0);
    int64_t elems = zero_elems, ntiled = 0, nslabs = 0;
    drecs.reserve(host.size());
    // Records that are a few contiguous rows go to the run table (see CopyRun): everything the kernel needs to move a run
    // is in its 32-byte entry.  Zero-fill boxes always do (they have no record path).
    for (size_t i = 0; i < host.size(); ++i) {
        const HostRec& h = host[i];
        const int last = h.nd - 1;
        if (h.nd > 2 || h.sstr[last] != 1 || h.dstr[last] != 1) continue;
        const int64_t L = h.ext[last], rows = h.nd == 2 ? h.ext[0] : 1;
        if (L * rows > kRunRecMax || (h.nd == 2 && L < kRunRowMin)) continue;
        on_runs[i] = 1;
        for (int64_t r = 0; r < rows; ++r) {
            const int64_t so = h.src_base + (h.nd == 2 ? r * h.sstr[0] : 0), dof = h.dst_base + (h.nd == 2 ? r * h.dstr[0] : 0);
            for (int64_t c = 0; c < L; c += item_elems) {     // pieces of one work item's size: small tensors stay spread out
                CopyRun run;
                run.src = so + c;
                run.dst = dof + c;
                run.n = (uint32_t)std::min<int64_t>(item_elems, L - c);
                run.flags = 0u;
                run.pad_[0] = run.pad_[1] = 0;
                runs.push_back(run);
            }
        }
        elems += L * rows;
    }
    {   // batches: up to 32 runs and about item_elems elements each
        size_t b = 0;
        while (b < runs.size()) {
            size_t e = b;
            uint64_t acc = 0;
            while (e < runs.size() && e - b < (size_t)kRunBatch && (acc == 0 || acc + runs[e].n <= item_elems)) acc += runs[e++].n;
            CopyItem it = {(int32_t)b, (int32_t)e, 0, 0, 5, {0, 0, 0}};
            items.push_back(it);
            b = e;
        }
    }
    // large records first in table order; tiny ones are packed afterwards
    std::vector<int> small_idx;
    for (size_t i = 0; i < host.size(); ++i) {
        const HostRec& h = host[i];
        CopyRec c;
        memset(&c, 0, sizeof(c));
        if (on_runs[i]) {          // keeps the record indices aligned with `host`; never referenced by an item
            c.nd = 1;
            c.tile_a = c.tile_b = -1;
            drecs.push_back(c);
            continue;
        }
        c.src_base = h.src_base;
        c.dst_base = h.dst_base;
        c.nd = h.nd;
        int64_t total = 1;
        for (int k = 0; k < kMaxDims; ++k) {
            c.ext[k] = 1;
            c.mul[k] = 0;
            c.shr[k] = 0;
        }
        for (int k = 0; k < c.nd; ++k) {
            FastDiv f = make_fastdiv((uint32_t)h.ext[k]);
            c.ext[k] = f.div;
            c.mul[k] = f.mul;
            c.shr[k] = f.shr;
            c.sstr[k] = (uint32_t)h.sstr[k];
            c.dstr[k] = (uint32_t)h.dstr[k];
            total *= h.ext[k];
        }
        c.total = (uint32_t)total;
        c.tile_a = c.tile_b = -1;
        elems += total;
        // tiled path: destination-fast dim (last) differs from the source-fast dim and both are wide
        int da = c.nd - 1, db = 0;
        for (int k = 1; k < c.nd; ++k)
            if (h.sstr[k] < h.sstr[db]) db = k;
        if (c.nd >= 2 && da != db && h.ext[da] >= 16 && h.ext[db] >= 16 && total >= 4096) {
            c.tile_a = da;
            c.tile_b = db;
            c.tiles_a = (uint32_t)((h.ext[da] + kTileA - 1) / kTileA);
            c.tiles_b = (uint32_t)((h.ext[db] + tile_b - 1) / tile_b);
            c.outer_total = (uint32_t)(total / (h.ext[da] * h.ext[db]));
            nslabs += (int64_t)c.tiles_a * c.tiles_b * c.outer_total;
            ++ntiled;
        }
        drecs.push_back(c);
    }
    // slabs per CTA item: about two items per resident CTA before items grow (same reasoning as item_size)
    const uint32_t per_item = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(4, nslabs / (2ll * sms * 4)));
    for (size_t i = 0; i < drecs.size(); ++i) {
        const CopyRec& c = drecs[i];
        if (on_runs[i]) continue;
        if (c.tile_a >= 0) {
            const uint64_t nslab = (uint64_t)c.tiles_a * c.tiles_b * c.outer_total;
            for (uint64_t s0 = 0; s0 < nslab; s0 += per_item) {
                CopyItem it = {(int32_t)i, (int32_t)i + 1, (uint32_t)s0, (uint32_t)std::min<uint64_t>(per_item, nslab - s0), 2, {0, 0, 0}};
                cta_items.push_back(it);
            }
        } else if (c.total >= kSmallRec) {
            const int last = c.nd - 1;
            const bool inner = c.sstr[last] == 1 && c.dstr[last] == 1;   // innermost dim is a run contiguous on both sides
            int kind = 0;
            if (inner && itemsize == 8 && c.ext[last] >= kRowMin) kind = 3;
            else if (inner && c.nd >= 2 && c.ext[last] >= kRowShortMin) kind = 4;
            for (uint64_t e0 = 0; e0 < c.total; e0 += item_elems) {
                CopyItem it = {(int32_t)i, (int32_t)i + 1, (uint32_t)e0, (uint32_t)std::min<uint64_t>(item_elems, c.total - e0), kind, {0, 0, 0}};
                items.push_back(it);
            }
        } else {
            small_idx.push_back((int)i);
        }
    }
    // packs need consecutive record indices: append re-ordered copies of the small records at the end
    if (!small_idx.empty()) {
        const int base = (int)drecs.size();
        for (int idx : small_idx) drecs.push_back(drecs[idx]);
        int begin = base;
        uint64_t acc = 0;
        for (int j = 0; j < (int)small_idx.size(); ++j) {
            acc += drecs[base + j].total;
            const bool last = (j + 1 == (int)small_idx.size());
            if (acc >= item_elems || (base + j + 1 - begin) >= kPackMaxRecs || last) {
                CopyItem it = {begin, base + j + 1, 0, 0, 1, {0, 0, 0}};
                items.push_back(it);
                begin = base + j + 1;
                acc = 0;
            }
        }
    }

    const int nwarp_items = (int)items.size();
    items.insert(items.end(), cta_items.begin(), cta_items.end());   // phase 2 of the kernel: one CTA per item

    yb_copy_plan* plan = new yb_copy_plan();
    plan->itemsize = itemsize;
    plan->device = device;
    plan->nitems = (int)items.size();
    plan->nwarp_items = nwarp_items;
    plan->elems = elems;
    plan->nrecs = (int64_t)host.size();
    plan->ntiled = ntiled;
    plan->nruns = (int64_t)runs.size();
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = kOk;
    if (cudaSetDevice(device) != cudaSuccess) rc = fail(kErrCuda, "yb_copy_plan_create: cudaSetDevice(%d) failed", device);
    if (rc == kOk) {
        TableBatch up;
        up.add(plan->recs, drecs.data(), drecs.size() * sizeof(CopyRec));
        up.add(plan->items, items.data(), items.size() * sizeof(CopyItem));
        up.add(plan->runs, runs.data(), runs.size() * sizeof(CopyRun));
        rc = up.commit();
    }
    if (rc == kOk) {
        plan->grid = std::min((nwarp_items + kCopyWarps - 1) / kCopyWarps, sms * 4);
        plan->grid_tiled = std::min((int)cta_items.size(), sms * 4);
    }
    cudaSetDevice(prev);
    if (rc != kOk) {
        plan->recs.release();
        plan->items.release();
        plan->runs.release();
        delete plan;
        return rc;
    }
    *out = plan;
    return kOk;
}

extern "C" int yb_copy_plan_info(const yb_copy_plan* plan, int64_t info[5]) {
    if (!plan || !info) return fail(kErrArg, "yb_copy_plan_info: null argument");
    info[0] = plan->nitems;
    info[1] = plan->elems;
    info[2] = plan->nrecs;
    info[3] = plan->ntiled;
    info[4] = plan->nruns;
    return kOk;
}

namespace {
template <typename T, bool CONJ>
int launch_copy(const yb_copy_plan* plan, const CopyRec* recs, const CopyItem* items, const CopyItem* titems, int ntiled_items,
                const T* src, T* dst, cudaStream_t st) {
    if (plan->nwarp_items > 0)
        copy_kernel<T, CONJ><<<plan->grid, kCopyThreads, 0, st>>>(recs, items, plan->nwarp_items, (const CopyRun*)plan->runs.ptr, src, dst);
    if (ntiled_items > 0) {
        tiled_kernel<T, CONJ><<<plan->grid_tiled, kCopyThreads, 0, st>>>(recs, titems, ntiled_items, src, dst);
    }
    return kOk;
}
}  // namespace

extern "C" int yb_copy_run(const yb_copy_plan* plan, const void* src, void* dst, int64_t dst_elems, int flags, void* stream) {
    if (!plan) return fail(kErrArg, "yb_copy_run: plan is null");
    cudaStream_t st = (cudaStream_t)stream;
    if ((flags & YB_COPY_ZERO_DST) && dst_elems > 0) {
        if (!dst) return fail(kErrArg, "yb_copy_run: dst is null");
        YB_CUDA(cudaMemsetAsync(dst, 0, (size_t)dst_elems * plan->itemsize, st));
    }
    if (plan->nitems == 0) return kOk;
    if (!src || !dst) return fail(kErrArg, "yb_copy_run: null data pointer");
    const bool conj = (flags & YB_COPY_CONJ) != 0;
    if (conj && plan->itemsize != 16) return fail(kErrArg, "yb_copy_run: conj needs a complex plan");
    int rc_launch = kOk;
    const CopyRec* recs = (const CopyRec*)plan->recs.ptr;
    const CopyItem* items = (const CopyItem*)plan->items.ptr;
    const CopyItem* titems = items + plan->nwarp_items;
    const int ntiled_items = plan->nitems - plan->nwarp_items;
    if (plan->itemsize == 8)
        rc_launch = launch_copy<double, false>(plan, recs, items, titems, ntiled_items, (const double*)src, (double*)dst, st);
    else if (conj)
        rc_launch = launch_copy<double2, true>(plan, recs, items, titems, ntiled_items, (const double2*)src, (double2*)dst, st);
    else
        rc_launch = launch_copy<double2, false>(plan, recs, items, titems, ntiled_items, (const double2*)src, (double2*)dst, st);
    if (rc_launch != kOk) return rc_launch;
    YB_CUDA(cudaGetLastError());
    return kOk;
}

extern "C" void yb_copy_plan_destroy(yb_copy_plan* plan) {
    if (!plan) return;
    plan->recs.release();
    plan->items.release();
    plan->runs.release();
    delete plan;
}
