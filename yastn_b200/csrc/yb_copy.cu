// Block copy engine: one launch moves every block of a symmetric tensor between two layouts.
//
// Serves backend.transpose_and_merge / unmerge / transpose and their adjoints (reference loops:
// yastn/backend/_backend_torch_backwards.py:315-319, 340-364, 397-408).  The reference issues one
// strided-copy launch per block; here the host normalises every block move into a "record"
// (dims sorted by destination stride, unit dims dropped, mergeable dims coalesced) and the device
// walks a work-item table:
//   * flat items   — a slice of one large record, destination-linear, fully coalesced writes; reads are
//                    coalesced too whenever the record keeps its innermost source dim innermost;
//   * pack items   — many tiny records, one warp per record (tiny blocks dominate block counts);
//   * tiled items  — (src-fast dim != dst-fast dim) a 2-d slab staged through shared memory so that both
//                    the global reads and the global writes are coalesced.
// HBM-bound: algorithmic bytes = itemsize * (elements read + elements written).
#include <algorithm>
#include <numeric>

#include "yb_common.h"

namespace yb {

constexpr int kMaxDims = 8;
constexpr int kCopyThreads = 256;
constexpr uint32_t kItemElems = 8192;      // elements per flat work item
constexpr uint32_t kPackElems = 4096;      // target elements per pack of small records
constexpr int kPackMaxRecs = 256;
constexpr int kTileDim = 32;               // tiled path: slab is kTileDim x kTileDim elements
constexpr int kRunUnroll = 4;              // row path: 16-byte loads in flight per lane
constexpr uint32_t kRowChunk = 32 * kRunUnroll * 2;   // row path: float64 elements a warp moves per trip
constexpr uint32_t kRowMin = 2048;         // row path only for long contiguous runs (short rows leave lanes idle: measured slower)

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

struct CopyItem {
    int32_t rec_begin, rec_end;  // [rec_begin, rec_end) ; a single record unless this is a pack
    uint32_t e0, ne;             // flat: element range; tiled: slab range
    int32_t kind;                // 0 flat, 1 pack, 2 tiled, 3 rows (flat range, innermost dim contiguous on both sides)
    int32_t pad_[3];
};

constexpr int kHostDims = 24;   // rank limit of an input record (YASTN tensors have at most ~12 legs)
struct HostRec {                 // fixed arrays: plan construction handles thousands of records, no per-record heap traffic
    int64_t src_base, dst_base;
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
            for (int u = 0; u < kRunUnroll; ++u)
                if (p0 + u * 32 < npair) v[u] = s2[p0 + u * 32];
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
            for (int u = 0; u < kRunUnroll; ++u)
                if (p0 + u * 32 < n) v[u] = load_elem<T, CONJ>(s + p0 + u * 32);
#pragma unroll
            for (int u = 0; u < kRunUnroll; ++u)
                if (p0 + u * 32 < n) d[p0 + u * 32] = v[u];
        }
    }
}

template <typename T, bool CONJ>
__global__ void __launch_bounds__(kCopyThreads, 3)
copy_kernel(const CopyRec* __restrict__ recs, const CopyItem* __restrict__ items, int nitems,
            const T* __restrict__ src, T* __restrict__ dst) {
    __shared__ CopyRec srec;
    __shared__ T tile[kTileDim][kTileDim + 1];
    const int tid = threadIdx.x;
    for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
        const CopyItem item = items[it];
        if (item.kind == 1) {
            // pack of tiny records: one warp per record
            const int warp = tid >> 5, lane = tid & 31;
            for (int r = item.rec_begin + warp; r < item.rec_end; r += kCopyThreads / 32) {
                const CopyRec& rec = recs[r];
                const T* s = src + rec.src_base;
                T* d = dst + rec.dst_base;
                const uint32_t total = rec.total;
                for (uint32_t e = lane; e < total; e += 32) {
                    uint32_t so, dof;
                    offsets_of(rec, e, so, dof);
                    d[dof] = load_elem<T, CONJ>(s + so);
                }
            }
            continue;
        }
        __syncthreads();  // previous item done with srec / tile
        {
            const uint32_t* g = reinterpret_cast<const uint32_t*>(recs + item.rec_begin);
            uint32_t* sm = reinterpret_cast<uint32_t*>(&srec);
            for (int w = tid; w < (int)(sizeof(CopyRec) / 4); w += kCopyThreads) sm[w] = g[w];
        }
        __syncthreads();
        const T* s = src + srec.src_base;
        T* d = dst + srec.dst_base;
        if (item.kind == 0) {
            // HBM latency x bandwidth needs ~35 KB in flight per SM: every thread keeps kUnroll independent loads
            // outstanding before the first store (one element per thread and trip reaches only half of the bandwidth)
            constexpr int kUnroll = 64 / sizeof(T);
            const uint32_t end = item.e0 + item.ne;
            for (uint32_t e = item.e0 + tid; e < end; e += kUnroll * kCopyThreads) {
                T v[kUnroll];
                uint32_t dofs[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {
                    const uint32_t ee = e + u * kCopyThreads;
                    if (ee < end) {
                        uint32_t so;
                        offsets_of(srec, ee, so, dofs[u]);
                        v[u] = load_elem<T, CONJ>(s + so);
                    }
                }
#pragma unroll
                for (int u = 0; u < kUnroll; ++u)
                    if (e + u * kCopyThreads < end) d[dofs[u]] = v[u];
            }
        } else if (sizeof(T) == 8 && item.kind == 3) {
            // rows (float64 only; complex128 elements are 16 bytes already and its flat path runs at 92 % of the HBM peak): the innermost dim is a long run contiguous on both sides.  A warp owns a chunk of the destination-linear
            // range, resolves the outer indices once per row and moves the run with 16-byte loads (copy_run).
            const int warp = tid >> 5, lane = tid & 31;
            const int last = srec.nd - 1;
            const uint32_t L = srec.ext[last];
            const uint32_t end = item.e0 + item.ne;
            for (uint32_t cb = item.e0 + warp * kRowChunk; cb < end; cb += (kCopyThreads / 32) * kRowChunk) {
                const uint32_t ce = min(cb + kRowChunk, end);
                uint32_t e = cb;
                while (e < ce) {
                    uint32_t row = fdiv(e, L, srec.mul[last], srec.shr[last]);
                    const uint32_t col = e - row * L;
                    const uint32_t run = min(L - col, ce - e);
                    uint32_t so = col, dof = col;
                    for (int k = last - 1; k >= 1; --k) {
                        uint32_t q = fdiv(row, srec.ext[k], srec.mul[k], srec.shr[k]);
                        uint32_t i = row - q * srec.ext[k];
                        so += i * srec.sstr[k];
                        dof += i * srec.dstr[k];
                        row = q;
                    }
                    if (last >= 1) {
                        so += row * srec.sstr[0];
                        dof += row * srec.dstr[0];
                    }
                    copy_run<T, CONJ>(s + so, d + dof, run, lane);
                    e += run;
                }
            }
        } else {
            // tiled transpose: slab index -> (outer index, slab coordinates along a and b)
            const int da = srec.tile_a, db = srec.tile_b;
            const uint32_t ea = srec.ext[da], eb = srec.ext[db];
            const uint32_t sa_s = srec.sstr[da], sa_d = srec.dstr[da];
            const uint32_t sb_s = srec.sstr[db], sb_d = srec.dstr[db];
            const int tx = tid & 31, ty = tid >> 5;  // 32 x 8
            for (uint32_t slab = item.e0; slab < item.e0 + item.ne; ++slab) {
                uint32_t ta = slab % srec.tiles_a;
                uint32_t rest = slab / srec.tiles_a;
                uint32_t tb = rest % srec.tiles_b;
                uint32_t outer = rest / srec.tiles_b;
                // outer index over the remaining dims (dst order, skipping da/db)
                uint32_t so = 0, dof = 0;
#pragma unroll
                for (int k = kMaxDims - 1; k >= 0; --k) {
                    if (k < srec.nd && k != da && k != db) {
                        uint32_t q = fdiv(outer, srec.ext[k], srec.mul[k], srec.shr[k]);
                        uint32_t i = outer - q * srec.ext[k];
                        so += i * srec.sstr[k];
                        dof += i * srec.dstr[k];
                        outer = q;
                    }
                }
                const uint32_t a0 = ta * kTileDim, b0 = tb * kTileDim;
                // read: threads contiguous along b (source-fast)
#pragma unroll
                for (int j = 0; j < kTileDim; j += 8) {
                    uint32_t ia = a0 + ty + j, ib = b0 + tx;
                    if (ia < ea && ib < eb) tile[ty + j][tx] = load_elem<T, CONJ>(s + so + ia * sa_s + ib * sb_s);
                }
                __syncthreads();
                // write: threads contiguous along a (destination-fast)
#pragma unroll
                for (int j = 0; j < kTileDim; j += 8) {
                    uint32_t ib = b0 + ty + j, ia = a0 + tx;
                    if (ia < ea && ib < eb) d[dof + ia * sa_d + ib * sb_d] = tile[tx][ty + j];
                }
                __syncthreads();
            }
        }
    }
}

}  // namespace yb

using namespace yb;

struct yb_copy_plan {
    int itemsize = 0, device = 0;
    int nitems = 0;
    int64_t elems = 0, nrecs = 0, ntiled = 0;
    DeviceTable recs, items;
    int grid = 0;
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
    const int64_t w = 2 + 3 * (int64_t)rank;
    for (int64_t i = 0; i < nrec; ++i) {
        const int64_t* p = recs + i * w;
        HostRec r;
        r.src_base = p[0];
        r.dst_base = p[1];
        r.nd = rank;
        for (int k = 0; k < rank; ++k) {
            r.ext[k] = p[2 + k];
            r.sstr[k] = p[2 + rank + k];
            r.dstr[k] = p[2 + 2 * rank + k];
        }
        for (int k = 0; k < rank; ++k)
            if (r.ext[k] < 0 || r.sstr[k] < 0 || r.dstr[k] < 0) return fail(kErrArg, "yb_copy_plan_create: negative extent/stride in record %lld", (long long)i);
        if (!normalise(r)) continue;
        split_to_limits(r, host);
    }

    std::vector<CopyRec> drecs;
    std::vector<CopyItem> items;
    int64_t elems = 0, ntiled = 0;
    drecs.reserve(host.size());
    // large records first in table order; tiny ones are packed afterwards
    std::vector<int> small_idx;
    for (size_t i = 0; i < host.size(); ++i) {
        const HostRec& h = host[i];
        CopyRec c;
        memset(&c, 0, sizeof(c));
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
        if (c.nd >= 2 && da != db && h.ext[da] >= 8 && h.ext[db] >= 8 && total >= 1024) {
            c.tile_a = da;
            c.tile_b = db;
            c.tiles_a = (uint32_t)((h.ext[da] + kTileDim - 1) / kTileDim);
            c.tiles_b = (uint32_t)((h.ext[db] + kTileDim - 1) / kTileDim);
            c.outer_total = (uint32_t)(total / (h.ext[da] * h.ext[db]));
            ++ntiled;
        }
        drecs.push_back(c);
    }
    for (size_t i = 0; i < drecs.size(); ++i) {
        const CopyRec& c = drecs[i];
        if (c.tile_a >= 0) {
            const uint64_t nslab = (uint64_t)c.tiles_a * c.tiles_b * c.outer_total;
            const uint32_t per_item = 8;  // 8 slabs of 32x32 = 8192 elements
            for (uint64_t s0 = 0; s0 < nslab; s0 += per_item) {
                CopyItem it = {(int32_t)i, (int32_t)i + 1, (uint32_t)s0, (uint32_t)std::min<uint64_t>(per_item, nslab - s0), 2, {0, 0, 0}};
                items.push_back(it);
            }
        } else if (c.total >= kPackElems / 4) {
            const int last = c.nd - 1;
            const bool rows = itemsize == 8 && c.sstr[last] == 1 && c.dstr[last] == 1 && c.ext[last] >= kRowMin;
            for (uint64_t e0 = 0; e0 < c.total; e0 += kItemElems) {
                CopyItem it = {(int32_t)i, (int32_t)i + 1, (uint32_t)e0, (uint32_t)std::min<uint64_t>(kItemElems, c.total - e0), rows ? 3 : 0, {0, 0, 0}};
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
            if (acc >= kPackElems || (base + j + 1 - begin) >= kPackMaxRecs || last) {
                CopyItem it = {begin, base + j + 1, 0, 0, 1, {0, 0, 0}};
                items.push_back(it);
                begin = base + j + 1;
                acc = 0;
            }
        }
    }

    yb_copy_plan* plan = new yb_copy_plan();
    plan->itemsize = itemsize;
    plan->device = device;
    plan->nitems = (int)items.size();
    plan->elems = elems;
    plan->nrecs = (int64_t)host.size();
    plan->ntiled = ntiled;
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = kOk;
    if (cudaSetDevice(device) != cudaSuccess) rc = fail(kErrCuda, "yb_copy_plan_create: cudaSetDevice(%d) failed", device);
    if (rc == kOk) rc = plan->recs.upload(drecs.data(), drecs.size() * sizeof(CopyRec));
    if (rc == kOk) rc = plan->items.upload(items.data(), items.size() * sizeof(CopyItem));
    if (rc == kOk) {
        static int sm_count[64] = {0};
        int sms = (device >= 0 && device < 64) ? sm_count[device] : 0;
        if (sms == 0) {
            sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
            if (device >= 0 && device < 64) sm_count[device] = sms;
        }
        plan->grid = std::max(1, std::min(plan->nitems, sms * 8));
    }
    cudaSetDevice(prev);
    if (rc != kOk) {
        plan->recs.release();
        plan->items.release();
        delete plan;
        return rc;
    }
    *out = plan;
    return kOk;
}

extern "C" int yb_copy_plan_info(const yb_copy_plan* plan, int64_t info[4]) {
    if (!plan || !info) return fail(kErrArg, "yb_copy_plan_info: null argument");
    info[0] = plan->nitems;
    info[1] = plan->elems;
    info[2] = plan->nrecs;
    info[3] = plan->ntiled;
    return kOk;
}

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
    const CopyRec* recs = (const CopyRec*)plan->recs.ptr;
    const CopyItem* items = (const CopyItem*)plan->items.ptr;
    if (plan->itemsize == 8)
        copy_kernel<double, false><<<plan->grid, kCopyThreads, 0, st>>>(recs, items, plan->nitems, (const double*)src, (double*)dst);
    else if (conj)
        copy_kernel<double2, true><<<plan->grid, kCopyThreads, 0, st>>>(recs, items, plan->nitems, (const double2*)src, (double2*)dst);
    else
        copy_kernel<double2, false><<<plan->grid, kCopyThreads, 0, st>>>(recs, items, plan->nitems, (const double2*)src, (double2*)dst);
    YB_CUDA(cudaGetLastError());
    return kOk;
}

extern "C" void yb_copy_plan_destroy(yb_copy_plan* plan) {
    if (!plan) return;
    plan->recs.release();
    plan->items.release();
    delete plan;
}
