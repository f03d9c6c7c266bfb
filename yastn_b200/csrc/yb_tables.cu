// Host-side meta pass: YASTN's meta tuples (flattened to int64 tables) -> the record tables of the copy / GEMM-scatter /
// elementwise plans.  Pure host code (no device calls) so that it is unit-tested on CPU; it lives in the library because a
// launch-bound run (DMRG at D=64: ~20 new block structures per bond) spends its time building plans, and the same passes in
// numpy cost 0.1-0.4 ms each (profiles/host_profile_r02.txt) against ~5 us here.
//
// Reference loops whose index arithmetic these tables encode:
//   yb_tables_merge    backend.transpose_and_merge   yastn/backend/_backend_torch_backwards.py:340-364, metas of
//                                                     yastn/tensor/_merging.py:137-187
//   yb_tables_scatter  backend.unmerge               yastn/backend/_backend_torch_backwards.py:397-408, metas of
//                                                     yastn/tensor/_merging.py:528-549 (applied inside the GEMM epilogue)
//   yb_tables_add      backend.add / sub             yastn/backend/backend_torch.py:518-534
#include <algorithm>
#include <numeric>

#include "yb_common.h"

namespace yb {
namespace {

std::vector<int64_t>& result_slot() {
    static thread_local std::vector<int64_t> buf;
    return buf;
}

constexpr int kMaxRank = 16;
constexpr int64_t kZeroRowsMax = 1 << 16;   // above this many zero rows a memset of the whole destination is cheaper

inline void cstrides(const int64_t* shape, int n, int64_t* st) {
    int64_t s = 1;
    for (int k = n - 1; k >= 0; --k) {
        st[k] = s;
        s *= shape[k];
    }
}

}  // namespace
}  // namespace yb

using namespace yb;

extern "C" int64_t yb_tables_result_size(void) { return (int64_t)result_slot().size(); }

extern "C" int yb_tables_result_fetch(int64_t* out, int64_t n) {
    auto& buf = result_slot();
    if (n != (int64_t)buf.size() || (n > 0 && !out)) return fail(kErrArg, "yb_tables_result_fetch: size mismatch");
    if (n) memcpy(out, buf.data(), (size_t)n * sizeof(int64_t));
    return kOk;
}

// mrg: n x wm rows [tn(T), slo(2), Do(r), Dslc(2g), Drsh(g)], neu: nnew x wn rows [tn(T), Dn(g), sln(2)], grp_in (optional):
// the row of `neu` every record of `mrg` merges into (otherwise derived from the order of the charges).
// Result: [status, rank, covered, nrec, recs(nrec x (2 + 3 rank))...]; status 0 ok, 1 = the records are not grouped in the
// order of `neu` (call again with grp_in), 2 = zero-fill records were not built (the caller clears the destination).
extern "C" int yb_tables_merge(const int64_t* mrg, int64_t n, int64_t wm, const int64_t* neu, int64_t nnew, int64_t wn,
                               const int64_t* order, int r, int g, int T, const int64_t* grp_in, int zero_records) {
    auto& out = result_slot();
    out.clear();
    if (n <= 0 || !mrg || !neu) return fail(kErrArg, "yb_tables_merge: empty input");
    if (r > kMaxRank || g > kMaxRank || r < 0 || g < 0) return fail(kErrArg, "yb_tables_merge: rank above %d", kMaxRank);
    if (wm != T + 2 + r + 3 * g || wn != T + g + 2) return fail(kErrArg, "transpose_and_merge: unexpected meta layout");
    std::vector<int64_t> grp((size_t)n, 0);
    if (grp_in) {
        for (int64_t i = 0; i < n; ++i) {
            if (grp_in[i] < 0 || grp_in[i] >= nnew) return fail(kErrArg, "yb_tables_merge: group index out of range");
            grp[(size_t)i] = grp_in[i];
        }
    } else {
        for (int64_t i = 1; i < n; ++i) {
            bool diff = false;
            for (int t = 0; t < T; ++t) diff |= mrg[i * wm + t] != mrg[(i - 1) * wm + t];
            grp[(size_t)i] = grp[(size_t)i - 1] + (diff ? 1 : 0);
        }
        bool ok = grp[(size_t)n - 1] < nnew;
        for (int64_t i = 0; ok && i < n; ++i)
            for (int t = 0; t < T; ++t) ok &= neu[grp[(size_t)i] * wn + t] == mrg[i * wm + t];
        if (!ok) {
            out.assign({1, 0, 0, 0});
            return kOk;
        }
    }
    const int rank = std::max(r, 1);
    const int w = 2 + 3 * rank;
    out.assign(4, 0);
    out[1] = rank;
    if (r == 0) {   // rank-0 tensor: one element per record
        out.resize(4 + (size_t)n * w);
        for (int64_t i = 0; i < n; ++i) {
            int64_t* rec = &out[4 + (size_t)i * w];
            rec[0] = mrg[i * wm + T];
            rec[1] = neu[grp[(size_t)i] * wn + T + g];
            rec[2] = rec[3] = rec[4] = 1;
        }
        out[2] = n;
        out[3] = n;
        return kOk;
    }
    out.resize(4 + (size_t)n * w);
    int64_t covered = 0;
    std::vector<int64_t> vol((size_t)nnew, 0);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t* m = mrg + i * wm;
        const int64_t* nw = neu + grp[(size_t)i] * wn;
        const int64_t* Do = m + T + 2;
        const int64_t* Dslc = Do + r;
        const int64_t* Drsh = Dslc + 2 * g;
        const int64_t* Dn = nw + T;
        int64_t cs[kMaxRank], P[kMaxRank], pstr[kMaxRank], rstr[kMaxRank], nstr[kMaxRank];
        cstrides(Do, r, cs);
        for (int k = 0; k < r; ++k) {
            if (order[k] < 0 || order[k] >= r) return fail(kErrArg, "transpose_and_merge: bad axis order");
            P[k] = Do[order[k]];
        }
        cstrides(P, r, pstr);
        cstrides(Drsh, g, rstr);
        cstrides(Dn, g, nstr);
        int64_t* rec = &out[4 + (size_t)i * w];
        int64_t dst = nw[T + g], elems = 1, box = 1;
        for (int q = 0; q < g; ++q) {
            dst += Dslc[2 * q] * nstr[q];
            box *= Dslc[2 * q + 1] - Dslc[2 * q];
        }
        rec[0] = m[T];
        rec[1] = dst;
        for (int k = 0; k < r; ++k) {
            // every permuted dim lies inside exactly one reshape group: rstr[q] <= pstr[k] and the dim fits in the group
            int q = -1;
            for (int c = 0; c < g; ++c)
                if (rstr[c] <= pstr[k] && pstr[k] * P[k] <= rstr[c] * Drsh[c]) {
                    q = c;
                    break;
                }
            if (P[k] > 1 && q < 0) return fail(kErrArg, "transpose_and_merge: reshape groups do not align with permuted dims");
            rec[2 + k] = P[k];
            rec[2 + r + k] = cs[order[k]];
            rec[2 + 2 * r + k] = (P[k] > 1 && rstr[q] > 0) ? (pstr[k] / rstr[q]) * nstr[q] : 0;
            elems *= P[k];
        }
        covered += elems;
        vol[(size_t)grp[(size_t)i]] += box;
    }
    int64_t nrec = n;
    int status = 0;
    if (zero_records) {
        // cells of the merged blocks that no source block covers (absent charge sectors): the source rectangles of one merged
        // block lie on a grid (one segment list per fused leg); the holes are the grid cells no record occupies
        std::vector<int64_t> holes;
        int64_t expect = 0;
        for (int64_t t = 0; t < nnew; ++t) {
            int64_t full = 1;
            for (int q = 0; q < g; ++q) full *= neu[t * wn + T + q];
            if (vol[(size_t)t] < full) {
                holes.push_back(t);
                expect += full - vol[(size_t)t];
            }
        }
        if (!holes.empty()) {
            if (g > rank || g == 0) {
                status = 2;
            } else {
                std::vector<int64_t> idx_order((size_t)n);
                std::iota(idx_order.begin(), idx_order.end(), 0);
                std::stable_sort(idx_order.begin(), idx_order.end(), [&](int64_t a, int64_t b) { return grp[(size_t)a] < grp[(size_t)b]; });
                std::vector<int64_t> bounds((size_t)nnew + 1, 0);
                for (int64_t i = 0; i < n; ++i) bounds[(size_t)grp[(size_t)i] + 1]++;
                for (int64_t t = 0; t < nnew; ++t) bounds[(size_t)t + 1] += bounds[(size_t)t];
                std::vector<int64_t> zrecs;
                int64_t rows = 0, zeros = 0;
                std::vector<int64_t> cuts[kMaxRank];
                std::vector<char> occ;
                for (int64_t t : holes) {
                    const int64_t* Dn = neu + t * wn + T;
                    int64_t nstr[kMaxRank], gdim[kMaxRank], gstr[kMaxRank];
                    cstrides(Dn, g, nstr);
                    for (int d = 0; d < g; ++d) {
                        auto& c = cuts[d];
                        c.clear();
                        c.push_back(0);
                        c.push_back(Dn[d]);
                        for (int64_t j = bounds[(size_t)t]; j < bounds[(size_t)t + 1]; ++j) {
                            const int64_t* Dslc = mrg + idx_order[(size_t)j] * wm + T + 2 + r;
                            c.push_back(Dslc[2 * d]);
                            c.push_back(Dslc[2 * d + 1]);
                        }
                        std::sort(c.begin(), c.end());
                        c.erase(std::unique(c.begin(), c.end()), c.end());
                        gdim[d] = (int64_t)c.size() - 1;
                    }
                    cstrides(gdim, g, gstr);
                    const int64_t cells = g > 0 ? gstr[0] * gdim[0] : 1;
                    occ.assign((size_t)cells, 0);
                    for (int64_t j = bounds[(size_t)t]; j < bounds[(size_t)t + 1]; ++j) {
                        const int64_t* Dslc = mrg + idx_order[(size_t)j] * wm + T + 2 + r;
                        int64_t c0[kMaxRank], c1[kMaxRank], cur[kMaxRank];
                        bool empty = false;
                        for (int d = 0; d < g; ++d) {
                            c0[d] = std::lower_bound(cuts[d].begin(), cuts[d].end(), Dslc[2 * d]) - cuts[d].begin();
                            c1[d] = std::lower_bound(cuts[d].begin(), cuts[d].end(), Dslc[2 * d + 1]) - cuts[d].begin();
                            cur[d] = c0[d];
                            empty |= c1[d] <= c0[d];
                        }
                        if (empty) continue;
                        while (true) {     // odometer over the occupied grid cells
                            int64_t lin = 0;
                            for (int d = 0; d < g; ++d) lin += cur[d] * gstr[d];
                            occ[(size_t)lin] = 1;
                            int d = g - 1;
                            while (d >= 0 && ++cur[d] == c1[d]) {
                                cur[d] = c0[d];
                                --d;
                            }
                            if (d < 0) break;
                        }
                    }
                    const int64_t last = gdim[g - 1];
                    for (int64_t k = 0; k < cells;) {
                        if (occ[(size_t)k]) {
                            ++k;
                            continue;
                        }
                        // free cells that follow each other along the last dim (same leading indices) make one box
                        int64_t e = k + 1;
                        while (e < cells && e % last != 0 && !occ[(size_t)e]) ++e;
                        int64_t rec[2 + 3 * kMaxRank];
                        for (int j = 0; j < w; ++j) rec[j] = 0;
                        for (int j = 0; j < rank; ++j) rec[2 + j] = 1;
                        int64_t rem = k, base = neu[t * wn + T + g], rprod = 1, eprod = 1;
                        for (int d = 0; d < g; ++d) {
                            const int64_t c = rem / gstr[d];
                            rem %= gstr[d];
                            const int64_t lo = cuts[d][(size_t)c];
                            const int64_t hi = d == g - 1 ? cuts[d][(size_t)(c + (e - k))] : cuts[d][(size_t)c + 1];
                            rec[2 + d] = hi - lo;
                            rec[2 + 2 * rank + d] = nstr[d];
                            base += lo * nstr[d];
                            if (d < g - 1) rprod *= hi - lo;
                            eprod *= hi - lo;
                        }
                        rec[0] = YB_COPY_SRC_ZERO;
                        rec[1] = base;
                        zrecs.insert(zrecs.end(), rec, rec + w);
                        rows += g > 1 ? rprod : 1;
                        zeros += eprod;
                        k = e;
                    }
                }
                if (rows > kZeroRowsMax || zeros != expect) {
                    status = 2;   // too fragmented (or overlapping source boxes): memset instead
                } else {
                    out.insert(out.end(), zrecs.begin(), zrecs.end());
                    nrec += (int64_t)zrecs.size() / w;
                    covered += zeros;
                }
            }
        }
    }
    out[0] = status;
    out[2] = covered;
    out[3] = nrec;
    return kOk;
}

// um: n x (10 + gn) rows [sln(2), Dn(gn), slo(2), Do(2), r0, r1, c0, c1]; md: nprob x 12 rows of meta_dot
// [slc(2), Dc(2), sla(2), Da(2), slb(2), Db(2)]; shift (optional): one destination shift per record of um.
// Result: [nprob, ng, nrow_cuts, ncol_cuts, ndst, scat_index(nprob), row_ptr(ng+1), row_cuts, col_ptr(ng+1), col_cuts,
//          dst_ptr(ng+1), dst] — the tables of yb_gemm_plan_create_scatter.
extern "C" int yb_tables_scatter(const int64_t* um, int64_t n, int gn, const int64_t* md, int64_t nprob, const int64_t* shift) {
    auto& out = result_slot();
    out.clear();
    if (n <= 0 || !um || (nprob > 0 && !md)) return fail(kErrArg, "yb_tables_scatter: empty input");
    const int w = 10 + gn;
    auto U = [&](int64_t i, int c) { return um[i * w + c]; };
    const int cSlo = 2 + gn, cM = 4 + gn, cN = 5 + gn, cR0 = 6 + gn, cR1 = 7 + gn, cC0 = 8 + gn, cC1 = 9 + gn;
    std::vector<int64_t> src((size_t)n);
    for (int64_t i = 0; i < n; ++i) src[(size_t)i] = U(i, cSlo);
    std::sort(src.begin(), src.end());
    src.erase(std::unique(src.begin(), src.end()), src.end());
    const int64_t ng = (int64_t)src.size();
    std::vector<int64_t> grp((size_t)n);
    std::vector<std::vector<int64_t>> rcuts((size_t)ng), ccuts((size_t)ng);
    std::vector<int64_t> count((size_t)ng, 0), Mg((size_t)ng, 0), Ng((size_t)ng, 0);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t gi = std::lower_bound(src.begin(), src.end(), U(i, cSlo)) - src.begin();
        grp[(size_t)i] = gi;
        rcuts[(size_t)gi].push_back(U(i, cR0));
        ccuts[(size_t)gi].push_back(U(i, cC0));
        count[(size_t)gi]++;
        Mg[(size_t)gi] = U(i, cM);
        Ng[(size_t)gi] = U(i, cN);
    }
    std::vector<int64_t> row_ptr((size_t)ng + 1, 0), col_ptr((size_t)ng + 1, 0), dst_ptr((size_t)ng + 1, 0);
    for (int64_t gi = 0; gi < ng; ++gi) {
        auto& rc = rcuts[(size_t)gi];
        auto& cc = ccuts[(size_t)gi];
        std::sort(rc.begin(), rc.end());
        rc.erase(std::unique(rc.begin(), rc.end()), rc.end());
        std::sort(cc.begin(), cc.end());
        cc.erase(std::unique(cc.begin(), cc.end()), cc.end());
        if (count[(size_t)gi] != (int64_t)(rc.size() * cc.size())) return fail(kErrArg, "unmerge rectangles of a block do not form a grid");
        if (rc[0] != 0 || cc[0] != 0) return fail(kErrArg, "unmerge rectangles do not tile the merged block");
        row_ptr[(size_t)gi + 1] = row_ptr[(size_t)gi] + (int64_t)rc.size() + 1;
        col_ptr[(size_t)gi + 1] = col_ptr[(size_t)gi] + (int64_t)cc.size() + 1;
        dst_ptr[(size_t)gi + 1] = dst_ptr[(size_t)gi] + (int64_t)(rc.size() * cc.size());
        rc.push_back(Mg[(size_t)gi]);
        cc.push_back(Ng[(size_t)gi]);
    }
    const int64_t ndst = dst_ptr[(size_t)ng];
    std::vector<int64_t> dst((size_t)ndst, -1);
    std::vector<char> placed((size_t)ndst, 0);
    for (int64_t i = 0; i < n; ++i) {
        const size_t gi = (size_t)grp[(size_t)i];
        auto& rc = rcuts[gi];
        auto& cc = ccuts[gi];
        const int64_t ri = std::lower_bound(rc.begin(), rc.end() - 1, U(i, cR0)) - rc.begin();
        const int64_t ci = std::lower_bound(cc.begin(), cc.end() - 1, U(i, cC0)) - cc.begin();
        if (rc[(size_t)ri + 1] != U(i, cR1) || cc[(size_t)ci + 1] != U(i, cC1))
            return fail(kErrArg, "unmerge rectangles do not tile the merged block");
        const int64_t ncs = (int64_t)cc.size() - 1;
        const size_t slot = (size_t)(dst_ptr[gi] + ri * ncs + ci);
        placed[slot] = 1;
        dst[slot] = U(i, 0) + (shift ? shift[i] : 0);
    }
    for (char p : placed)
        if (!p) return fail(kErrArg, "unmerge rectangles do not tile the merged block");
    std::vector<int64_t> scat((size_t)nprob, -1);
    std::vector<char> used((size_t)ng, 0);
    for (int64_t p = 0; p < nprob; ++p) {
        const int64_t slc0 = md[p * 12 + 0], Mp = md[p * 12 + 2], Np = md[p * 12 + 3];
        const int64_t pos = std::lower_bound(src.begin(), src.end(), slc0) - src.begin();
        const bool hit = pos < ng && src[(size_t)pos] == slc0, empty = Mp * Np == 0;
        if (!hit && !empty) return fail(kErrArg, "unmerge meta does not cover every block produced by dot");
        if (hit && !empty) {
            if (Mg[(size_t)pos] != Mp || Ng[(size_t)pos] != Np) return fail(kErrArg, "unmerge source shape differs from the dot block shape");
            scat[(size_t)p] = pos;
            used[(size_t)pos] = 1;
        }
    }
    for (char u : used)
        if (!u) return fail(kErrArg, "unmerge meta references blocks that dot does not produce");
    const int64_t nrow = row_ptr[(size_t)ng], ncol = col_ptr[(size_t)ng];
    out.reserve((size_t)(5 + nprob + 3 * (ng + 1) + nrow + ncol + ndst));
    out.assign({nprob, ng, nrow, ncol, ndst});
    out.insert(out.end(), scat.begin(), scat.end());
    out.insert(out.end(), row_ptr.begin(), row_ptr.end());
    for (auto& rc : rcuts) out.insert(out.end(), rc.begin(), rc.end());
    out.insert(out.end(), col_ptr.begin(), col_ptr.end());
    for (auto& cc : ccuts) out.insert(out.end(), cc.begin(), cc.end());
    out.insert(out.end(), dst_ptr.begin(), dst_ptr.end());
    out.insert(out.end(), dst.begin(), dst.end());
    return kOk;
}

// ops: nrec x 4 rows [operand, c0, c1, a0] (new[c0:c1] (+/-)= data_operand[a0 : a0 + c1 - c0]); signs: one per operand.
// Result: [nrounds, then per round: nrec, nslots, slots[4] (operand index, -1 = the running sum), recs (nrec x 16)].
// One launch adds up to four operands; later rounds read the running sum as their first source.
extern "C" int yb_tables_add(const int64_t* ops, int64_t nrec, int64_t n_ops, const int64_t* signs) {
    auto& out = result_slot();
    out.clear();
    if (n_ops <= 0 || (nrec > 0 && !ops)) return fail(kErrArg, "yb_tables_add: empty input");
    constexpr int kSources = 4;
    std::vector<int64_t> cuts;
    cuts.reserve((size_t)nrec * 2);
    for (int64_t i = 0; i < nrec; ++i) {
        cuts.push_back(ops[i * 4 + 1]);
        cuts.push_back(ops[i * 4 + 2]);
    }
    std::sort(cuts.begin(), cuts.end());
    cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
    const int64_t nint = cuts.empty() ? 0 : (int64_t)cuts.size() - 1;
    // elementary output intervals and who writes them
    std::vector<std::vector<std::pair<int64_t, int64_t>>> writers((size_t)nint);
    for (int64_t i = 0; i < nrec; ++i) {
        const int64_t k = ops[i * 4], c0 = ops[i * 4 + 1], c1 = ops[i * 4 + 2], a0 = ops[i * 4 + 3];
        if (k < 0 || k >= n_ops) return fail(kErrArg, "yb_tables_add: operand index out of range");
        if (c1 <= c0) continue;
        int64_t j = std::lower_bound(cuts.begin(), cuts.end() - 1, c0) - cuts.begin();
        for (; j < nint && cuts[(size_t)j] < c1; ++j) writers[(size_t)j].push_back({k, a0 + cuts[(size_t)j] - c0});
    }
    out.push_back(0);
    int64_t next_op = 0, nrounds = 0;
    bool first = true;
    while (next_op < n_ops) {
        const int64_t ntake = std::min<int64_t>(n_ops - next_op, first ? kSources : kSources - 1);
        int64_t slots[kSources] = {0, 0, 0, 0};
        int nslots = 0;
        if (!first) slots[nslots++] = -1;
        for (int64_t k = next_op; k < next_op + ntake; ++k) slots[nslots++] = k;
        const size_t head = out.size();
        out.insert(out.end(), {0, (int64_t)nslots, slots[0], slots[1], slots[2], slots[3]});
        int64_t m = 0;
        for (int64_t i = 0; i < nint; ++i) {
            const int64_t lo = cuts[(size_t)i], hi = cuts[(size_t)i + 1];
            int64_t srcs[kSources] = {YB_EW_ABSENT, YB_EW_ABSENT, YB_EW_ABSENT, YB_EW_ABSENT};
            int64_t neg = 0;
            bool hit = false;
            for (auto& wr : writers[(size_t)i]) {
                if (wr.first < next_op || wr.first >= next_op + ntake) continue;
                const int j = (int)(wr.first - next_op) + (first ? 0 : 1);
                if (srcs[j] != YB_EW_ABSENT) return fail(kErrArg, "add: an operand writes an output element twice");
                srcs[j] = wr.second;
                hit = true;
                if (signs && signs[wr.first] < 0) neg |= (int64_t)1 << j;
            }
            if (!first) {
                if (!hit) continue;   // nothing new for this interval: the running sum stays
                srcs[0] = lo;
            } else if (writers[(size_t)i].empty()) {
                continue;             // gap between blocks: belongs to no output block
            }
            const int64_t rec[16] = {YB_EW_LINCOMB, lo, hi - lo, srcs[0], srcs[1], srcs[2], srcs[3], neg, 0, 1, 1, 0, 0, 0, 0, 0};
            out.insert(out.end(), rec, rec + 16);
            ++m;
        }
        out[head] = m;
        next_op += ntake;
        first = false;
        ++nrounds;
    }
    out[0] = nrounds;
    return kOk;
}
