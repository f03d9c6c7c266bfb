// Batched one-sided Jacobi SVD of the small charge sectors (up to 64 x 64): ONE launch factorises every small sector of a
// block matrix, one CTA per sector, the sector and the accumulated rotations resident in shared memory.
//
// Reference: yastn/backend/_backend_torch_backwards.py:26-39 (kernel_svd.forward) loops over the sectors and calls
// torch.linalg.svd(driver='gesvd') on each (backend/linalg/torch_svd_gesdd.py:17): for a D = 64 Heisenberg chain that is 8
// cuSOLVER calls of 0.5-1 ms plus a host synchronisation each per bond — half of a DMRG sweep (profiles/e2e_r02_c.jsonl), and
// 5.8 ms for one 64 x 64 complex128 sector (profiles/svd_probe_r02.jsonl).  Small matrices want a different algorithm:
// Hestenes' one-sided Jacobi orthogonalises the columns of G (= A, or A^H for a wide sector) by plane rotations,
//     G V = U S   =>   A = U S V^H,
// converges quadratically in ~6-10 sweeps, needs only dot products and column updates — all in shared memory — and
// computes even the tiny singular values to high RELATIVE accuracy (better than the bidiagonalisation route).  Column pairs of
// a sweep follow the round-robin tournament order, k/2 disjoint pairs per step, one warp per pair, so the result does not
// depend on scheduling: bit-reproducible.
#include <algorithm>

#include "yb_common.h"

namespace yb {

constexpr int kSvdThreads = 256;
constexpr int kSvdWarps = kSvdThreads / 32;
constexpr int kSvdMax = 64;

struct SvdRec {
    int64_t offA, offU, offS, offV;
    int32_t m, n;
};

__device__ __forceinline__ double sv_abs2(double v) { return v * v; }
__device__ __forceinline__ double sv_abs2(double2 v) { return v.x * v.x + v.y * v.y; }
__device__ __forceinline__ double sv_conj(double v) { return v; }
__device__ __forceinline__ double2 sv_conj(double2 v) { return make_double2(v.x, -v.y); }
// conj(a) * b
__device__ __forceinline__ double sv_cdot(double a, double b) { return a * b; }
__device__ __forceinline__ double2 sv_cdot(double2 a, double2 b) { return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double sv_add(double a, double b) { return a + b; }
__device__ __forceinline__ double2 sv_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double sv_zero(double) { return 0.0; }
__device__ __forceinline__ double2 sv_zero(double2) { return make_double2(0.0, 0.0); }
__device__ __forceinline__ double sv_scale(double v, double s) { return v * s; }
__device__ __forceinline__ double2 sv_scale(double2 v, double s) { return make_double2(v.x * s, v.y * s); }
__device__ __forceinline__ double sv_one(double) { return 1.0; }
__device__ __forceinline__ double2 sv_one(double2) { return make_double2(1.0, 0.0); }
__device__ __forceinline__ double sv_wsum(double v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ double2 sv_wsum(double2 v) { return make_double2(sv_wsum(v.x), sv_wsum(v.y)); }

// x' = c x - s e^{-i phi} y ;  y' = s x + c e^{-i phi} y      (ph = e^{-i phi}; real: ph = +-1)
__device__ __forceinline__ void sv_rot(double& x, double& y, double c, double s, double ph) {
    const double yy = ph * y, xo = x;
    x = c * xo - s * yy;
    y = s * xo + c * yy;
}
__device__ __forceinline__ void sv_rot(double2& x, double2& y, double c, double s, double2 ph) {
    const double2 yy = make_double2(ph.x * y.x - ph.y * y.y, ph.x * y.y + ph.y * y.x), xo = x;
    x = make_double2(c * xo.x - s * yy.x, c * xo.y - s * yy.y);
    y = make_double2(s * xo.x + c * yy.x, s * xo.y + c * yy.y);
}
__device__ __forceinline__ double sv_phase(double g, double /*absg*/) { return g < 0.0 ? -1.0 : 1.0; }
__device__ __forceinline__ double2 sv_phase(double2 g, double absg) { return make_double2(g.x / absg, -g.y / absg); }   // e^{-i phi}
__device__ __forceinline__ double sv_mag(double g) { return fabs(g); }
__device__ __forceinline__ double sv_mag(double2 g) { return hypot(g.x, g.y); }

template <typename T>
__global__ void __launch_bounds__(kSvdThreads) svd_jacobi_kernel(const SvdRec* __restrict__ recs, const T* __restrict__ A, T* __restrict__ U,
                                                                 double* __restrict__ S, T* __restrict__ Vh, int* __restrict__ status, int ldg,
                                                                 int ldv, int max_sweeps, int vectors) {
    extern __shared__ __align__(16) char svd_smem[];
    const SvdRec rec = recs[blockIdx.x];
    const int m = rec.m, n = rec.n;
    const bool tall = m >= n;
    const int L = tall ? m : n, k = tall ? n : m;            // G is L x k (columns are orthogonalised)
    T* G = reinterpret_cast<T*>(svd_smem);                    // G[j * ldg + r]  column j, row r
    T* V = G + (size_t)kSvdMax * ldg;                          // V[j * ldv + i]  column j of the accumulated rotations
    double* sig = reinterpret_cast<double*>(V + (size_t)kSvdMax * ldv);
    int* perm = reinterpret_cast<int*>(sig + kSvdMax);
    __shared__ int rotated;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const T* a = A + rec.offA;
    for (int e = tid; e < L * k; e += kSvdThreads) {
        const int j = e / L, r = e % L;
        G[j * ldg + r] = tall ? a[(int64_t)r * n + j] : sv_conj(a[(int64_t)j * n + r]);
    }
    for (int e = tid; e < k * k; e += kSvdThreads) {
        const int j = e / k, i = e % k;
        V[j * ldv + i] = i == j ? sv_one(T{}) : sv_zero(T{});
    }
    __syncthreads();
    const int kk = k + (k & 1);                               // players of the round-robin tournament (one dummy when k is odd)
    const double tol2 = 2.5e-31 * (double)L;                  // (sqrt(L) * eps / 2 ... )^2 : |g|^2 <= tol2 * a * b counts as orthogonal
    int sweeps = 0;
    bool converged = k < 2;
    while (!converged && sweeps < max_sweeps) {
        if (tid == 0) rotated = 0;
        __syncthreads();
        for (int step = 0; step < kk - 1; ++step) {
            for (int pi = warp; pi < kk / 2; pi += kSvdWarps) {
                int p, q;
                if (pi == 0) {
                    p = kk - 1;
                    q = step;
                } else {
                    p = (step + pi) % (kk - 1);
                    q = (step - pi + (kk - 1)) % (kk - 1);
                }
                if (p > q) {
                    const int t = p;
                    p = q;
                    q = t;
                }
                if (q >= k) continue;                         // the dummy player
                T* gp = G + p * ldg;
                T* gq = G + q * ldg;
                double al = 0.0, be = 0.0;
                T ga = sv_zero(T{});
                for (int r = lane; r < L; r += 32) {
                    const T x = gp[r], y = gq[r];
                    al += sv_abs2(x);
                    be += sv_abs2(y);
                    ga = sv_add(ga, sv_cdot(x, y));
                }
                al = sv_wsum(al);
                be = sv_wsum(be);
                ga = sv_wsum(ga);
                const double g2 = sv_abs2(ga);
                if (g2 > tol2 * al * be && g2 > 0.0) {
                    const double absg = sv_mag(ga);
                    const double zeta = (be - al) / (2.0 * absg);
                    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                    const auto ph = sv_phase(ga, absg);
                    for (int r = lane; r < L; r += 32) {
                        T x = gp[r], y = gq[r];
                        sv_rot(x, y, c, s, ph);
                        gp[r] = x;
                        gq[r] = y;
                    }
                    if (vectors) {
                        T* vp = V + p * ldv;
                        T* vq = V + q * ldv;
                        for (int r = lane; r < k; r += 32) {
                            T x = vp[r], y = vq[r];
                            sv_rot(x, y, c, s, ph);
                            vp[r] = x;
                            vq[r] = y;
                        }
                    }
                    if (lane == 0) rotated = 1;
                }
            }
            __syncthreads();
        }
        converged = rotated == 0;
        ++sweeps;
        __syncthreads();
    }
    // singular values = column norms; descending order
    for (int j = warp; j < k; j += kSvdWarps) {
        double al = 0.0;
        for (int r = lane; r < L; r += 32) al += sv_abs2(G[j * ldg + r]);
        al = sv_wsum(al);
        if (lane == 0) sig[j] = sqrt(al);
    }
    __syncthreads();
    for (int j = tid; j < k; j += kSvdThreads) {
        int rank = 0;
        const double sj = sig[j];
        for (int i = 0; i < k; ++i) rank += (sig[i] > sj || (sig[i] == sj && i < j)) ? 1 : 0;
        perm[rank] = j;
    }
    __syncthreads();
    int bad = converged ? 0 : 1;
    double smax = k > 0 ? sig[perm[0]] : 0.0;
    for (int jj = tid; jj < k; jj += kSvdThreads) S[rec.offS + jj] = sig[perm[jj]];
    if (vectors) {
        if (k > 0 && !(sig[perm[k - 1]] > 0.0) ) bad = 2;     // a zero (or NaN) singular value: its left vector is undefined here
        T* u = U + rec.offU;
        T* vh = Vh + rec.offV;
        if (tall) {          // U (m x k) = normalised columns of G ; Vh (k x n) = V^H
            for (int e = tid; e < m * k; e += kSvdThreads) {
                const int r = e / k, jj = e % k, j = perm[jj];
                u[e] = sv_scale(G[j * ldg + r], 1.0 / sig[j]);
            }
            for (int e = tid; e < k * n; e += kSvdThreads) {
                const int jj = e / n, c = e % n;
                vh[e] = sv_conj(V[perm[jj] * ldv + c]);
            }
        } else {             // G = A^H:  U (m x k) = accumulated rotations ; Vh (k x n) = (normalised columns of G)^H
            for (int e = tid; e < m * k; e += kSvdThreads) {
                const int i = e / k, jj = e % k;
                u[e] = V[perm[jj] * ldv + i];
            }
            for (int e = tid; e < k * n; e += kSvdThreads) {
                const int jj = e / n, c = e % n, j = perm[jj];
                vh[e] = sv_conj(sv_scale(G[j * ldg + c], 1.0 / sig[j]));
            }
        }
    }
    (void)smax;
    if (tid == 0) status[blockIdx.x] = bad;
}

}  // namespace yb

using namespace yb;

struct yb_svd_plan {
    int itemsize = 0, device = 0, nrec = 0;
    int ldg = 0, ldv = 0;
    size_t smem = 0;
    DeviceTable recs;
};

// recs: nrec x 6 int64 rows [offA, m, n, offU, offS, offV] (element offsets; A row-major m x n, U row-major m x k, S k reals,
// Vh row-major k x n, k = min(m, n)); every sector must satisfy max(m, n) <= 64.
extern "C" int yb_svd_plan_create(const int64_t* recs, int64_t nrec, int itemsize, int device, yb_svd_plan** out) {
    if (!out) return fail(kErrArg, "yb_svd_plan_create: out is null");
    *out = nullptr;
    if (nrec < 0 || (nrec > 0 && !recs)) return fail(kErrArg, "yb_svd_plan_create: bad table");
    if (itemsize != 8 && itemsize != 16) return fail(kErrUnsupported, "yb_svd_plan_create: itemsize %d (8 or 16)", itemsize);
    std::vector<SvdRec> h((size_t)nrec);
    int Lmax = 1, kmax = 1;
    for (int64_t i = 0; i < nrec; ++i) {
        const int64_t* q = recs + i * 6;
        if (q[1] < 1 || q[2] < 1 || q[1] > kSvdMax || q[2] > kSvdMax)
            return fail(kErrUnsupported, "yb_svd_plan_create: sector %lld is %lld x %lld (1..%d supported)", (long long)i, (long long)q[1], (long long)q[2], kSvdMax);
        h[(size_t)i] = {q[0], q[3], q[4], q[5], (int32_t)q[1], (int32_t)q[2]};
        Lmax = std::max<int>(Lmax, (int)std::max(q[1], q[2]));
        kmax = std::max<int>(kmax, (int)std::min(q[1], q[2]));
    }
    yb_svd_plan* plan = new yb_svd_plan();
    plan->itemsize = itemsize;
    plan->device = device;
    plan->nrec = (int)nrec;
    plan->ldg = Lmax + 1;      // odd leading dimensions: conflict-free column walks
    plan->ldv = kmax + 1;
    plan->smem = ((size_t)kSvdMax * plan->ldg + (size_t)kSvdMax * plan->ldv) * itemsize + kSvdMax * (sizeof(double) + sizeof(int));
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = kOk;
    if (cudaSetDevice(device) != cudaSuccess) rc = fail(kErrCuda, "yb_svd_plan_create: cudaSetDevice(%d) failed", device);
    if (rc == kOk) rc = plan->recs.upload(h.data(), h.size() * sizeof(SvdRec));
    if (rc == kOk) {
        cudaError_t e = itemsize == 8 ? cudaFuncSetAttribute(svd_jacobi_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
                                      : cudaFuncSetAttribute(svd_jacobi_kernel<double2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) rc = fail(kErrCuda, "yb_svd_plan_create: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    }
    cudaSetDevice(prev);
    if (rc != kOk) {
        plan->recs.release();
        delete plan;
        return rc;
    }
    *out = plan;
    return kOk;
}

// status: nrec int32 on the device — 0 ok, 1 not converged within max_sweeps, 2 a singular value is zero (U / Vh of that sector are
// not orthonormal: refactorise it with another routine).  vectors = 0 computes singular values only.
extern "C" int yb_svd_run(const yb_svd_plan* plan, const void* A, void* U, void* S, void* Vh, void* status, int max_sweeps, int vectors,
                          void* stream) {
    if (!plan) return fail(kErrArg, "yb_svd_run: plan is null");
    if (plan->nrec == 0) return kOk;
    if (!A || !S || !status || (vectors && (!U || !Vh))) return fail(kErrArg, "yb_svd_run: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->itemsize == 8)
        svd_jacobi_kernel<double><<<plan->nrec, kSvdThreads, plan->smem, st>>>((const SvdRec*)plan->recs.ptr, (const double*)A, (double*)U, (double*)S,
                                                                               (double*)Vh, (int*)status, plan->ldg, plan->ldv, max_sweeps, vectors);
    else
        svd_jacobi_kernel<double2><<<plan->nrec, kSvdThreads, plan->smem, st>>>((const SvdRec*)plan->recs.ptr, (const double2*)A, (double2*)U, (double*)S,
                                                                                (double2*)Vh, (int*)status, plan->ldg, plan->ldv, max_sweeps, vectors);
    YB_CUDA(cudaGetLastError());
    return kOk;
}

extern "C" void yb_svd_plan_destroy(yb_svd_plan* plan) {
    if (!plan) return;
    plan->recs.release();
    delete plan;
}
