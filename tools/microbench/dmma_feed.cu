// Microbenchmark: what does feeding the FP64 tensor pipe cost?  A GEMM-like inner loop (MT x NT DMMA.8x8x4 tiles per
// warp per k-step) with the operand-feed components switched on one by one:
//   LDS  : fragments re-loaded from shared memory every k-step (LDS.64, or LDS.128 fetching two k-steps at once)
//   SYNC : __syncthreads every 4 k-steps (one pipeline stage)
//   CPA  : 12 x 16-byte cp.async per thread per stage from an L2-resident buffer (+ commit / wait_group)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_feed dmma_feed.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds128(unsigned a) { double2 v; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }

template <int MT, int NT, int LDS /*0 none, 1 LDS.64, 2 LDS.128*/, bool SYNC, bool CPA, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) feed(double* out, const double* src, int stages) {
    extern __shared__ __align__(128) char smem_raw[];
    const unsigned smem = (unsigned)__cvta_generic_to_shared(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    double* sm = reinterpret_cast<double*>(smem_raw);
    for (int i = tid; i < 3 * 24576 / 8; i += THREADS) sm[i] = 1e-3 * (i & 127);
    __syncthreads();
    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double af[MT], bf[NT];
#pragma unroll
    for (int i = 0; i < MT; ++i) af[i] = 1e-3 * (lane + i);
#pragma unroll
    for (int j = 0; j < NT; ++j) bf[j] = 1e-3 * (lane - j);
    // conflict-free fragment addresses: lane -> 8 rows x 4 k, rows 128 B apart with a chunk swizzle
    const int lx = lane >> 2, lk = lane & 3;
    unsigned aoff[MT], boff[NT];
#pragma unroll
    for (int i = 0; i < MT; ++i) aoff[i] = (i * 8 + lx) * 128 + ((((lk * 8) >> 4) ^ ((lx & 3) << 1)) << 4) + ((lk * 8) & 15);
#pragma unroll
    for (int j = 0; j < NT; ++j) boff[j] = 8192 + (j * 8 + lx) * 128 + ((((lk * 8) >> 4) ^ ((lx & 3) << 1)) << 4) + ((lk * 8) & 15);
    const char* g = reinterpret_cast<const char*>(src) + (size_t)(blockIdx.x % 64) * 24576 + tid * 16;
    for (int s = 0; s < stages; ++s) {
        const unsigned sa = smem + (s % 3) * 24576;
        if (CPA) {
            asm volatile("cp.async.wait_group 1;");
        }
        if (SYNC) __syncthreads();
        if (CPA) {
            const unsigned sd = smem + ((s + 2) % 3) * 24576 + tid * 16;
#pragma unroll
            for (int c = 0; c < 12 * 128 / THREADS; ++c)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sd + c * THREADS * 16), "l"(g + c * THREADS * 16));
            asm volatile("cp.async.commit_group;");
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            if (LDS == 1) {
#pragma unroll
                for (int i = 0; i < MT; ++i) af[i] = lds64(sa + (aoff[i] ^ (ks * 32)));
#pragma unroll
                for (int j = 0; j < NT; ++j) bf[j] = lds64(sa + (boff[j] ^ (ks * 32)));
            } else if (LDS == 2) {
                if ((ks & 1) == 0) {
                    // one LDS.128 per fragment per TWO k-steps (k permuted inside the stage)
                    double2 t;
#pragma unroll
                    for (int i = 0; i < MT; ++i) { t = lds128(sa + ((aoff[i] & ~15u) ^ (ks * 32))); af[i] = t.x + t.y; }
#pragma unroll
                    for (int j = 0; j < NT; ++j) { t = lds128(sa + ((boff[j] & ~15u) ^ (ks * 32))); bf[j] = t.x - t.y; }
                }
            }
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    double sum = 0;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) sum += acc[i][j][0] + acc[i][j][1];
    out[(size_t)blockIdx.x * THREADS + tid] = sum;
}

template <typename K>
void bench(const char* name, K kern, int threads, int per_sm, int mt, int nt, int sms, double* out, const double* src) {
    const int stages = 4000, smem = 3 * 24576;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
    const int blocks = sms * per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<blocks, threads, smem>>>(out, src, stages); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); kern<<<blocks, threads, smem>>>(out, src, stages); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double flops = 512.0 * mt * nt * 4 * stages * (double)blocks * (threads / 32);
    printf("{\"variant\": \"%s\", \"threads\": %d, \"ctas_per_sm\": %d, \"occ\": %d, \"ms\": %.3f, \"tflops\": %.2f, \"err\": \"%s\"}\n", name, threads, per_sm, occ,
           best, flops / best * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *out, *src;
    cudaMalloc(&out, sizeof(double) * sms * 4 * 1024);
    cudaMalloc(&src, 64 * 24576 + 65536);
    cudaMemset(src, 0, 64 * 24576 + 65536);
#define B(name, MT, NT, LDS, SYNC, CPA, TH, MINB) bench(name, feed<MT, NT, LDS, SYNC, CPA, TH, MINB>, TH, MINB, MT, NT, sms, out, src)
    B("4x8 regs only", 4, 8, 0, false, false, 128, 2);
    B("4x8 +LDS64", 4, 8, 1, false, false, 128, 2);
    B("4x8 +LDS128", 4, 8, 2, false, false, 128, 2);
    B("4x8 +LDS64 +sync", 4, 8, 1, true, false, 128, 2);
    B("4x8 +LDS64 +sync +cpasync", 4, 8, 1, true, true, 128, 2);
    B("4x8 +LDS128 +sync +cpasync", 4, 8, 2, true, true, 128, 2);
    B("4x8 regs +sync +cpasync", 4, 8, 0, true, true, 128, 2);
    B("8x4 +LDS64 +sync +cpasync", 8, 4, 1, true, true, 128, 2);
    B("4x4 +LDS64 +sync +cpasync 256thr", 4, 4, 1, true, true, 256, 2);
    B("4x8 +LDS64 +sync +cpasync 256thr x1", 4, 8, 1, true, true, 256, 1);
    B("4x4 +LDS64 +sync +cpasync 128thr x4", 4, 4, 1, true, true, 128, 3);
    return 0;
}
