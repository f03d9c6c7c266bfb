// Microbenchmark: FP64 pipe throughput on sm_100a (DMMA.8x8x4 vs DFMA), used to pick the
// grouped-GEMM inner instruction and to record the FP64 roofline denominator.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dmma_loop(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    double c[ILP][2];
#pragma unroll
    for (int j = 0; j < ILP; j++) { c[j][0] = j; c[j][1] = -j; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dfma_loop(double* out, int iters) {
    double a = threadIdx.x * 1e-3 + 1.0, b = threadIdx.x * 2e-9;
    double c[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) c[j] = j;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) c[j] = fma(c[j], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) s += c[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, sms);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    int warps_list[] = {4, 8, 16, 32};
    for (int w : warps_list) {
        int threads = w * 32 > 1024 ? 1024 : w * 32;
        int blocks = sms * ((w * 32 + threads - 1) / threads);
        {
            float ms = time_it([&] { dmma_loop<8><<<blocks, threads>>>(out, iters); });
            double flops = 2.0 * 8 * 8 * 4 * 8.0 * iters * (double)blocks * (threads / 32);
            printf("{\"op\": \"DMMA.8x8x4\", \"ilp\": 8, \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", w, ms, flops / ms * 1e-9);
        }
        {
            float ms = time_it([&] { dmma_loop<2><<<blocks, threads>>>(out, iters); });
            double flops = 2.0 * 8 * 8 * 4 * 2.0 * iters * (double)blocks * (threads / 32);
            printf("{\"op\": \"DMMA.8x8x4\", \"ilp\": 2, \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", w, ms, flops / ms * 1e-9);
        }
        {
            float ms = time_it([&] { dfma_loop<8><<<blocks, threads>>>(out, iters); });
            double flops = 2.0 * 8.0 * iters * (double)blocks * threads;
            printf("{\"op\": \"DFMA\", \"ilp\": 8, \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", w, ms, flops / ms * 1e-9);
        }
    }
    cudaFree(out);
    return 0;
}
