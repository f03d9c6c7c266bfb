// Shared helpers of the yastn_b200 CUDA library (host side error handling, device table upload).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/yastn_b200.h"

namespace yb {

enum : int { kOk = 0, kErrArg = -1, kErrCuda = -2, kErrUnsupported = -3 };

std::string& last_error_slot();
int fail(int code, const char* fmt, ...);

#define YB_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t err__ = (expr);                                                           \
        if (err__ != cudaSuccess)                                                             \
            return yb::fail(yb::kErrCuda, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
    } while (0)

// Owns one device allocation holding a host-built table.
struct DeviceTable {
    void* ptr = nullptr;
    size_t bytes = 0;
    int upload(const void* host, size_t nbytes) {
        bytes = nbytes;
        if (nbytes == 0) return kOk;
        YB_CUDA(cudaMalloc(&ptr, nbytes));
        YB_CUDA(cudaMemcpy(ptr, host, nbytes, cudaMemcpyHostToDevice));
        return kOk;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
    }
};

// Division by a run-time invariant (valid for numerators < 2^31).
struct FastDiv {
    uint32_t div = 1, mul = 0, shr = 0;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.div = d;
    if (d <= 1) return f;
    uint32_t lg = 0;
    while ((1ull << lg) < d) ++lg;
    uint32_t p = 31 + lg;
    f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
    f.shr = p - 32;
    return f;
}

}  // namespace yb
