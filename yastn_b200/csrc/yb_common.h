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

// Device memory for plan tables comes from a per-device pool (power-of-two size classes, blocks are recycled when plans
// are destroyed) and is filled through a dedicated non-blocking copy stream: building a plan neither calls cudaMalloc in
// steady state nor synchronises with the kernels already queued on the caller's stream (a plain cudaMemcpy would: it
// runs on the legacy default stream).
int pool_alloc(int device, size_t bytes, void** ptr, size_t* cap);
void pool_free(int device, void* ptr, size_t cap);
int pool_upload(int device, void* dst, const void* host, size_t bytes);

// Scratch of the kernels that reduce across CTAs (stream-K partial tiles, skinny partial sums): one buffer per
// (device, stream), looked up at launch time and grown on demand, so launches on different streams never share slots.
// `flags` are zero when handed out for the first time and every kernel leaves them at zero.  A buffer that has been
// handed out is never freed (queued kernels and captured CUDA graphs keep using it); growth at least doubles it.
int stream_workspace(int device, cudaStream_t st, size_t ws_bytes, size_t nflags, void** ws, int** flags);
// Called when a plan that needs a workspace is created: keeps one spare default-sized workspace per device, so that the first
// launch on a stream never seen before — torch's CUDA-graph capture stream, where cudaMalloc is not allowed — finds one.
int workspace_reserve(int device);

// Owns one device allocation holding a host-built table (the current device must be the plan's device).
struct DeviceTable {
    void* ptr = nullptr;
    size_t bytes = 0, cap = 0;
    int device = 0;
    bool owner = true;      // false: a sub-range of the block another table of the same plan owns (TableBatch)
    int upload(const void* host, size_t nbytes) {
        bytes = nbytes;
        if (nbytes == 0) return kOk;
        if (cudaGetDevice(&device) != cudaSuccess) return fail(kErrCuda, "cudaGetDevice failed");
        int rc = pool_alloc(device, nbytes, &ptr, &cap);
        if (rc != kOk) return rc;
        return pool_upload(device, ptr, host, nbytes);
    }
    void release() {
        if (ptr && owner) pool_free(device, ptr, cap);
        ptr = nullptr;
    }
};

// All tables of one plan in ONE pool block filled by ONE host->device copy (through a pinned staging buffer): creating a plan
// costs one copy + one event wait instead of one per table (plan creation is what a D=64 DMRG sweep spends its time on:
// ~20 new structures per bond, profiles/host_profile_r02.txt).  Sub-ranges are 256-byte aligned.
struct TableBatch {
    struct Item {
        DeviceTable* table;
        const void* host;
        size_t bytes;
    };
    std::vector<Item> items;
    void add(DeviceTable& t, const void* host, size_t bytes) { items.push_back({&t, host, bytes}); }
    int commit();
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
