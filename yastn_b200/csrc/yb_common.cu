#include "yb_common.h"

#include <mutex>

namespace yb {

std::string& last_error_slot() {
    static thread_local std::string msg;
    return msg;
}

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_slot() = buf;
    return code;
}

namespace {
struct DevicePool {
    std::mutex mu;
    std::vector<void*> free_blocks[48];
    std::vector<void*> pending[48];   // released by a destroyed plan; kernels reading them may still be queued
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_done = nullptr;
};
DevicePool& pool_of(int device) {
    static DevicePool pools[64];
    return pools[(device >= 0 && device < 64) ? device : 0];
}
int size_class(size_t bytes) {
    int c = 9;   // 512 B minimum
    while (((size_t)1 << c) < bytes) ++c;
    return c;
}
}  // namespace

int pool_alloc(int device, size_t bytes, void** ptr, size_t* cap) {
    DevicePool& p = pool_of(device);
    const int c = size_class(bytes);
    *cap = (size_t)1 << c;
    {
        std::lock_guard<std::mutex> lk(p.mu);
        if (p.free_blocks[c].empty() && !p.pending[c].empty()) {
            // recycle: one device-wide sync makes every block released so far safe to overwrite
            YB_CUDA(cudaDeviceSynchronize());
            for (int k = 0; k < 48; ++k) {
                p.free_blocks[k].insert(p.free_blocks[k].end(), p.pending[k].begin(), p.pending[k].end());
                p.pending[k].clear();
            }
        }
        if (!p.free_blocks[c].empty()) {
            *ptr = p.free_blocks[c].back();
            p.free_blocks[c].pop_back();
            return kOk;
        }
    }
    YB_CUDA(cudaMalloc(ptr, *cap));
    return kOk;
}

void pool_free(int device, void* ptr, size_t cap) {
    DevicePool& p = pool_of(device);
    std::lock_guard<std::mutex> lk(p.mu);
    p.pending[size_class(cap)].push_back(ptr);
}

int pool_upload(int device, void* dst, const void* host, size_t bytes) {
    DevicePool& p = pool_of(device);
    std::lock_guard<std::mutex> lk(p.mu);
    if (!p.copy_stream) {
        YB_CUDA(cudaStreamCreateWithFlags(&p.copy_stream, cudaStreamNonBlocking));
        YB_CUDA(cudaEventCreateWithFlags(&p.copy_done, cudaEventDisableTiming));
    }
    YB_CUDA(cudaMemcpyAsync(dst, host, bytes, cudaMemcpyHostToDevice, p.copy_stream));
    YB_CUDA(cudaEventRecord(p.copy_done, p.copy_stream));
    YB_CUDA(cudaEventSynchronize(p.copy_done));   // the host waits for this copy only, not for the compute streams
    return kOk;
}

}  // namespace yb

extern "C" int yb_abi_version(void) { return YB_ABI_VERSION; }
extern "C" const char* yb_last_error(void) { return yb::last_error_slot().c_str(); }
