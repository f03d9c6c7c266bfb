#include "yb_common.h"

#include <algorithm>
#include <map>
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
    char* staging = nullptr;          // pinned host buffer of the batched table uploads
    size_t staging_bytes = 0;
    char* slab = nullptr;             // current slab small blocks are carved from
    size_t slab_used = 0;
    size_t pending_bytes = 0;
};
constexpr size_t kStagingMax = 64u << 20;
constexpr size_t kSlabBytes = 64u << 20;    // blocks up to 8 MiB are carved from slabs
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

namespace {
// Blocks released by destroyed plans may still be read by queued kernels: they wait in `pending` until a device-wide
// synchronisation has happened.  That synchronisation stalls the caller's pipeline, so it is only paid when a lot of memory is
// waiting (or the device is out of memory) — never just because one plan was evicted from a cache.
constexpr size_t kRecycleBytes = 512u << 20;

int recycle_pending(DevicePool& p) {     // p.mu held
    YB_CUDA(cudaDeviceSynchronize());
    for (int k = 0; k < 48; ++k) {
        p.free_blocks[k].insert(p.free_blocks[k].end(), p.pending[k].begin(), p.pending[k].end());
        p.pending[k].clear();
    }
    p.pending_bytes = 0;
    return kOk;
}

cudaError_t new_slab(DevicePool& p) {    // p.mu held; the rest of the previous slab is abandoned (less than one block)
    void* slab = nullptr;
    // stream-ordered allocation on the pool's private stream: a plain cudaMalloc waits for the kernels queued on every stream
    // (95-110 ms stalls in the middle of a D = 4096 sweep, profiles/plan_trace_hubbard_r02.jsonl)
    cudaError_t err = cudaErrorNotSupported;
    if (!p.copy_stream) {
        if (cudaStreamCreateWithFlags(&p.copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&p.copy_done, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            p.copy_stream = nullptr;
        }
    }
    if (p.copy_stream) {
        err = cudaMallocAsync(&slab, kSlabBytes, p.copy_stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(p.copy_stream);
        if (err != cudaSuccess) {
            cudaGetLastError();
            slab = nullptr;
        }
    }
    if (err != cudaSuccess) err = cudaMalloc(&slab, kSlabBytes);
    if (err != cudaSuccess) return err;
    p.slab = (char*)slab;
    p.slab_used = 0;
    return cudaSuccess;
}
}  // namespace

int pool_alloc(int device, size_t bytes, void** ptr, size_t* cap) {
    DevicePool& p = pool_of(device);
    const int c = size_class(bytes);
    *cap = (size_t)1 << c;
    std::lock_guard<std::mutex> lk(p.mu);
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (p.free_blocks[c].empty() && !p.pending[c].empty() && (p.pending_bytes >= kRecycleBytes || attempt == 1)) {
            int rc = recycle_pending(p);
            if (rc != kOk) return rc;
        }
        if (!p.free_blocks[c].empty()) {
            *ptr = p.free_blocks[c].back();
            p.free_blocks[c].pop_back();
            return kOk;
        }
        // No recycled block: carve one from the current slab.  A cudaMalloc per plan costs 100-250 us (and more for the MiB-sized
        // tables of a D = 4096 structure); a sweep creates thousands of plans that live as long as the plan cache
        // (profiles/host_profile_r02.txt).  A slab costs one cudaMalloc per 64 MiB.
        cudaError_t err = cudaSuccess;
        if (*cap <= kSlabBytes / 8) {
            // blocks are powers of two >= 512 B: aligning the bump pointer to the block size keeps every block naturally aligned
            size_t off = p.slab ? ((p.slab_used + *cap - 1) & ~(*cap - 1)) : kSlabBytes;
            if (off + *cap > kSlabBytes) {
                err = new_slab(p);
                off = 0;
            }
            if (err == cudaSuccess) {
                *ptr = p.slab + off;
                p.slab_used = off + *cap;
                return kOk;
            }
        } else {
            err = cudaMalloc(ptr, *cap);
            if (err == cudaSuccess) return kOk;
        }
        cudaGetLastError();
        if (attempt == 1 || p.pending_bytes == 0)
            return fail(kErrCuda, "yastn_b200: device allocation of %zu bytes for plan tables failed: %s", *cap, cudaGetErrorString(err));
        // out of memory with blocks waiting: synchronise, recycle everything and try once more
        int rc = recycle_pending(p);
        if (rc != kOk) return rc;
    }
    return fail(kErrCuda, "yastn_b200: device allocation of %zu bytes for plan tables failed", *cap);
}

void pool_free(int device, void* ptr, size_t cap) {
    DevicePool& p = pool_of(device);
    std::lock_guard<std::mutex> lk(p.mu);
    p.pending[size_class(cap)].push_back(ptr);
    p.pending_bytes += cap;
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

int TableBatch::commit() {
    size_t total = 0;
    for (auto& it : items) total += (it.bytes + 255) & ~(size_t)255;
    if (total == 0) return kOk;
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return fail(kErrCuda, "cudaGetDevice failed");
    if (total > kStagingMax) {     // large plans: one block and one (pageable) copy per table, as before
        for (auto& it : items) {
            int rc = it.table->upload(it.host, it.bytes);
            if (rc != kOk) return rc;
        }
        return kOk;
    }
    void* block = nullptr;
    size_t cap = 0;
    int rc = pool_alloc(device, total, &block, &cap);
    if (rc != kOk) return rc;
    DevicePool& p = pool_of(device);
    std::lock_guard<std::mutex> lk(p.mu);
    bool first = true;
    size_t off = 0;
    // hand the block out before anything can fail: the plan's destroy path then returns it to the pool
    for (auto& it : items) {
        DeviceTable& t = *it.table;
        t.bytes = it.bytes;
        t.device = device;
        if (it.bytes == 0) continue;
        t.ptr = (char*)block + off;
        t.owner = first;
        t.cap = first ? cap : 0;
        first = false;
        off += (it.bytes + 255) & ~(size_t)255;
    }
    if (!p.copy_stream) {
        YB_CUDA(cudaStreamCreateWithFlags(&p.copy_stream, cudaStreamNonBlocking));
        YB_CUDA(cudaEventCreateWithFlags(&p.copy_done, cudaEventDisableTiming));
    }
    if (p.staging_bytes < total) {
        if (p.staging) cudaFreeHost(p.staging);
        p.staging = nullptr;
        p.staging_bytes = 0;
        size_t want = 1u << 16;
        while (want < total) want <<= 1;
        YB_CUDA(cudaHostAlloc((void**)&p.staging, want, cudaHostAllocDefault));
        p.staging_bytes = want;
    }
    off = 0;
    for (auto& it : items) {
        if (it.bytes == 0) continue;
        memcpy(p.staging + off, it.host, it.bytes);
        off += (it.bytes + 255) & ~(size_t)255;
    }
    YB_CUDA(cudaMemcpyAsync(block, p.staging, off, cudaMemcpyHostToDevice, p.copy_stream));
    YB_CUDA(cudaEventRecord(p.copy_done, p.copy_stream));
    YB_CUDA(cudaEventSynchronize(p.copy_done));   // the staging buffer is free again; the host waited for this copy only
    return kOk;
}

namespace {
struct StreamWs {
    void* ws = nullptr;
    int* flags = nullptr;
    size_t ws_bytes = 0, nflags = 0;
};
constexpr size_t kWsDefaultBytes = 32u << 20;     // stream-K: 296 slots x 64 KB = 18.5 MiB; skinny partials: 0.5..1 KB per run
constexpr size_t kWsDefaultFlags = 1u << 16;
std::mutex ws_mu;
std::map<std::pair<int, cudaStream_t>, StreamWs> ws_table;
std::vector<StreamWs> ws_spare[64];

// cudaMalloc + zeroed flags, without touching any caller stream (the pool's copy stream, host-synchronised)
int ws_alloc(int device, size_t ws_bytes, size_t nflags, StreamWs* out) {
    StreamWs e;
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != device) YB_CUDA(cudaSetDevice(device));
    cudaError_t err = cudaMalloc(&e.ws, ws_bytes);
    if (err == cudaSuccess) err = cudaMalloc((void**)&e.flags, nflags * sizeof(int));
    if (err == cudaSuccess) {
        DevicePool& pool = pool_of(device);
        std::lock_guard<std::mutex> lk(pool.mu);
        if (!pool.copy_stream) {
            err = cudaStreamCreateWithFlags(&pool.copy_stream, cudaStreamNonBlocking);
            if (err == cudaSuccess) err = cudaEventCreateWithFlags(&pool.copy_done, cudaEventDisableTiming);
        }
        if (err == cudaSuccess) err = cudaMemsetAsync(e.flags, 0, nflags * sizeof(int), pool.copy_stream);
        if (err == cudaSuccess) err = cudaEventRecord(pool.copy_done, pool.copy_stream);
        if (err == cudaSuccess) err = cudaEventSynchronize(pool.copy_done);
    }
    if (prev != device) cudaSetDevice(prev);
    if (err != cudaSuccess) {
        cudaGetLastError();
        return fail(kErrCuda, "yastn_b200: allocating a %zu-byte reduction workspace failed: %s (a stream that is being captured into "
                    "a CUDA graph cannot allocate: run the call once on that stream before capturing)", ws_bytes, cudaGetErrorString(err));
    }
    e.ws_bytes = ws_bytes;
    e.nflags = nflags;
    *out = e;
    return kOk;
}
}  // namespace

int workspace_reserve(int device) {
    if (device < 0 || device >= 64) return fail(kErrArg, "yastn_b200: device index %d out of range", device);
    std::lock_guard<std::mutex> lk(ws_mu);
    if (!ws_spare[device].empty()) return kOk;
    StreamWs e;
    int rc = ws_alloc(device, kWsDefaultBytes, kWsDefaultFlags, &e);
    if (rc != kOk) return rc;
    ws_spare[device].push_back(e);
    return kOk;
}

int stream_workspace(int device, cudaStream_t st, size_t ws_bytes, size_t nflags, void** ws, int** flags) {
    if (device < 0 || device >= 64) return fail(kErrArg, "yastn_b200: device index %d out of range", device);
    std::lock_guard<std::mutex> lk(ws_mu);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) {
        cudaGetLastError();
        cap = cudaStreamCaptureStatusNone;
    }
    const bool capturing = cap != cudaStreamCaptureStatusNone;
    auto it = ws_table.find({device, st});
    if (it == ws_table.end()) {
        StreamWs e;
        if (!ws_spare[device].empty() && ws_spare[device].back().ws_bytes >= ws_bytes && ws_spare[device].back().nflags >= nflags) {
            e = ws_spare[device].back();     // a stream seen for the first time (e.g. torch's graph-capture stream): no allocation
            ws_spare[device].pop_back();
        } else {
            int rc = ws_alloc(device, std::max(ws_bytes, kWsDefaultBytes), std::max(nflags, kWsDefaultFlags), &e);
            if (rc != kOk) return rc;
        }
        it = ws_table.emplace(std::make_pair(device, st), e).first;
    }
    // keep one spare around for the next unknown stream; allocation is only legal outside stream capture
    if (!capturing && ws_spare[device].empty()) {
        StreamWs s;
        if (ws_alloc(device, kWsDefaultBytes, kWsDefaultFlags, &s) == kOk) ws_spare[device].push_back(s);
    }
    StreamWs& e = it->second;
    if (e.ws_bytes < ws_bytes || e.nflags < nflags) {
        // grow: the old buffers stay allocated (kernels already queued, or captured in a graph, still use them)
        StreamWs g;
        int rc = ws_alloc(device, std::max(ws_bytes, 2 * e.ws_bytes), std::max(nflags, 2 * e.nflags), &g);
        if (rc != kOk) return rc;
        e = g;
    }
    *ws = e.ws;
    *flags = e.flags;
    return kOk;
}

}  // namespace yb

extern "C" int yb_abi_version(void) { return YB_ABI_VERSION; }
extern "C" const char* yb_last_error(void) { return yb::last_error_slot().c_str(); }
