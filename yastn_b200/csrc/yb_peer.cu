// Peer arenas: device buffers of one rank (process) mapped into the address space of the other ranks of the same box.
//
// The multi-GPU path shards a contraction by charge sector and exchanges whole blocks only between contractions
// (SURVEY.md 8e).  With every rank's operand buffer living in a peer arena, that exchange is not a send/recv pair per
// block but ONE launch of the block-copy kernel (yb_copy.cu) whose records carry the *peer's* mapped address as their
// destination base: the SMs store the blocks straight into the remote HBM over NVLink / NVSwitch, all peers at once, and
// the fused unmerge epilogue of the grouped GEMM (yb_gemm.cu) can do the same through its destination-offset table —
// the result blocks of a contraction land on the rank that multiplies them next while the other tiles are still in the
// tensor pipe.  A plain stream-ordered collective (one tiny all-reduce) then publishes the data.
//
// The handles are CUDA IPC handles (cudaIpcGetMemHandle): 64 opaque bytes that the host side ships through
// torch.distributed (all_gather_object).  Opening a handle of a buffer on another device enables peer access lazily.
#include "yb_common.h"

using namespace yb;

extern "C" int yb_peer_alloc(int64_t bytes, int device, void** ptr, unsigned char handle[64]) {
    if (!ptr || !handle || bytes <= 0) return fail(kErrArg, "yb_peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    int prev = 0;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return fail(kErrCuda, "yb_peer_alloc: cudaSetDevice(%d) failed", device);
    int rc = kOk;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) rc = fail(kErrCuda, "yb_peer_alloc: cudaMalloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
    if (rc == kOk) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, p);
        if (e != cudaSuccess) {
            rc = fail(kErrCuda, "yb_peer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
            cudaFree(p);
        } else {
            memcpy(handle, &h, 64);
            *ptr = p;
        }
    }
    cudaSetDevice(prev);
    return rc;
}

extern "C" int yb_peer_open(const unsigned char handle[64], int device, void** ptr) {
    if (!ptr || !handle) return fail(kErrArg, "yb_peer_open: bad argument");
    int prev = 0;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return fail(kErrCuda, "yb_peer_open: cudaSetDevice(%d) failed", device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(kErrCuda, "yb_peer_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    *ptr = p;
    return kOk;
}

extern "C" int yb_peer_close(void* ptr) {
    if (!ptr) return kOk;
    YB_CUDA(cudaIpcCloseMemHandle(ptr));
    return kOk;
}

extern "C" int yb_peer_free(void* ptr) {
    if (!ptr) return kOk;
    YB_CUDA(cudaFree(ptr));
    return kOk;
}
