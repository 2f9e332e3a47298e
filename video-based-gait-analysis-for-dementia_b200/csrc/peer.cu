// Peer-memory plumbing for the final gather of a sequence-sharded run: the root rank owns one device
// buffer, every other rank (one process per GPU) maps it through CUDA IPC and writes its block of meshes / joints straight
// into it over NVLink - with the skinning kernel's own stores (the `verts` argument of gait_smpl_lbs_tc_ex is then a peer
// address) or with asynchronous peer copies on the copy engines.  Replaces the per-process device-to-host boundary of the
// reference (batch_generation.py:316-323, demo.py:183-188) when one consumer wants all sequences.
#include <string.h>

#include "common.cuh"

using namespace gait;

extern "C" {

int gait_peer_alloc(void** ptr, size_t bytes) {
    GAIT_REQUIRE(ptr != nullptr && bytes > 0, "peer_alloc: null pointer or zero size");
    *ptr = nullptr;
    GAIT_CUDA(cudaMalloc(ptr, bytes));          // a whole cudaMalloc allocation: its IPC handle maps exactly this buffer
    return GAIT_OK;
}

int gait_peer_free(void* ptr) {
    if (ptr) GAIT_CUDA(cudaFree(ptr));
    return GAIT_OK;
}

int gait_peer_export(const void* ptr, unsigned char handle[GAIT_PEER_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == GAIT_PEER_HANDLE_BYTES, "IPC handle size");
    GAIT_REQUIRE(ptr != nullptr && handle != nullptr, "peer_export: null pointer");
    cudaIpcMemHandle_t h;
    GAIT_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
    memcpy(handle, &h, sizeof(h));
    return GAIT_OK;
}

int gait_peer_open(const unsigned char handle[GAIT_PEER_HANDLE_BYTES], void** ptr) {
    GAIT_REQUIRE(ptr != nullptr && handle != nullptr, "peer_open: null pointer");
    *ptr = nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    GAIT_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return GAIT_OK;
}

int gait_peer_close(void* ptr) {
    if (ptr) GAIT_CUDA(cudaIpcCloseMemHandle(ptr));
    return GAIT_OK;
}

int gait_peer_copy(void* dst, const void* src, size_t bytes, gait_stream_t stream) {
    GAIT_REQUIRE(bytes == 0 || (dst && src), "peer_copy: null pointer");
    if (bytes == 0) return GAIT_OK;
    GAIT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, as_stream(stream)));   // UVA: local or peer-mapped
    return GAIT_OK;
}

}  // extern "C"
