// Shared helpers for the gaitb200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "gaitb200.h"

namespace gait {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline cudaStream_t as_stream(gait_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Check the launch that was just enqueued; records the error text for gait_last_error().
inline int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("%s: %s", what, cudaGetErrorString(e));
        return GAIT_ERR_CUDA;
    }
    count_launch();
    return GAIT_OK;
}

#define GAIT_REQUIRE(cond, ...)                  \
    do {                                         \
        if (!(cond)) {                           \
            ::gait::set_error(__VA_ARGS__);      \
            return GAIT_ERR_INVALID;             \
        }                                        \
    } while (0)

#define GAIT_CUDA(call)                                                              \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            ::gait::set_error("%s: %s", #call, cudaGetErrorString(e__));             \
            return GAIT_ERR_CUDA;                                                    \
        }                                                                            \
    } while (0)

#define GAIT_TRY(call)                 \
    do {                               \
        int rc__ = (call);             \
        if (rc__ != GAIT_OK) return rc__; \
    } while (0)

// Per-device caches (a process may drive several GPUs, e.g. module.to('cuda:1') or one thread per device): SM count, and a
// once-per-device guard for cudaFuncSetAttribute (the attribute applies to the current device only).
int device_sm_count();                                   // api.cu
struct PerDeviceOnce {
    std::atomic<unsigned long long> done{0};
    // true exactly until mark() was called for the current device
    bool needed(int* dev) const {
        cudaGetDevice(dev);
        return *dev >= 64 || !((done.load(std::memory_order_acquire) >> *dev) & 1ull);
    }
    void mark(int dev) { if (dev < 64) done.fetch_or(1ull << dev, std::memory_order_release); }
};

// Programmatic dependent launch.  A kernel launched through launch_pdl() may begin while its predecessor in the stream is
// still draining: its CTAs become resident as SMs free up and run their prologue (barrier init, tensor-memory allocation,
// descriptor prefetch), then block in pdl_wait() until the predecessor has completed and its writes are visible.  Rules kept
// by every kernel launched this way: (1) ALL threads execute pdl_wait() before their first global-memory access (reads of
// constants included: a constant may have been written by the kernel just before), so "this grid completed" still implies
// "everything before it completed"; (2) pdl_trigger() right after it lets the successor be scheduled as soon as every CTA
// of this grid is resident.  Both are no-ops for a normal launch.
// GAITB200_PDL is a mask of the kernel kinds that are launched this way: 1 = tensor-core GEMM, 2 = skinning, 4 = the small
// kernels (chain, joint assembly, split-K reductions).  Measured in CUDA-graph replays of the C2 step (scripts/
// pdl_graph_check.py, profiles/r02za_pdl.md).  With every thread in griddepcontrol.wait: mask 0 729.6 us, 1 720.0, 2 743.5,
// 7 735.7 - CTAs that wait next to a running kernel slow it down.  With ONE waiting thread per CTA (pdl_wait_cta below; the
// others sit at the CTA barrier) the skinning kernel gains too: mask 1 700.7 us, 3 696.3, 5 704.2, 7 708.0 (after the other
// late changes).  The small kernels still lose, so the default is 3.
int pdl_mask();                                          // api.cu: bit 0 GEMM, bit 1 skinning, bit 2 small kernels
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// variant for kernels whose CTAs may sit next to a running kernel: ONE thread blocks in griddepcontrol.wait, the others at a
// CTA barrier (GAIT_PDL_WAIT_ONE=0 restores the all-threads wait)
#ifndef GAIT_PDL_WAIT_ONE
#define GAIT_PDL_WAIT_ONE 1
#endif
__device__ __forceinline__ void pdl_wait_cta() {
#if GAIT_PDL_WAIT_ONE
    if (threadIdx.x == 0) pdl_wait();
    __syncthreads();
#else
    pdl_wait();
#endif
}
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int kind, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl_mask() & kind) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// internal (not exported) variants used by composite entry points
int linear_launch(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                  const float* Cin, int64_t ldcin, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                  cudaStream_t stream);

int linear_simt_launch(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                       const float* Cin, int64_t ldcin, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                       cudaStream_t stream);
// tcgen05 split-TF32 path (linear_tc.cu)
bool linear_tc_eligible(const float* A, int64_t lda, const float* W, int64_t ldw, int64_t M, int64_t N, int64_t K);
int linear_tc_launch(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, const float* Cin,
                     int64_t ldcin, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int splits,
                     int64_t split_stride, cudaStream_t stream);
int linear_path();
// prepared (pre-split) weight lookup of the tensor-core GEMM; an explicit operand set with PreparedOverride wins over the registry
struct PreparedOverride {
    const float* prev_w; const float* prev_hilo; int64_t prev_n;
    PreparedOverride(const float* W, const float* hilo, int64_t n);
    ~PreparedOverride();
};
// Call-site knob of the tensor-core GEMM: k-blocks (32 k) accumulated in tensor memory between promotions to FP32 registers
// (0 = by K: 1 for K >= 512, else 2).  1 is the more accurate, 2 the faster setting (see linear_tc.cu).
extern thread_local int g_linear_promote_kb;
struct LinearPromote {
    int prev;
    explicit LinearPromote(int kb) : prev(g_linear_promote_kb) { g_linear_promote_kb = kb; }
    ~LinearPromote() { g_linear_promote_kb = prev; }
};
// tma.cu: CUtensorMap (128-byte opaque) builder
int make_tensor_map_2d(void* map, int elem_bytes, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t row_stride_bytes,
                       uint32_t box0, uint32_t box1, bool swizzle128);

int make_tensor_map_3d_f32(void* map, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t dim2, uint64_t stride1_bytes,
                           uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2, bool swizzle128);
// internal return code of gru_recurrent_launch: the launch was refused, run the per-step path instead
constexpr int GAIT_GRU_RETRY_PER_STEP = 1;
// gru_rec.cu: the whole recurrence of one GRU layer/direction in one persistent cluster kernel
bool gru_recurrent_eligible(const float* gi, const float* W_hh, const float* h0, const float* y, int64_t ldy,
                            const float* resid, int64_t ldres, const float* out, int64_t ldout, int64_t S, int64_t T,
                            int64_t H);
int gru_recurrent_launch(const float* gi, const float* W_hh, const float* b_hh, const float* h0, float* y, int64_t ldy,
                         const float* resid, int64_t ldres, float* out, int64_t ldout, float* hn, int64_t S, int64_t T,
                         int64_t H, int reverse, unsigned int* counter, float* hlo, cudaStream_t stream);

// gru_small.cu: weight-stationary recurrence for 1-2 sequences (W_hh resident in registers + shared memory for all T steps)
bool gru_small_eligible(const float* gi, const float* W_hh, const float* h0, const float* y, int64_t ldy, int64_t S, int64_t T,
                        int64_t H);
int gru_small_launch(const float* gi, const float* W_hh, const float* b_hh, const float* h0, float* y, int64_t ldy,
                     const float* resid, int64_t ldres, float* out, int64_t ldout, float* hn, int64_t S, int64_t T, int reverse,
                     void* exchange, cudaStream_t stream);

}  // namespace gait
