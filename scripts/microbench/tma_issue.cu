// Microbenchmark: is the TMA request rate of an SM limited per issuing thread?  128 CTAs stream an L2-resident buffer
// with cp.async.bulk (1D) issued by `warps` warps x `lanes` lanes, each issuer with its own 2-stage ring.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
constexpr int STAGES = 2;
// 2D tensor loads: matrix ROWS x 2048 floats; CTA owns ROWS/gridDim rows; issuer `me` loads boxes (32 floats x box_rows)
__global__ void k2d(const __grid_constant__ CUtensorMap tm, int rows_per_cta, int box_rows, int lanes, int reps, unsigned long long* clk, int issue_warps, int poll_mode) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = issue_warps;
    const int n_issuers = warps * lanes, me = warp * lanes + lane;
    const int chunk = box_rows * 128;
    const uint32_t bars = base + n_issuers * STAGES * chunk;
    const uint32_t dummy = bars + 8 * n_issuers * STAGES;
    volatile int* done = reinterpret_cast<volatile int*>(smem_raw + (base - smem_u32(smem_raw)) + n_issuers * STAGES * chunk + 8 * n_issuers * STAGES + 8);
    if (warp < warps && lane < lanes) for (int s = 0; s < STAGES; ++s) mbar_init(bars + 8 * (me * STAGES + s), 1);
    if (threadIdx.x == 0) { mbar_init(dummy, 1); *done = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const unsigned long long t0 = clock64();
    if (warp >= warps) {
        // pollers: spin like consumer warps waiting for a phase that does not complete
        while (!*done) {
            uint32_t ok;
            if (poll_mode == 0)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(dummy), "r"(0) : "memory");
            else
                ok = mbar_test(dummy, 0);
        }
    } else if (lane < lanes) {
        const int boxes_per_kb = rows_per_cta / box_rows, n = boxes_per_kb * 64;      // 64 k-blocks of 32 floats
        const int mine = n / n_issuers, total = mine * reps;
        for (int it = 0; it < total + STAGES; ++it) {
            const int s = it % STAGES;
            const uint32_t bar = bars + 8 * (me * STAGES + s);
            if (it >= STAGES) while (!mbar_test(bar, ((it - STAGES) / STAGES) & 1)) {}
            if (it < total) {
                const int idx = (it % mine) * n_issuers + me, kb = idx / boxes_per_kb, b = idx % boxes_per_kb;
                mbar_expect(bar, chunk);
                tma2d(base + (me * STAGES + s) * chunk, &tm, kb * 32, blockIdx.x * rows_per_cta + b * box_rows, bar);
            }
        }
    }
    if (warp < warps) { asm volatile("bar.sync 1, %0;" ::"r"(32 * warps) : "memory"); if (threadIdx.x == 0) *done = 1; }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = clock64() - t0;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const uint8_t* __restrict__ W, size_t bytes_per_cta, int chunk, int lanes, int reps, unsigned long long* clk) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    const int n_issuers = warps * lanes;
    const int me = warp * lanes + lane;
    const uint32_t bars = base + n_issuers * STAGES * chunk;
    if (lane < lanes) {
        for (int s = 0; s < STAGES; ++s) mbar_init(bars + 8 * (me * STAGES + s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const unsigned long long t0 = clock64();
    if (lane < lanes) {
        const uint8_t* src = W + (size_t)blockIdx.x * bytes_per_cta;
        const int n = (int)(bytes_per_cta / chunk);              // chunks per pass, dealt round-robin to issuers
        const int mine = n / n_issuers, total = mine * reps;
        for (int it = 0; it < total + STAGES; ++it) {
            const int s = it % STAGES;
            const uint32_t bar = bars + 8 * (me * STAGES + s);
            if (it >= STAGES) while (!mbar_test(bar, ((it - STAGES) / STAGES) & 1)) {}
            if (it < total) {
                mbar_expect(bar, chunk);
                bulk1d(base + (me * STAGES + s) * chunk, src + (size_t)((it % mine) * n_issuers + me) * chunk, chunk, bar);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = clock64() - t0;
}
int main() {
    const size_t total = 48u << 20;                                // 48 MB, L2 resident
    const int NCTA = 128;
    uint8_t* W; CK(cudaMalloc(&W, total)); CK(cudaMemset(W, 0, total));
    unsigned long long* clk; CK(cudaMalloc(&clk, 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int reps = 10;
    printf("%-60s %9s %9s %10s %12s\n", "config", "us", "TB/s", "B/clk/SM", "clk/instr/SM");
    for (int chunk : {8192}) for (int warps : {4}) for (int lanes : {1}) {
        if ((size_t)warps * lanes * STAGES * chunk > 200 * 1024) continue;
        const int smem = warps * lanes * STAGES * chunk + warps * lanes * STAGES * 8 + 1024 + 64;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k<<<NCTA, 32 * warps, smem>>>(W, total / NCTA, chunk, lanes, 1, clk);
        CK(cudaEventRecord(e0));
        k<<<NCTA, 32 * warps, smem>>>(W, total / NCTA, chunk, lanes, reps, clk);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        unsigned long long c; CK(cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost));
        const double bytes = (double)total * reps;
        const double instr_per_sm = bytes / NCTA / chunk;
        char nm[128]; snprintf(nm, sizeof nm, "bulk 1D %5d B chunks, %d warps x %2d lanes issuing", chunk, warps, lanes);
        printf("%-60s %9.1f %9.2f %10.1f %12.1f\n", nm, ms * 1e3, bytes / ms * 1e-9, bytes / NCTA / (double)c, (double)c / instr_per_sm);
    }
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int K = 2048;
    const int rows_per_cta = 32;          // 128 CTAs x 32 rows = 4096 rows = 33.5 MB
    for (int pollers : {0, 8, 16}) for (int poll_mode : {0, 1}) for (int box_rows : {32}) for (int warps : {2}) for (int lanes : {1, 4}) {
        if (pollers == 0 && poll_mode == 1) continue;
        if (rows_per_cta / box_rows * 64 % (warps * lanes)) continue;
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)4096}; cuuint64_t strides[1] = {(cuuint64_t)K * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, W, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int chunk = box_rows * 128;
        const int smem = warps * lanes * STAGES * chunk + warps * lanes * STAGES * 8 + 1024 + 64;
        if (smem > 220 * 1024) continue;
        CK(cudaFuncSetAttribute(k2d, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k2d<<<NCTA, 32 * (warps + pollers), smem>>>(tm, rows_per_cta, box_rows, lanes, 1, clk, warps, poll_mode);
        CK(cudaEventRecord(e0));
        k2d<<<NCTA, 32 * (warps + pollers), smem>>>(tm, rows_per_cta, box_rows, lanes, reps, clk, warps, poll_mode);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        unsigned long long c; CK(cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost));
        const double bytes = (double)NCTA * rows_per_cta * K * 4 * reps;
        const double instr_per_sm = bytes / NCTA / chunk;
        char nm[128]; snprintf(nm, sizeof nm, "TMA 2D 128Bx%2d rows, %dw x %2dl issuing, %2d poller warps (%s)", box_rows, warps, lanes, pollers, poll_mode ? "test_wait" : "try_wait");
        printf("%-60s %9.1f %9.2f %10.1f %12.1f\n", nm, ms * 1e3, bytes / ms * 1e-9, bytes / NCTA / (double)c, (double)c / instr_per_sm);
    }
    return 0;
}
