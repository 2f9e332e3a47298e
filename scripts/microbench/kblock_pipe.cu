// Microbenchmark of the operand-feed pipeline of the persistent GRU kernel: per k-block a CTA needs a 96-row x 128-byte
// tile of W (rows 8 KB apart, 3 groups of 32 rows) and a 64-row x 128-byte tile of h (rows 128 KB apart), both L2
// resident, into a ring of `stages` buffers; a consumer thread releases a stage as soon as it has landed.
// Strategies differ in how the TMA instructions are cut and who issues them.  Reports cycles per k-block.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
constexpr int NKB = 32, WROWS = 96, HROWS = 64, STAGE = (WROWS + HROWS) * 128;

// rows_per_box: box height of both tensor maps.  nwarps producer warps alternate k-blocks; within a warp, one lane per box.
// serial != 0: a single lane issues all boxes of a k-block one after the other (the original scheme).
__global__ void pipe(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH, int rows_per_box,
                     int nwarps, int serial, int stages, int steps, unsigned long long* clk) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + stages * STAGE;
    auto FULL = [&](int s) { return bars + 8u * s; };
    auto EMPTY = [&](int s) { return bars + 8u * (stages + s); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wboxes = WROWS / rows_per_box, hboxes = HROWS / rows_per_box, nboxes = wboxes + hboxes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(FULL(s), serial ? 1 : nboxes); mbar_init(EMPTY(s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned long long t0 = clock64();
    const int total = NKB * steps;
    const int kslice = (blockIdx.x & 1) * 1024, rowbase = (blockIdx.x >> 1) * 32;
    if (warp < nwarps) {
        for (int it = warp; it < total; it += nwarps) {
            const int s = it % stages, kb = it % NKB;
            mbar_wait(EMPTY(s), ((it / stages) & 1) ^ 1);
            const uint32_t st = base + s * STAGE;
            if (serial) {
                if (lane == 0) {
                    mbar_expect(FULL(s), STAGE);
                    for (int b = 0; b < wboxes; ++b) {
                        const int r = b * rows_per_box, g = r / 32, rr = r % 32;
                        tma2d(st + r * 128, &tmW, kslice + kb * 32, g * 2048 + rowbase + rr, FULL(s));
                    }
                    for (int b = 0; b < hboxes; ++b) tma2d(st + (WROWS + b * rows_per_box) * 128, &tmH, kslice + kb * 32, b * rows_per_box, FULL(s));
                }
            } else if (lane < nboxes) {
                mbar_expect(FULL(s), rows_per_box * 128);
                if (lane < wboxes) {
                    const int r = lane * rows_per_box, g = r / 32, rr = r % 32;
                    tma2d(st + r * 128, &tmW, kslice + kb * 32, g * 2048 + rowbase + rr, FULL(s));
                } else {
                    const int b = lane - wboxes;
                    tma2d(st + (WROWS + b * rows_per_box) * 128, &tmH, kslice + kb * 32, b * rows_per_box, FULL(s));
                }
            }
            __syncwarp();
        }
    } else if (warp == nwarps) {
        for (int it = 0; it < total; ++it) {
            const int s = it % stages;
            mbar_wait(FULL(s), (it / stages) & 1);
            if (lane == 0) mbar_arrive(EMPTY(s));
            __syncwarp();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = clock64() - t0;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    float *W, *Hh;
    CK(cudaMalloc(&W, (size_t)6144 * 2048 * 4)); CK(cudaMemset(W, 0, (size_t)6144 * 2048 * 4));
    CK(cudaMalloc(&Hh, (size_t)64 * 16 * 2048 * 4)); CK(cudaMemset(Hh, 0, (size_t)64 * 16 * 2048 * 4));
    unsigned long long* clk; CK(cudaMalloc(&clk, 8));
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    printf("%-70s %12s %10s\n", "strategy", "clk/k-block", "B/clk/SM");
    for (int stages : {4, 5}) for (int serial : {1, 0}) for (int rows_per_box : {32, 16, 8}) for (int nwarps : {1, 2}) {
        if (serial && rows_per_box != 32) continue;
        CUtensorMap tmW, tmH;
        cuuint64_t dW[2] = {2048, 6144}, sW[1] = {2048 * 4}, dH[2] = {2048, 64}, sH[1] = {16 * 2048 * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)rows_per_box}, es[2] = {1, 1};
        if (enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, W, dW, sW, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
            enc(&tmH, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, Hh, dH, sH, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
        const int smem = stages * STAGE + 1024 + 256;
        CK(cudaFuncSetAttribute(pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int steps = 15;
        pipe<<<128, 32 * (nwarps + 1), smem>>>(tmW, tmH, rows_per_box, nwarps, serial, stages, 2, clk);
        pipe<<<128, 32 * (nwarps + 1), smem>>>(tmW, tmH, rows_per_box, nwarps, serial, stages, steps, clk);
        CK(cudaDeviceSynchronize());
        unsigned long long c; CK(cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost));
        char nm[160];
        snprintf(nm, sizeof nm, "%d stages, %s, boxes of %2d rows, %d producer warp(s)", stages, serial ? "one lane issues all boxes" : "one lane per box      ", rows_per_box, nwarps);
        printf("%-70s %12.1f %10.1f\n", nm, (double)c / (NKB * steps), (double)STAGE * NKB * steps / c);
    }
    return 0;
}
