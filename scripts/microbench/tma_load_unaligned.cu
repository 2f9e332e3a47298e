// Does a TMA tensor LOAD accept a box whose start address is only 8-byte aligned?  (stores fault: tma_store.cu)
// 4-byte elements, box starts at element 2 and at element 20670 of a row-pair view of a (F, 20670) float array.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void load_kernel(const __grid_constant__ CUtensorMap tm, int x, int y, float* out, int n) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), dst = (uint32_t)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n * 4));
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(x), "r"(y), "r"(b) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

int main() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int V3 = 20670, F = 16;
    std::vector<float> h((size_t)F * V3);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 64 * 32 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    for (int sw = 0; sw < 2; ++sw) for (int x : {0, 2, V3, V3 + 32}) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)2 * V3, (cuuint64_t)F / 2};
        cuuint64_t strides[1] = {(cuuint64_t)2 * V3 * 4};
        cuuint32_t box[2] = {32, 8}, es[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        load_kernel<<<1, 128, 8 * 32 * 4 + 1024>>>(tm, x, 1, o, 8 * 32);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("swizzle %d x=%d: KERNEL ERROR %s\n", sw, x, cudaGetErrorString(e)); return 1; }
        float res[4];
        cudaMemcpy(res, o, 16, cudaMemcpyDeviceToHost);
        printf("swizzle %d x=%5d (address %% 16 = %d): ok, first elements %.0f %.0f (expected %.0f)\n", sw, x, (x * 4) % 16, res[0], res[1],
               (float)(2 * V3 + x));
    }
    return 0;
}
