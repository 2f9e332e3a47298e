// Which TMA tensor-store configurations does the hardware accept?  (nvcc -arch=sm_100a tma_store.cu -o tma_store -lcuda)
// Case A: 8-byte elements, aligned box start.  B: box start at an odd element (8-byte aligned address, not 16).
// C: box clipped by the inner extent.  D: 4-byte elements, start at element offset 2 (8-byte aligned).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void store_kernel(const __grid_constant__ CUtensorMap tm, int x, int y, int bytes) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* s = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) s[i] = 1000.f + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t src = (uint32_t)__cvta_generic_to_shared(smem);
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                     ::"l"(reinterpret_cast<uint64_t>(&tm)), "r"(x), "r"(y), "r"(src) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

int main() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int V = 6890, F = 8;
    float* d;
    cudaMalloc(&d, (size_t)F * V * 3 * 4 + 64);
    struct Case { const char* name; int elem; uint64_t base_off_bytes; uint64_t inner; uint64_t rows; uint64_t stride; uint32_t bx, by; int x, y; };
    const uint64_t st2 = (uint64_t)V * 24;
    Case cases[] = {
        {"A even frames, 8B elems, x=192", 8, 0, (uint64_t)V * 3 / 2, F / 2, st2, 192, 4, 192, 0},
        {"B odd frames via shifted base, x=1+192", 8, (uint64_t)V * 12 - 8, (uint64_t)V * 3 / 2 + 1, F / 2, st2, 192, 4, 193, 0},
        {"C last tile clipped, x=53*192", 8, 0, (uint64_t)V * 3 / 2, F / 2, st2, 192, 4, 53 * 192, 0},
        {"D 4B elems, x=2", 4, 0, (uint64_t)V * 3, F / 2, st2, 256, 4, 2, 0},
        {"E 8B elems rows clipped y=2 (rows=4, box 4)", 8, 0, (uint64_t)V * 3 / 2, F / 2, st2, 192, 4, 0, 2},
    };
    for (auto& c : cases) {
        cudaMemset(d, 0, (size_t)F * V * 3 * 4 + 64);
        CUtensorMap tm;
        cuuint64_t dims[2] = {c.inner, c.rows};
        cuuint64_t strides[1] = {c.stride};
        cuuint32_t box[2] = {c.bx, c.by};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tm, c.elem == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                         (uint8_t*)d + c.base_off_bytes, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
        const int bytes = c.bx * c.by * c.elem;
        store_kernel<<<1, 128, bytes>>>(tm, c.x, c.y, bytes);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: KERNEL ERROR %s\n", c.name, cudaGetErrorString(e)); return 1; }
        std::vector<float> h((size_t)F * V * 3 + 16);
        cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
        size_t nz = 0, first = 0, last = 0;
        for (size_t i = 0; i < h.size(); ++i) if (h[i] != 0.f) { if (!nz) first = i; last = i; ++nz; }
        printf("%s: ok, %zu floats written, first at float %zu (value %.0f), last at %zu\n", c.name, nz, first, h[first], last);
    }
    return 0;
}
