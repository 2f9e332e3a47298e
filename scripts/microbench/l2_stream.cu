// Microbenchmark: how fast can the SMs stream an L2-resident K-major FP32 operand (rows 8 KB apart, 128-byte row
// segments per k-block) - with LDG.128, and with TMA 2D SWIZZLE_128B boxes into a shared-memory ring?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o l2_stream l2_stream.cu   Run: ./l2_stream
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
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// matrix: ROWS x K floats (row-major). CTA b owns rows [b*R, (b+1)*R); sweeps k-blocks of `bk` floats, `reps` times.
// mode 0: LDG.128, each warp reads whole 128-byte (or bk*4-byte) row segments
__global__ void ldg_kernel(const float* __restrict__ W, int K, int LD, int R, int bk, int reps, float* sink) {
    const float4* base = reinterpret_cast<const float4*>(W + (size_t)blockIdx.x * R * LD);
    const int f4_per_seg = bk / 4;                       // float4 per row segment
    const int segs_per_pass = blockDim.x / f4_per_seg;   // rows covered by the CTA at once
    const int r0 = threadIdx.x / f4_per_seg, c = threadIdx.x % f4_per_seg;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int rep = 0; rep < reps; ++rep)
        for (int kb = 0; kb < K / bk; ++kb) {
#pragma unroll 4
            for (int r = r0; r < R; r += segs_per_pass) {
                const float4 v = __ldcg(base + ((size_t)r * LD + (size_t)kb * bk) / 4 + c);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
}

// mode 1: TMA ring.  One thread issues box loads (bk x box_rows, R/box_rows boxes per k-block) into `stages` buffers and
// re-issues as soon as a buffer has landed (nobody consumes: pure load throughput).
__global__ void tma_kernel(const __grid_constant__ CUtensorMap tm, int K, int R, int bk, int box_rows, int stages, int reps,
                           unsigned long long* lat) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_bytes = (uint32_t)R * bk * 4;
    const uint32_t bars = base + stages * stage_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int nkb = K / bk, total = nkb * reps;
        int issued = 0;
        unsigned long long t_first = 0;
        for (int it = 0; it < total + stages; ++it) {
            const int s = it % stages;
            if (it >= stages) mbar_spin(bars + 8 * s, ((it - stages) / stages) & 1);
            if (it == stages && lat && blockIdx.x == 0) lat[0] = clock64() - t_first;
            if (it < total) {
                const int kb = it % nkb;
                if (it == 0) t_first = clock64();
                mbar_expect(bars + 8 * s, stage_bytes);
                for (int b = 0; b < R / box_rows; ++b)
                    tma2d(base + s * stage_bytes + b * box_rows * bk * 4, &tm, kb * bk, blockIdx.x * R + b * box_rows, bars + 8 * s);
                ++issued;
            }
        }
    }
}

// mode 2: 1D bulk copies of contiguous `chunk`-byte pieces (pre-tiled operand layout)
__global__ void bulk_kernel(const float* __restrict__ W, size_t bytes_per_cta, int chunk, int stages, int reps) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + stages * chunk;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint8_t* src = reinterpret_cast<const uint8_t*>(W) + (size_t)blockIdx.x * bytes_per_cta;
        const int n = (int)(bytes_per_cta / chunk), total = n * reps;
        for (int it = 0; it < total + stages; ++it) {
            const int s = it % stages;
            if (it >= stages) mbar_spin(bars + 8 * s, ((it - stages) / stages) & 1);
            if (it < total) {
                mbar_expect(bars + 8 * s, chunk);
                bulk1d(base + s * chunk, src + (size_t)(it % n) * chunk, chunk, bars + 8 * s);
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int K = argc > 1 ? atoi(argv[1]) : 2048; const int ROWS = 6144, NCTA = 128, R = ROWS / NCTA;      // 48 rows per CTA, 50 MB total: L2 resident
    float* W;
    CK(cudaMalloc(&W, (size_t)ROWS * (K + 64) * 4));
    CK(cudaMemset(W, 0, (size_t)ROWS * (K + 64) * 4));
    float* sink; CK(cudaMalloc(&sink, 4));
    unsigned long long* lat; CK(cudaMalloc(&lat, 8));
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int reps = 20 * (2048 / K);
    const double bytes = (double)ROWS * K * 4 * reps;
    int dev_clk = 0; CK(cudaDeviceGetAttribute(&dev_clk, cudaDevAttrClockRate, 0));
    auto report = [&](const char* name, float ms) {
        printf("%-80s %8.1f us  %7.2f TB/s  %6.1f B/clk/SM\n", name, ms * 1e3, bytes / ms * 1e-9,
               bytes / NCTA / (ms * 1e-3 * dev_clk * 1e3));
    };
    printf("%.1f MB FP32 matrix 6144 x %d, 128 CTAs x 48 rows, %d MHz\n", ROWS * (double)K * 4e-6, K, dev_clk / 1000);
    for (int LD : {K, K + 32}) for (int bk : {32, 128}) {
        const int threads = 1024;
        ldg_kernel<<<NCTA, threads>>>(W, K, LD, R, bk, 2, sink);   // warm
        CK(cudaEventRecord(e0));
        ldg_kernel<<<NCTA, threads>>>(W, K, LD, R, bk, reps, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        char nm[128]; snprintf(nm, sizeof nm, "LDG.128 row segments of %d B, row stride %d B, %d threads", bk * 4, LD * 4, threads);
        report(nm, ms);
    }
    for (int LD : {K, 32}) for (int box_rows : {48, 16}) for (int stages : {2, 8}) {
        const int bk = 32;
        CUtensorMap tm;
        // LD == 32: the same bytes viewed as a fully contiguous (ROWS*K/32) x 32 matrix (pre-tiled operand)
        const cuuint64_t rows = LD == 32 ? (cuuint64_t)ROWS * K / 32 : ROWS;
        cuuint64_t dims[2] = {(cuuint64_t)(LD == 32 ? 32 : K), rows}; cuuint64_t strides[1] = {(cuuint64_t)LD * 4};
        cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, W, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int smem = stages * R * bk * 4 + 1024 + 128;
        CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        // for LD == 32 the kernel's (kb*bk, row) coordinates become (0, tile row): reuse by passing K=32*nkb.. keep simple:
        if (LD == 32) {
            tma_kernel<<<NCTA, 32, smem>>>(tm, 32, R, bk, box_rows, stages, 2 * (K / 32), nullptr);
            CK(cudaEventRecord(e0));
            tma_kernel<<<NCTA, 32, smem>>>(tm, 32, R, bk, box_rows, stages, reps * (K / 32), lat);
        } else {
            tma_kernel<<<NCTA, 32, smem>>>(tm, K, R, bk, box_rows, stages, 2, nullptr);
            CK(cudaEventRecord(e0));
            tma_kernel<<<NCTA, 32, smem>>>(tm, K, R, bk, box_rows, stages, reps, lat);
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        unsigned long long hl; CK(cudaMemcpy(&hl, lat, 8, cudaMemcpyDeviceToHost));
        char nm[160]; snprintf(nm, sizeof nm, "TMA 2D SW128 box 128B x %d rows, row stride %d B, %d stages (t[stage 0 landed] %llu clk)", box_rows, LD * 4, stages, hl);
        report(nm, ms);
    }
    for (int chunk : {2048, 4096, 8192, 12288, 24576}) for (int stages : {2, 4, 8}) {
        const int smem = stages * chunk + 1024 + 128;
        CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        bulk_kernel<<<NCTA, 32, smem>>>(W, (size_t)R * K * 4, chunk, stages, 2);
        CK(cudaEventRecord(e0));
        bulk_kernel<<<NCTA, 32, smem>>>(W, (size_t)R * K * 4, chunk, stages, reps);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        char nm[128]; snprintf(nm, sizeof nm, "cp.async.bulk 1D contiguous %d B chunks, %d stages", chunk, stages);
        report(nm, ms);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
