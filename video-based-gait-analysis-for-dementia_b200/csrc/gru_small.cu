// The GRU recurrence for ONE or TWO sequences (BASELINE configs[3]: a single long gait clip, T up to 900 frames; demo.py:149
// feeds one person track at a time) as a weight-stationary persistent kernel.
//
// With one sequence a recurrence step is a matrix-vector product: every one of the 3H x H = 12.6 M weights of W_hh is used
// exactly once per step, so the step time is the time to get 50 MB of weights to the FMA units.  The tensor-core kernel
// (gru_rec.cu) streams them from L2 every step with 63 of its 64 operand rows as padding: ~20 us per step.  Here W_hh never
// moves after the first step: 128 CTAs (one per SM, co-resident, cooperative launch) each own 16 hidden units = 48 rows of
// W_hh = 393 KB, kept where the B200 SM has room for it -
//     * 4 of the 16 k-chunks of every row in REGISTERS  (48 registers per thread, 98 KB per SM),
//     * 9 (8 for two sequences) k-chunks in SHARED MEMORY (221 KB per SM),
//     * the remaining 3 (4) k-chunks are re-read from L2 each step (75 KB per SM; the loads are issued BEFORE the step's
//       flag wait, so their latency hides under the inter-CTA synchronisation).
// A warp owns one hidden unit (its r, z and n rows), lanes split K, so the three gate pre-activations of a unit meet in one
// warp after a shuffle reduction and the gate math needs no cross-warp exchange.  Exact FP32 FMA arithmetic (no tensor cores,
// no operand split).  Steps are chained like in gru_rec.cu: every CTA publishes a step counter (st.release, own 128-byte line)
// after its 16 units of h_t are in global memory; 128 threads of every CTA poll the 128 counters, then the CTA reloads h_t
// (8 KB per sequence, L2) into shared memory.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gait {
namespace grusmall {
using tcu::SpinGuard;

constexpr int HH = 2048;                    // hidden size this kernel is laid out for
constexpr int UC = 16;                      // hidden units (= warps) per CTA
constexpr int NCTA = HH / UC;               // 128
constexpr int THREADS = UC * 32;
constexpr int NCH = HH / 128;               // 16 k-chunks of 128 floats (one float4 per lane)
constexpr int RC = 4;                       // chunks held in registers
constexpr int FLAG_STRIDE = 32;             // words between the step counters of consecutive CTAs

template <int SB> struct Cfg {
    static constexpr int SC = (SB == 1) ? 9 : 8;            // chunks held in shared memory
    static constexpr int GC = NCH - RC - SC;                // chunks streamed from L2 every step
    static constexpr int W_BYTES = 3 * UC * SC * 512;
    static constexpr int H_BYTES = SB * HH * 4;
    static constexpr int SMEM = W_BYTES + H_BYTES;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}

template <int SB>
__global__ void __launch_bounds__(THREADS, 1)
gru_small_kernel(const float* __restrict__ gi, const float* __restrict__ W_hh, const float* __restrict__ b_hh,
                 const float* __restrict__ h0, float* y, int64_t ldy, const float* __restrict__ resid, int64_t ldres,
                 float* __restrict__ out, int64_t ldout, float* __restrict__ hn, int S, int T, int reverse, unsigned* flags) {
    using C = Cfg<SB>;
    extern __shared__ __align__(16) uint8_t smem[];
    float4* Ws = reinterpret_cast<float4*>(smem);                        // [(warp*3+g)][chunk][lane]
    float* hs = reinterpret_cast<float*>(smem + C::W_BYTES);             // [SB][HH]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int unit = blockIdx.x * UC + warp;

    // ---- weights become resident: registers and shared memory (once; 50 MB from HBM for the whole grid)
    float4 wreg[3][RC];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        const float4* row = reinterpret_cast<const float4*>(W_hh + ((int64_t)g * HH + unit) * HH) + lane;
#pragma unroll
        for (int c = 0; c < RC; ++c) wreg[g][c] = __ldg(row + c * 32);
#pragma unroll
        for (int c = 0; c < C::SC; ++c) Ws[((warp * 3 + g) * C::SC + c) * 32 + lane] = __ldg(row + (RC + c) * 32);
    }
    float bias[3], hprev = 0.f;                                          // lane s < S handles the gate math of sequence s
#pragma unroll
    for (int g = 0; g < 3; ++g) bias[g] = b_hh[g * HH + unit];
    if (lane < S && h0) hprev = h0[(int64_t)lane * HH + unit];
    for (int i = tid; i < SB * HH; i += THREADS) {
        const int s = i / HH;
        hs[i] = (h0 && s < S) ? h0[(int64_t)s * HH + (i % HH)] : 0.f;
    }
    __syncthreads();

    for (int step = 0; step < T; ++step) {
        const int t = reverse ? (T - 1 - step) : step;
        // ---- independent of h_{t-1}: the streamed weight chunks and this frame's input projection, issued before the wait
        float4 wst[3][C::GC];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            const float4* row = reinterpret_cast<const float4*>(W_hh + ((int64_t)g * HH + unit) * HH) + lane;
#pragma unroll
            for (int c = 0; c < C::GC; ++c) wst[g][c] = __ldg(row + (RC + C::SC + c) * 32);
        }
        float gin[3] = {0.f, 0.f, 0.f}, rs = 0.f;
        if (lane < S) {
            const int64_t f = (int64_t)lane * T + t;
#pragma unroll
            for (int g = 0; g < 3; ++g) gin[g] = gi[f * 3 * HH + g * HH + unit];
            if (out) rs = resid[f * ldres + unit];
        }
        if (step > 0) {
            // ---- h_{t-1} is complete when every CTA has published `step`
            if (tid < NCTA) {
                const unsigned* fl = flags + (size_t)tid * FLAG_STRIDE;
                SpinGuard guard;
                for (;;) {
                    unsigned v;
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(fl) : "memory");
                    if (v >= (unsigned)step) break;
                    guard.tick();
                }
            }
            __syncthreads();
            const int tp = reverse ? t + 1 : t - 1;
            for (int i = tid; i < S * (HH / 4); i += THREADS) {
                const int s = i / (HH / 4), k4 = i % (HH / 4);
                const float4 v = __ldcg(reinterpret_cast<const float4*>(y + ((int64_t)s * T + tp) * ldy) + k4);   // L2, never a stale L1 line
                reinterpret_cast<float4*>(hs + s * HH)[k4] = v;
            }
            __syncthreads();
        }
        // ---- 3 gate rows x SB sequences, K split over the lanes
        float acc[3][SB];
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int s = 0; s < SB; ++s) acc[g][s] = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            float4 hv[SB];
#pragma unroll
            for (int s = 0; s < SB; ++s) hv[s] = reinterpret_cast<const float4*>(hs + s * HH)[c * 32 + lane];
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                float4 w;
                if (c < RC) w = wreg[g][c];
                else if (c < RC + C::SC) w = Ws[((warp * 3 + g) * C::SC + (c - RC)) * 32 + lane];
                else w = wst[g][c - RC - C::SC];
#pragma unroll
                for (int s = 0; s < SB; ++s) acc[g][s] = dot4(w, hv[s], acc[g][s]);
            }
        }
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int s = 0; s < SB; ++s)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[g][s] += __shfl_xor_sync(0xffffffffu, acc[g][s], o);
        // ---- gates: lane s finalises sequence s of this warp's unit (torch.nn.GRU: r, z, n; h' = (1-z) n + z h)
        if (lane < S) {
            float ar = acc[0][0], az = acc[1][0], an = acc[2][0];
#pragma unroll
            for (int s = 1; s < SB; ++s)
                if (lane == s) { ar = acc[0][s]; az = acc[1][s]; an = acc[2][s]; }
            const float r = sigmoidf_(gin[0] + ar + bias[0]);
            const float z = sigmoidf_(gin[1] + az + bias[1]);
            const float n = tanhf(gin[2] + r * (an + bias[2]));
            const float h = (1.f - z) * n + z * hprev;
            hprev = h;
            const int64_t f = (int64_t)lane * T + t;
            y[f * ldy + unit] = h;
            if (out) out[f * ldout + unit] = h + rs;
            if (hn && step == T - 1) hn[(int64_t)lane * HH + unit] = h;
        }
        if (step < T - 1) {
            __syncthreads();                       // the CTA's 16 units of h_t are stored; everyone is done reading hs
            if (tid == 0)           // release: cumulative over the CTA's stores of h_t ordered before it by the barrier
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + (size_t)blockIdx.x * FLAG_STRIDE), "r"((unsigned)(step + 1)) : "memory");
        }
    }
}

}  // namespace grusmall

bool gru_small_eligible(const float* gi, const float* W_hh, const float* h0, const float* y, int64_t ldy, int64_t S, int64_t T,
                        int64_t H) {
    using namespace grusmall;
    if (S < 1 || S > 2 || T < 1 || H != HH) return false;
    if (!aligned16(W_hh) || !aligned16(y) || (ldy & 3)) return false;
    (void)gi; (void)h0;
    return device_sm_count() >= NCTA;
}

// GAIT_GRU_RETRY_PER_STEP when the cooperative launch is refused (the caller then takes a path without a residency requirement)
int gru_small_launch(const float* gi, const float* W_hh, const float* b_hh, const float* h0, float* y, int64_t ldy,
                     const float* resid, int64_t ldres, float* out, int64_t ldout, float* hn, int64_t S, int64_t T, int reverse,
                     unsigned int* flags, cudaStream_t stream) {
    using namespace grusmall;
    GAIT_CUDA(cudaMemsetAsync(flags, 0, sizeof(unsigned int) * (size_t)NCTA * FLAG_STRIDE, stream));
    static PerDeviceOnce attr_once;
    int dev = 0;
    if (attr_once.needed(&dev)) {
        GAIT_CUDA(cudaFuncSetAttribute(gru_small_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM));
        GAIT_CUDA(cudaFuncSetAttribute(gru_small_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2>::SMEM));
        attr_once.mark(dev);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(NCTA);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = S == 1 ? Cfg<1>::SMEM : Cfg<2>::SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;            // all 128 CTAs resident at once (they spin on each other's flags)
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const int Si = (int)S, Ti = (int)T;
    cudaError_t e = S == 1
        ? cudaLaunchKernelEx(&cfg, gru_small_kernel<1>, gi, W_hh, b_hh, h0, y, ldy, resid, ldres, out, ldout, hn, Si, Ti, reverse, flags)
        : cudaLaunchKernelEx(&cfg, gru_small_kernel<2>, gi, W_hh, b_hh, h0, y, ldy, resid, ldres, out, ldout, hn, Si, Ti, reverse, flags);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("gru(weight-stationary): cooperative launch refused: %s", cudaGetErrorString(e));
        return GAIT_GRU_RETRY_PER_STEP;
    }
    return check_launch("gru(weight-stationary recurrent)");
}

}  // namespace gait
