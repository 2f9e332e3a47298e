// The GRU recurrence for ONE or TWO sequences (BASELINE configs[3]: a single long gait clip, T up to 900 frames; demo.py:149
// feeds one person track at a time) as a weight-stationary persistent kernel.
//
// With one sequence a recurrence step is a matrix-vector product: every one of the 3H x H = 12.6 M weights of W_hh is used
// exactly once per step, so the step time is the time to get 50 MB of weights to the FMA units.  The tensor-core kernel
// (gru_rec.cu) streams them from L2 every step with 63 of its 64 operand rows as padding: ~20 us per step.  Here W_hh never
// moves after the first step: 128 CTAs (one per SM, co-resident, cooperative launch) each own 16 hidden units = 48 rows of
// W_hh = 393 KB, kept where the B200 SM has room for it -
//     * 16 (12 for two sequences) of the 48 rows in REGISTERS (64 registers per thread, 131 KB per SM),
//     * 26 rows in SHARED MEMORY (213 KB per SM),
//     * the remaining 6 (10) rows are re-read from L2 each step (49 KB per SM; the loads are issued BEFORE the step's
//       flag wait, so their latency hides under the inter-CTA synchronisation).
// K is split over the 16 warps (a warp owns 128 columns, a lane one float4 of every row), so a warp needs only its own 512-byte
// chunk of h_{t-1} - produced by 8 CTAs - polls only those 8 step counters, takes the chunk from L2 straight into registers and
// starts its FMAs without any CTA-wide wait; a transposed-butterfly shuffle reduction and ONE CTA barrier per step bring the
// 16 partial sums of every row to the 16 x S gate threads.  Exact FP32 FMA arithmetic (no tensor cores, no operand split).
// Steps are chained through the data itself: h_t travels between the CTAs as 8-byte (value, step tag) pairs written with
// single 64-bit stores and polled by the readers - no flag, no release / acquire fence (the flag protocol of gru_rec.cu
// costs ~3.5 us per step: stores, fence, barrier, release store, propagation, acquire poll, reload).  (The first version gave each warp one hidden unit and all of K: every warp then read
// all of h from shared memory - 128 KB of smem traffic per step on top of the weights - behind two CTA barriers and a poll of
// all 128 counters: 4.4 us per step.)
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gait {
namespace grusmall {
using tcu::SpinGuard;

constexpr int HH = 2048;                    // hidden size this kernel is laid out for
constexpr int UC = 16;                      // hidden units per CTA (48 rows of W_hh: gates r, z, n)
constexpr int NCTA = HH / UC;               // 128
constexpr int NWARP = 16;                   // warp w owns k-chunk w: k in [128 w, 128 w + 128), one float4 per lane
constexpr int THREADS = NWARP * 32;
constexpr int NROW = 3 * UC;                // 48
static_assert(NWARP * 128 == HH && NCTA * UC == HH, "k-chunks and unit blocks tile the hidden state");

// rows of the CTA's (48 x 2048) weight slice by where they live: registers / shared memory / re-read from L2 every step
template <int SB> struct Cfg {
    static constexpr int RR = (SB == 1) ? 16 : 12;
    static constexpr int SR = 26;
    static constexpr int GR = NROW - RR - SR;               // 6 (one sequence) or 10 (two)
    static constexpr int W_BYTES = SR * NWARP * 512;        // 212 992
    static constexpr int RED_BYTES = 2 * NWARP * NROW * SB * 4;     // per-warp partial sums, double-buffered by step parity
    static constexpr int SMEM = W_BYTES + RED_BYTES;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
// 16 per-lane values -> their 32-lane sums, transposed butterfly (each level a lane keeps one half of its values and sends
// the other): 16 shuffles; afterwards lane l holds the sum of value 8 b16 + 4 b8 + 2 b4 + b2 (b = bits of l)
__device__ __forceinline__ float reduce16(const float (&v)[16], int lane) {
    const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2;
    float a[8], b[4], c[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (b16 ? v[i + 8] : v[i]) + __shfl_xor_sync(0xffffffffu, b16 ? v[i] : v[i + 8], 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = (b8 ? a[i + 4] : a[i]) + __shfl_xor_sync(0xffffffffu, b8 ? a[i] : a[i + 4], 8);
#pragma unroll
    for (int i = 0; i < 2; ++i) c[i] = (b4 ? b[i + 2] : b[i]) + __shfl_xor_sync(0xffffffffu, b4 ? b[i] : b[i + 2], 4);
    float r = (b2 ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, b2 ? c[0] : c[1], 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

// K is split over the warps: warp w needs only chunk w of h_{t-1} - the 128 units that CTAs 8w .. 8w+7 publish - so it polls
// those eight step counters, loads its 512 bytes of h straight from L2 into registers (no shared-memory staging of h, no CTA
// barrier before the arithmetic) and forms the partial sums of all 48 rows over its chunk.  ONE CTA barrier per step then
// separates the per-warp partials (shared memory, double-buffered by step parity) from the 16 x S gate threads of warp 0, which
// add them, apply the gates, store h_t and publish the CTA's counter.
template <int SB>
__global__ void __launch_bounds__(THREADS, 1)
gru_small_kernel(const float* __restrict__ gi, const float* __restrict__ W_hh, const float* __restrict__ b_hh,
                 const float* __restrict__ h0, float* y, int64_t ldy, const float* __restrict__ resid, int64_t ldres,
                 float* __restrict__ out, int64_t ldout, float* __restrict__ hn, int S, int T, int reverse, uint2* hx) {
    using C = Cfg<SB>;
    extern __shared__ __align__(16) uint8_t smem[];
    float4* Ws = reinterpret_cast<float4*>(smem);                        // [warp][SR rows][lane]
    float* red = reinterpret_cast<float*>(smem + C::W_BYTES);            // [parity][warp][row][SB]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int u0 = blockIdx.x * UC;
    // local row lr = gate * 16 + unit  ->  row of W_hh: gate * HH + u0 + unit; this thread's float4: columns 128 warp + 4 lane
    auto wrow = [&](int lr) {
        return reinterpret_cast<const float4*>(W_hh + ((int64_t)(lr / UC) * HH + u0 + (lr % UC)) * HH + 128 * warp) + lane;
    };

    // ---- weights become resident (once; 50 MB from HBM for the whole grid): rows [0, RR) in registers, [RR, RR + SR) in smem
    float4 wreg[C::RR];
#pragma unroll
    for (int r = 0; r < C::RR; ++r) wreg[r] = __ldg(wrow(r));
#pragma unroll 1
    for (int r = 0; r < C::SR; ++r) Ws[(warp * C::SR + r) * 32 + lane] = __ldg(wrow(C::RR + r));

    // gate threads: warp 0, thread = (sequence, unit)
    const bool gate = tid < UC * S;
    const int gu = tid % UC, gs = tid / UC;
    float bias[3] = {0.f, 0.f, 0.f}, hprev = 0.f;
    if (gate) {
#pragma unroll
        for (int g = 0; g < 3; ++g) bias[g] = b_hh[g * HH + u0 + gu];
        if (h0) hprev = h0[(int64_t)gs * HH + u0 + gu];
    }
    __syncthreads();

    for (int step = 0; step < T; ++step) {
        const int t = reverse ? (T - 1 - step) : step;
        // ---- independent of h_{t-1}: the streamed weight rows and this frame's input projection, issued before the wait
        float4 wst[C::GR];
#pragma unroll
        for (int r = 0; r < C::GR; ++r) wst[r] = __ldg(wrow(C::RR + C::SR + r));
        float gin[3] = {0.f, 0.f, 0.f}, rs = 0.f;
        if (gate) {
            const int64_t f = (int64_t)gs * T + t;
#pragma unroll
            for (int g = 0; g < 3; ++g) gin[g] = gi[f * 3 * HH + g * HH + u0 + gu];
            if (out) rs = resid[f * ldres + u0 + gu];
        }
        // ---- this warp's chunk of h_{t-1}: every value travels as an 8-byte pair (h, tag = step it is the input of), written
        // with one 64-bit store (single-copy atomic), so the reader polls THE DATA: no flag, no release / acquire fence, one
        // L2 round trip after the value becomes visible.  Two exchange buffers alternate by step parity: a CTA cannot reach
        // the gates of step t+2 before every CTA has consumed all of h_t (it needs their h_{t+1}).
        float4 hv[SB];
        if (step > 0) {
#pragma unroll
            for (int s = 0; s < SB; ++s) {
                hv[s] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (s < S) {
                    const uint4* src = reinterpret_cast<const uint4*>(hx + ((size_t)((step & 1) * SB + s)) * HH + 128 * warp + 4 * lane);
                    SpinGuard guard;
                    for (;;) {
                        // one lane spins on one pair of the chunk (512 threads x 128 CTAs polling their own data slowed the
                        // writers down: 8 us per step); when it has arrived the others usually have too, and every lane
                        // still verifies the tags of the values it uses
                        if (lane == 0) {
                            const unsigned* probe = reinterpret_cast<const unsigned*>(hx + ((size_t)((step & 1) * SB + s)) * HH + 128 * warp + 127) + 1;
                            unsigned tg;
                            do {
                                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(tg) : "l"(probe) : "memory");
                                if (tg != (unsigned)step) guard.tick();
                            } while (tg != (unsigned)step);
                        }
                        __syncwarp();
                        uint4 p0, p1;
                        asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(p0.x), "=r"(p0.y), "=r"(p0.z), "=r"(p0.w) : "l"(src) : "memory");
                        asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(p1.x), "=r"(p1.y), "=r"(p1.z), "=r"(p1.w) : "l"(src + 1) : "memory");
                        if (p0.y == (unsigned)step && p0.w == (unsigned)step && p1.y == (unsigned)step && p1.w == (unsigned)step) {
                            hv[s] = make_float4(__uint_as_float(p0.x), __uint_as_float(p0.z), __uint_as_float(p1.x), __uint_as_float(p1.z));
                        }
                        const bool all_ok = __all_sync(0xffffffffu, p0.y == (unsigned)step && p0.w == (unsigned)step &&
                                                                    p1.y == (unsigned)step && p1.w == (unsigned)step);
                        if (all_ok) break;
                        guard.tick();
                    }
                }
            }
        } else {
#pragma unroll
            for (int s = 0; s < SB; ++s)
                hv[s] = (h0 && s < S) ? __ldg(reinterpret_cast<const float4*>(h0 + (int64_t)s * HH + 128 * warp) + lane)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // ---- partial sums of the 48 rows over this chunk, one gate (16 rows) at a time
        float* myred = red + ((step & 1) * NWARP + warp) * (NROW * SB);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
#pragma unroll
            for (int s = 0; s < SB; ++s) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int lr = g * 16 + i;
                    float4 w;
                    if (lr < C::RR) w = wreg[lr];
                    else if (lr < C::RR + C::SR) w = Ws[(warp * C::SR + (lr - C::RR)) * 32 + lane];
                    else w = wst[lr - C::RR - C::SR];
                    v[i] = dot4(w, hv[s]);
                }
                const float r = reduce16(v, lane);
                const int ui = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                if ((lane & 1) == 0) myred[(g * 16 + ui) * SB + s] = r;
            }
        }
        __syncthreads();                            // the only CTA barrier of a step: all 16 partials of every row are in smem
        // ---- gates: thread (sequence, unit) of warp 0 (torch.nn.GRU: r, z, n; h' = (1-z) n + z h)
        if (gate) {
            const float* rp = red + (step & 1) * NWARP * (NROW * SB);
            float a[3];
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                float acc = 0.f;
#pragma unroll
                for (int w = 0; w < NWARP; ++w) acc += rp[w * (NROW * SB) + (g * 16 + gu) * SB + gs];     // fixed order
                a[g] = acc;
            }
            const float r = sigmoidf_(gin[0] + a[0] + bias[0]);
            const float z = sigmoidf_(gin[1] + a[1] + bias[1]);
            const float n = tanhf(gin[2] + r * (a[2] + bias[2]));
            const float h = (1.f - z) * n + z * hprev;
            hprev = h;
            if (step < T - 1) {                     // first: what the other CTAs wait for - (h_t, tag of the step that consumes it)
                uint2* dst = hx + ((size_t)(((step + 1) & 1) * SB + gs)) * HH + u0 + gu;
                asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(__float_as_uint(h)), "r"((unsigned)(step + 1)) : "memory");
            }
            const int64_t f = (int64_t)gs * T + t;
            y[f * ldy + u0 + gu] = h;
            if (out) out[f * ldout + u0 + gu] = h + rs;
            if (hn && step == T - 1) hn[(int64_t)gs * HH + u0 + gu] = h;
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
                     void* exchange, cudaStream_t stream) {
    using namespace grusmall;
    // exchange buffer [2 parities][S][H] of (value, tag) pairs; tag 0 is never awaited (steps are tagged from 1)
    uint2* hx = static_cast<uint2*>(exchange);
    GAIT_CUDA(cudaMemsetAsync(hx, 0, sizeof(uint2) * 2 * (size_t)S * HH, stream));
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
        ? cudaLaunchKernelEx(&cfg, gru_small_kernel<1>, gi, W_hh, b_hh, h0, y, ldy, resid, ldres, out, ldout, hn, Si, Ti, reverse, hx)
        : cudaLaunchKernelEx(&cfg, gru_small_kernel<2>, gi, W_hh, b_hh, h0, y, ldy, resid, ldres, out, ldout, hn, Si, Ti, reverse, hx);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("gru(weight-stationary): cooperative launch refused: %s", cudaGetErrorString(e));
        return GAIT_GRU_RETRY_PER_STEP;
    }
    return check_launch("gru(weight-stationary recurrent)");
}

}  // namespace gait
