// gait_gru_layer: one layer / one direction of torch.nn.GRU (gate order r,z,n;
// h' = (1-z) n + z h), the arithmetic behind TemporalEncoder and
// lib/models/layers/gait_feat_encoder.py:51-57,88.
//
// Input projection for all S*T frames is one GEMM; the recurrence is T steps of
// (S,H).(H,3H) plus a fused gate kernel that also emits the TemporalEncoder residual sum.
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"

namespace gait {

constexpr int kGruMaxSplits = 4;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// gi: (S*T, 3H) input projection incl. b_ih; gh: (S, 3H) hidden projection incl. b_hh, or NULL at
// the first step with h0 == NULL (then gh == b_hh).
__global__ void gru_gate_kernel(const float* __restrict__ gi, const float* __restrict__ gh, int gh_parts,
                                int64_t gh_part_stride, const float* __restrict__ b_hh, const float* __restrict__ hprev, int64_t ldh,
                                float* __restrict__ y, int64_t ldy, const float* __restrict__ resid, int64_t ldres,
                                float* __restrict__ out, int64_t ldout, float* __restrict__ hn, int S, int T,
                                int H, int t) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (u >= H) return;
    const int64_t f = (int64_t)s * T + t;
    const float* g = gi + f * 3 * H;
    float hr, hz, hnn;
    if (gh) {
        const float* q = gh + (int64_t)s * 3 * H;
        hr = q[u]; hz = q[H + u]; hnn = q[2 * H + u];
        for (int part = 1; part < gh_parts; ++part) {         // split-K partial sums of the recurrent GEMM
            q += gh_part_stride;
            hr += q[u]; hz += q[H + u]; hnn += q[2 * H + u];
        }
    } else {
        hr = b_hh[u]; hz = b_hh[H + u]; hnn = b_hh[2 * H + u];
    }
    const float hp = hprev ? hprev[(int64_t)s * ldh + u] : 0.f;
    const float r = sigmoidf_(g[u] + hr);
    const float z = sigmoidf_(g[H + u] + hz);
    const float n = tanhf(g[2 * H + u] + r * hnn);
    const float h = (1.f - z) * n + z * hp;
    y[f * ldy + u] = h;
    if (out) out[f * ldout + u] = h + resid[f * ldres + u];
    if (hn) hn[(int64_t)s * H + u] = h;
}

// Same, four hidden units per thread with 128-bit accesses (H % 4 == 0, every base pointer 16-byte aligned and every row
// stride a multiple of 4 floats): at 1024 sequences the scalar kernel took 25 us per step.
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__global__ void gru_gate4_kernel(const float* __restrict__ gi, const float* __restrict__ gh, int gh_parts,
                                 int64_t gh_part_stride, const float* __restrict__ b_hh, const float* __restrict__ hprev, int64_t ldh,
                                 float* __restrict__ y, int64_t ldy, const float* __restrict__ resid, int64_t ldres,
                                 float* __restrict__ out, int64_t ldout, float* __restrict__ hn, int S, int T,
                                 int H, int t) {
    const int u = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int s = blockIdx.y;
    if (u >= H) return;
    const int64_t f = (int64_t)s * T + t;
    const float* g = gi + f * 3 * H;
    float4 hr, hz, hnn;
    if (gh) {
        const float* q = gh + (int64_t)s * 3 * H;
        hr = ld4(q + u); hz = ld4(q + H + u); hnn = ld4(q + 2 * H + u);
        for (int part = 1; part < gh_parts; ++part) {
            q += gh_part_stride;
            const float4 a = ld4(q + u), b = ld4(q + H + u), c = ld4(q + 2 * H + u);
            hr.x += a.x; hr.y += a.y; hr.z += a.z; hr.w += a.w;
            hz.x += b.x; hz.y += b.y; hz.z += b.z; hz.w += b.w;
            hnn.x += c.x; hnn.y += c.y; hnn.z += c.z; hnn.w += c.w;
        }
    } else {
        hr = ld4(b_hh + u); hz = ld4(b_hh + H + u); hnn = ld4(b_hh + 2 * H + u);
    }
    const float4 hp = hprev ? ld4(hprev + (int64_t)s * ldh + u) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 gr = ld4(g + u), gz = ld4(g + H + u), gn = ld4(g + 2 * H + u);
    float4 h;
#define GAIT_GATE1(c)                                                  \
    {                                                                  \
        const float r = sigmoidf_(gr.c + hr.c);                        \
        const float z = sigmoidf_(gz.c + hz.c);                        \
        const float n = tanhf(gn.c + r * hnn.c);                       \
        h.c = (1.f - z) * n + z * hp.c;                                \
    }
    GAIT_GATE1(x) GAIT_GATE1(y) GAIT_GATE1(z) GAIT_GATE1(w)
#undef GAIT_GATE1
    *reinterpret_cast<float4*>(y + f * ldy + u) = h;
    if (out) {
        const float4 rs = ld4(resid + f * ldres + u);
        *reinterpret_cast<float4*>(out + f * ldout + u) = make_float4(h.x + rs.x, h.y + rs.y, h.z + rs.z, h.w + rs.w);
    }
    if (hn) *reinterpret_cast<float4*>(hn + (int64_t)s * H + u) = h;
}

__global__ void relu_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) y[i] = fmaxf(x[i], 0.f);
}

}  // namespace gait

using namespace gait;

extern "C" {

int gait_relu(const float* x, float* y, int64_t n, gait_stream_t stream) {
    GAIT_REQUIRE(n >= 0 && (n == 0 || (x && y)), "relu: null pointer or negative n");
    if (n == 0) return GAIT_OK;
    relu_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(x, y, n);
    return check_launch("relu");
}

size_t gait_gru_workspace_bytes(int64_t S, int64_t T, int64_t H) {
    if (S <= 0 || T <= 0 || H <= 0) return 0;
    return (size_t)(S * T * 3 * H + kGruMaxSplits * S * 3 * H) * sizeof(float) + 128 * (size_t)(H / 16 + 2)
           + (size_t)(2 * S * H) * sizeof(float);   // + per-CTA step flags (one 128-byte line each) and the h_lo scratch of the persistent kernel
}

int gait_gru_plan(int64_t S, int64_t T, int64_t H) {
    if (S <= 0 || T <= 0 || H <= 0) return 0;
    const char* e = getenv("GAITB200_GRU_PATH");
    if ((e && atoi(e) == 1) || linear_path() == 1) return 0;
    const float* a = reinterpret_cast<const float*>(uintptr_t(256));        // stands for any 16-byte aligned operand
    const char* sm = getenv("GAITB200_GRU_SMALL");
    if (!(sm && atoi(sm) == 0) && T >= 4 && gru_small_eligible(a, a, nullptr, a, H, S, T, H)) return 2;
    const char* mc = getenv("GAITB200_GRU_MAXCHUNKED");
    const int64_t max_chunked = mc ? std::max<int64_t>(0, atoll(mc)) : (H >= 1024 ? 64 : 320);
    if (S <= max_chunked && gru_recurrent_eligible(a, a, nullptr, a, H, a, H, a, H, std::min<int64_t>(S, 64), T, H)) return 1;
    return 0;
}

int gait_gru_layer(const float* x, int64_t ldx, const float* W_ih, const float* W_hh, const float* b_ih,
                   const float* b_hh, const float* h0, float* y, int64_t ldy, const float* resid,
                   int64_t ldres, float* out, int64_t ldout, float* hn, int64_t S, int64_t T, int64_t I,
                   int64_t H, int reverse, void* workspace, size_t workspace_bytes, gait_stream_t stream) {
    GAIT_REQUIRE(S >= 0 && T >= 0 && I > 0 && H > 0, "gru_layer: bad sizes");
    if (S == 0 || T == 0) return GAIT_OK;
    GAIT_REQUIRE(x && W_ih && W_hh && b_ih && b_hh && y && workspace, "gru_layer: null pointer");
    GAIT_REQUIRE(ldx >= I && ldy >= H, "gru_layer: stride smaller than row");
    GAIT_REQUIRE(out == nullptr || (resid != nullptr && ldout >= H && ldres >= H), "gru_layer: out needs resid and strides");
    GAIT_REQUIRE(S < 65536, "gru_layer: at most 65535 sequences per call");
    if (workspace_bytes < gait_gru_workspace_bytes(S, T, H)) {
        set_error("gru_layer: workspace %zu < %zu bytes", workspace_bytes, gait_gru_workspace_bytes(S, T, H));
        return GAIT_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    float* gi = static_cast<float*>(workspace);
    float* gh = gi + S * T * 3 * H;
    const int64_t F = S * T;
    // gi = x . W_ih^T + b_ih for every frame
    {
        // the projection feeds the gates' sigmoid / tanh and a contractive recurrence: 64-deep partial sums are accurate
        // enough here (encoder output rms error 8.8e-8 against FP64, FP32 torch 5.8e-8; 4.8e-8 with 32-deep sums) and the
        // GEMM is 10 % faster than with 32-deep ones
        LinearPromote promote(2);
        GAIT_TRY(linear_launch(x, ldx, W_ih, I, b_ih, nullptr, 0, gi, 3 * H, F, 3 * H, I, st));
    }
    // whole recurrence in one persistent cluster kernel (gru_rec.cu) when the shape allows;
    // GAITB200_GRU_PATH=1 forces the per-step path below, =2 makes ineligibility an error
    static int gru_path = -1;
    if (gru_path < 0) {
        const char* e = getenv("GAITB200_GRU_PATH");
        gru_path = e ? atoi(e) : 0;
    }
    // one or two sequences (a single long clip, BASELINE configs[3]): weight-stationary kernel, W_hh stays in registers and
    // shared memory for all T steps (gru_small.cu); GAITB200_GRU_SMALL=0 disables it (A/B)
    static int gru_small = -1;
    if (gru_small < 0) {
        const char* e = getenv("GAITB200_GRU_SMALL");
        gru_small = (e && atoi(e) == 0) ? 0 : 1;
    }
    if (gru_path != 1 && gru_small && gru_small_eligible(gi, W_hh, h0, y, ldy, S, T, H) && T >= 4) {
        // the per-step path's partial-sum area (4 x S x 3H floats, unused here) holds the (value, tag) exchange buffer: 2 x S x H pairs
        const int rc = gru_small_launch(gi, W_hh, b_hh, h0, y, ldy, resid, ldres, out, ldout, hn, S, T, reverse, gh, st);
        if (rc != GAIT_GRU_RETRY_PER_STEP) return rc;
    }
    if (gru_path != 1 && linear_path() != 1) {
        // consecutive 64-sequence launches of the persistent kernel (kept for small layers and as an A/B switch; at H = 2048 the
        // per-step GEMMs, whose efficiency grows with the number of rows, win as soon as there is more than one launch)
        constexpr int64_t kChunk = 64;
        // crossover measured on B200 (scripts/gru_s_sweep.py, profiles/r02zf_gru_s_sweep.md): GRU stage M frames/s at
        // S = 64/128/192/256/320/512: 64-sequence chunks 2.27/2.37/2.46/2.45/2.49/2.44, per-step 2.08/2.71/2.50/3.03/3.50/3.23
        // (per-step path with W_hh prepared, the vectorised gate kernel and the tile-width cost model of linear_tc.cu; before
        // those it won only from 512 sequences on); 72 / 96 / 112 sequences: chunks 1.53 / 1.95 / 2.17, per-step 1.95 / 2.40 / 2.64.
        // So one launch of the persistent kernel for up to 64 sequences, the per-step path beyond.
        static int64_t kMaxChunked = -2;              // GAITB200_GRU_MAXCHUNKED: A/B switch for the crossover (-1: not set)
        if (kMaxChunked == -2) {
            const char* e = getenv("GAITB200_GRU_MAXCHUNKED");
            kMaxChunked = e ? std::max<int64_t>(0, atoll(e)) : -1;
        }
        // not set = by H: the crossover was measured at H = 2048 only (above); small layers, whose per-step GEMMs are
        // launch-bound, keep the multi-launch persistent path up to 320 sequences
        const int64_t max_chunked = kMaxChunked >= 0 ? kMaxChunked : (H >= 1024 ? 64 : 320);
        const int64_t S0 = std::min<int64_t>(S, kChunk);
        if (S <= max_chunked && gru_recurrent_eligible(gi, W_hh, h0, y, ldy, resid, ldres, out, ldout, S0, T, H)) {
            unsigned int* counter = reinterpret_cast<unsigned int*>(gh + kGruMaxSplits * S * 3 * H);
            uintptr_t lo_addr = (reinterpret_cast<uintptr_t>(counter) + 128 * (size_t)(H / 16 + 1) + 127) & ~(uintptr_t)127;
            bool refused = false;
            for (int64_t s0 = 0; s0 < S && !refused; s0 += kChunk) {
                const int64_t Sc = std::min<int64_t>(kChunk, S - s0);
                const int rc = gru_recurrent_launch(gi + s0 * T * 3 * H, W_hh, b_hh, h0 ? h0 + s0 * H : nullptr, y + s0 * T * ldy, ldy,
                                                    resid ? resid + s0 * T * ldres : nullptr, ldres, out ? out + s0 * T * ldout : nullptr,
                                                    ldout, hn ? hn + s0 * H : nullptr, Sc, T, H, reverse, counter,
                                                    reinterpret_cast<float*>(lo_addr), st);
                if (rc == GAIT_GRU_RETRY_PER_STEP) refused = true;      // no co-residency guarantee: per-step path below
                else if (rc != GAIT_OK) return rc;
            }
            if (!refused) return GAIT_OK;
            if (gru_path == 2) return GAIT_ERR_CUDA;                    // set_error was filled by the launcher
        }
        else if (gru_path == 2) {
            set_error("gru_layer: persistent recurrent kernel not eligible for S=%lld T=%lld H=%lld", (long long)S, (long long)T, (long long)H);
            return GAIT_ERR_UNSUPPORTED;
        }
    }
    const dim3 block(256), grid((unsigned)ceil_div(H, 256), (unsigned)S);
    // recurrent GEMM (S,H).(H,3H): tensor-core path with split-K so that ~all SMs get a tile
    const int64_t hstride = T * ldy;
    const bool tc = linear_path() != 1 && (T > 1) && linear_tc_eligible(y, hstride, W_hh, H, S, 3 * H, H);
    int splits = 1;
    if (tc) {
        const int64_t tiles = ceil_div(3 * H, 128) * ceil_div(S, S <= 64 ? 64 : 128);
        const int64_t nkb = ceil_div(H, 32);
        splits = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(kGruMaxSplits, 148 / std::max<int64_t>(tiles, 1)), nkb));
        while (splits > 1 && ceil_div(nkb, ceil_div(nkb, splits)) != splits) --splits;
    }
    for (int64_t step = 0; step < T; ++step) {
        const int64_t t = reverse ? (T - 1 - step) : step;
        const float* hprev = nullptr;
        int64_t ldh = 0;
        if (step == 0) {
            hprev = h0; ldh = H;
        } else {
            const int64_t tp = reverse ? t + 1 : t - 1;
            hprev = y + tp * ldy; ldh = T * ldy;        // row s of h_{t-1} is y[s, tp, :]
        }
        const float* ghp = nullptr;
        int parts = 1;
        if (hprev) {
            if (tc && linear_tc_eligible(hprev, ldh, W_hh, H, S, 3 * H, H)) {
                GAIT_TRY(linear_tc_launch(hprev, ldh, W_hh, H, b_hh, nullptr, 0, gh, 3 * H, S, 3 * H, H, splits, S * 3 * H, st));
                parts = splits;
            } else {
                GAIT_TRY(linear_launch(hprev, ldh, W_hh, H, b_hh, nullptr, 0, gh, 3 * H, S, 3 * H, H, st));
            }
            ghp = gh;
        }
        const bool vec4 = (H % 4 == 0) && aligned16(gi) && aligned16(gh) && aligned16(b_hh) && aligned16(y) && (ldy % 4 == 0) &&
                          (!hprev || (aligned16(hprev) && ldh % 4 == 0)) &&
                          (!out || (aligned16(out) && aligned16(resid) && ldout % 4 == 0 && ldres % 4 == 0)) && (!hn || aligned16(hn));
        if (vec4) {
            const dim3 grid4((unsigned)ceil_div(H / 4, 256), (unsigned)S);
            gru_gate4_kernel<<<grid4, block, 0, st>>>(gi, ghp, parts, S * 3 * H, b_hh, hprev, ldh, y, ldy, resid, ldres, out, ldout,
                                                      (step == T - 1) ? hn : nullptr, (int)S, (int)T, (int)H, (int)t);
        } else
        gru_gate_kernel<<<grid, block, 0, st>>>(gi, ghp, parts, S * 3 * H, b_hh, hprev, ldh, y, ldy, resid, ldres, out, ldout,
                                                (step == T - 1) ? hn : nullptr, (int)S, (int)T, (int)H, (int)t);
        GAIT_TRY(check_launch("gru_gate"));
    }
    return GAIT_OK;
}

}  // extern "C"
