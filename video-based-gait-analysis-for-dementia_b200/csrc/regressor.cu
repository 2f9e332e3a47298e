// gait_hmr_regressor: the iterative HMR/VIBE regressor loop of lib/models/spin.py:244-265.
//
//   for it in range(n_iter):
//       xc = cat([x, pose, shape, cam]); h = fc2(fc1(xc)); state += dec(h)
//
// fc1 is split column-wise: the x part (Din of the 2205 columns) does not change between
// iterations, so x.W1x^T + b1 is computed once; each iteration then needs only the 157-column
// state part.  There is no non-linearity between the layers (dropout is the identity in eval),
// exactly as in the reference.  The three decoders are one (157,Dh) matrix.
#include <algorithm>

#include "common.cuh"

namespace gait {

constexpr int kState = 157;
constexpr int kStateLd = 160;

__global__ void broadcast_state_kernel(const float* __restrict__ init, int64_t init_rows, float* __restrict__ state,
                                       int64_t F) {
    pdl_wait_cta();
    pdl_trigger();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= F * kStateLd) return;
    const int64_t f = i / kStateLd;
    const int c = (int)(i % kStateLd);
    const float v = (c < kState) ? init[(init_rows == 1 ? 0 : f) * kStateLd + c] : 0.f;
    state[i] = v;
}

// Shared initial state (init_rows == 1): init.W1s^T is the same vector v for every frame, so it is computed once per call
// (one warp per hidden unit) and carried by the bias: the x-part GEMM gets b1 + v and directly yields fc1 of iteration 0;
// later iterations add state.W1s^T - v to it.  bias1 = [b1 + v | -v].
// The same launch also broadcasts the initial state (blocks below `state_blocks`).
__global__ void init_term_kernel(const float* __restrict__ W1s, const float* __restrict__ init, const float* __restrict__ b1,
                                 float* __restrict__ bias1, int64_t Dh, float* __restrict__ state, int64_t F,
                                 unsigned state_blocks) {
    pdl_wait_cta();
    pdl_trigger();
    if (blockIdx.x < state_blocks) {
        const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
        if (i < F * kStateLd) {
            const int c = (int)(i % kStateLd);
            state[i] = (c < kState) ? init[c] : 0.f;
        }
        return;
    }
    const int64_t row = ((blockIdx.x - state_blocks) * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= Dh) return;
    float v = 0.f;
    for (int c = lane; c < kState; c += 32) v = fmaf(W1s[row * kStateLd + c], init[c], v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) {
        bias1[row] = b1[row] + v;
        bias1[Dh + row] = -v;
    }
}

constexpr int kDecSplits = 6;          // split-K of the 157-row decoder GEMM (only 24 output tiles otherwise)

// state[f, c] += sum over split-K partials (partial 0 already holds bias); one thread per element
__global__ void decoder_reduce_kernel(const float* __restrict__ part, int parts, int64_t part_stride,
                                      float* __restrict__ state, int64_t F) {
    pdl_wait_cta();
    pdl_trigger();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= F * kStateLd) return;
    const int c = (int)(i % kStateLd);
    if (c >= kState) return;
    float acc = 0.f;
    for (int p = 0; p < parts; ++p) acc += part[p * part_stride + i];
    state[i] += acc;
}

// state[f, c] = sum over split-K partials (partial 0 holds the bias); padding columns are zeroed
__global__ void folded_reduce_kernel(const float* __restrict__ part, int parts, int64_t part_stride,
                                     float* __restrict__ state, int64_t F) {
    pdl_wait_cta();
    pdl_trigger();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= F * kStateLd) return;
    const int c = (int)(i % kStateLd);
    float acc = 0.f;
    if (c < kState)
        for (int p = 0; p < parts; ++p) acc += part[p * part_stride + i];
    state[i] = acc;
}

static int folded_splits(const float* x, int64_t ldx, const float* Wf, int64_t F, int64_t Din) {
    if (linear_path() == 1 || !linear_tc_eligible(x, ldx, Wf, Din, F, kState, Din)) return 1;
    const int64_t nkb = ceil_div(Din, 32);
    int splits = (int)std::min<int64_t>(kDecSplits, nkb);
    while (splits > 1 && ceil_div(nkb, ceil_div(nkb, splits)) != splits) --splits;
    return splits;
}

}  // namespace gait

using namespace gait;

extern "C" {

size_t gait_hmr_workspace_bytes(int64_t F, int64_t Dh) {
    if (F <= 0 || Dh <= 0) return 0;
    return (size_t)(3 * F * Dh + kDecSplits * F * kStateLd + 2 * Dh) * sizeof(float);
}

size_t gait_hmr_folded_workspace_bytes(int64_t F) {
    if (F <= 0) return 0;
    return (size_t)(kDecSplits * F * kStateLd) * sizeof(float);
}

int gait_hmr_regressor_folded(const float* x, int64_t ldx, const float* Wf, const float* bf, float* state_out,
                              int64_t F, int64_t Din, void* workspace, size_t workspace_bytes, gait_stream_t stream) {
    GAIT_REQUIRE(F >= 0 && Din > 0, "hmr_regressor_folded: bad sizes");
    if (F == 0) return GAIT_OK;
    GAIT_REQUIRE(x && Wf && bf && state_out && workspace, "hmr_regressor_folded: null pointer");
    GAIT_REQUIRE(ldx >= Din, "hmr_regressor_folded: ldx < Din");
    if (workspace_bytes < gait_hmr_folded_workspace_bytes(F)) {
        set_error("hmr_regressor_folded: workspace %zu < %zu bytes", workspace_bytes, gait_hmr_folded_workspace_bytes(F));
        return GAIT_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    float* part = static_cast<float*>(workspace);
    // F x 157 x Din: few output tiles, so K is cut across CTAs when the tensor-core path takes it
    const int splits = folded_splits(x, ldx, Wf, F, Din);
    if (splits > 1) {
        GAIT_TRY(linear_tc_launch(x, ldx, Wf, Din, bf, nullptr, 0, part, kStateLd, F, kState, Din, splits, F * kStateLd, st));
    } else {
        GAIT_TRY(linear_launch(x, ldx, Wf, Din, bf, nullptr, 0, part, kStateLd, F, kState, Din, st));
    }
    launch_pdl(4, folded_reduce_kernel, dim3((unsigned)ceil_div(F * kStateLd, 256)), dim3(256), 0, st, part, splits, F * kStateLd, state_out, F);
    return check_launch("hmr folded_reduce");
}

int gait_hmr_regressor(const float* x, int64_t ldx, const float* W1x, const float* W1s, const float* b1,
                       const float* W2, const float* b2, const float* Wd, const float* bd,
                       const float* init, int64_t init_rows, int n_iter, float* state_out, int64_t F,
                       int64_t Din, int64_t Dh, void* workspace, size_t workspace_bytes,
                       gait_stream_t stream) {
    GAIT_REQUIRE(F >= 0 && Din > 0 && Dh > 0 && n_iter >= 0, "hmr_regressor: bad sizes");
    if (F == 0) return GAIT_OK;
    GAIT_REQUIRE(x && W1x && W1s && b1 && W2 && b2 && Wd && bd && init && state_out && workspace,
                 "hmr_regressor: null pointer");
    GAIT_REQUIRE(init_rows == 1 || init_rows == F, "hmr_regressor: init must have 1 or F rows");
    GAIT_REQUIRE(ldx >= Din, "hmr_regressor: ldx < Din");
    if (workspace_bytes < gait_hmr_workspace_bytes(F, Dh)) {
        set_error("hmr_regressor: workspace %zu < %zu bytes", workspace_bytes, gait_hmr_workspace_bytes(F, Dh));
        return GAIT_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    float* hx = static_cast<float*>(workspace);
    float* h1 = hx + F * Dh;
    float* h2 = h1 + F * Dh;
    float* dpart = h2 + F * Dh;                          // split-K partials of the decoder GEMM
    float* bias1 = dpart + kDecSplits * F * kStateLd;    // [b1 + init.W1s^T | -init.W1s^T] when the init is shared
    const bool shared_init = init_rows == 1;
    // decoder GEMM (F x 157 x Dh): few output tiles, so cut K across CTAs when the tensor-core path takes it
    int dsplits = 1;
    if (linear_path() != 1 && linear_tc_eligible(h2, Dh, Wd, Dh, F, kState, Dh)) {
        const int64_t nkb = ceil_div(Dh, 32);
        dsplits = (int)std::min<int64_t>(kDecSplits, nkb);
        while (dsplits > 1 && ceil_div(nkb, ceil_div(nkb, dsplits)) != dsplits) --dsplits;
    }
    const unsigned state_blocks = (unsigned)ceil_div(F * kStateLd, 256);
    if (shared_init && n_iter > 0) {
        launch_pdl(4, init_term_kernel, dim3(state_blocks + (unsigned)ceil_div(Dh * 32, 256)), dim3(256), 0, st, W1s, init, b1, bias1, Dh,
                   state_out, F, state_blocks);
        GAIT_TRY(check_launch("hmr init_term + broadcast_state"));
    } else {
        launch_pdl(4, broadcast_state_kernel, dim3(state_blocks), dim3(256), 0, st, init, init_rows, state_out, F);
        GAIT_TRY(check_launch("hmr broadcast_state"));
    }
    if (n_iter == 0) return GAIT_OK;
    // iteration-invariant part of fc1 (with a shared init: all of fc1 for iteration 0)
    GAIT_TRY(linear_launch(x, ldx, W1x, Din, shared_init ? bias1 : b1, nullptr, 0, hx, Dh, F, Dh, Din, st));
    for (int it = 0; it < n_iter; ++it) {
        const float* fc1 = h1;
        if (shared_init && it == 0) fc1 = hx;
        else GAIT_TRY(linear_launch(state_out, kStateLd, W1s, kStateLd, shared_init ? bias1 + Dh : nullptr, hx, Dh, h1, Dh, F, Dh,
                                    kStateLd, st));
        GAIT_TRY(linear_launch(fc1, Dh, W2, Dh, b2, nullptr, 0, h2, Dh, F, Dh, Dh, st));
        if (dsplits > 1) {
            GAIT_TRY(linear_tc_launch(h2, Dh, Wd, Dh, bd, nullptr, 0, dpart, kStateLd, F, kState, Dh, dsplits,
                                      F * kStateLd, st));
            launch_pdl(4, decoder_reduce_kernel, dim3((unsigned)ceil_div(F * kStateLd, 256)), dim3(256), 0, st, dpart, dsplits,
                       F * kStateLd, state_out, F);
            GAIT_TRY(check_launch("hmr decoder_reduce"));
        } else {
            GAIT_TRY(linear_launch(h2, Dh, Wd, Dh, bd, state_out, kStateLd, state_out, kStateLd, F, kState, Dh, st));
        }
    }
    return GAIT_OK;
}

}  // extern "C"
