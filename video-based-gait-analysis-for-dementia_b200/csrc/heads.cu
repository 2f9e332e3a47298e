// Small heads next to the regression path:
//   gait_keypoint_attention   lib/models/layers/keypoint_attention.py:34-55 (softmax over the pixels of each joint's
//                             heat map, then attention pooling of the feature map)
//   gait_locally_connected    lib/models/layers/locallyconnected2d.py:39-49 for kernel_size = 1, output_size = [J, 1]:
//                             J unshared per-joint linear maps (PARE pose MLP pare.py:422-430; cparam MLP
//                             gait_feat_encoder.py:43-49), also used with a shared weight for the 1x1 smpl_final_layer
//   gait_activation           LeakyReLU(slope) / Tanh between the Linear layers of gait_feat_encoder.py:58-78
#include "common.cuh"

namespace gait {

// grid (J, B), 256 threads.  heat (B,J,HW), feat (B,C,HW) -> out[b*sob + c*soc + j*soj].
__global__ void keypoint_attention_kernel(const float* __restrict__ feat, const float* __restrict__ heat, float scale,
                                          float* __restrict__ out, int C, int J, int HW, int64_t sob, int64_t soc, int64_t soj) {
    extern __shared__ float w[];                      // HW softmax weights
    __shared__ float red[32];
    const int j = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const float* h = heat + ((int64_t)b * J + j) * HW;
    float m = -INFINITY;
    for (int i = tid; i < HW; i += blockDim.x) { const float v = h[i] * scale; w[i] = v; m = fmaxf(m, v); }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int i = 1; i < nw; ++i) m = fmaxf(m, red[i]);
    __syncthreads();
    float s = 0.f;
    for (int i = tid; i < HW; i += blockDim.x) { const float e = expf(w[i] - m); w[i] = e; s += e; }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    s = 0.f;
    for (int i = 0; i < nw; ++i) s += red[i];
    const float inv = 1.f / s;
    // attention pooling: one warp per channel, coalesced row reads (rows are re-read by the J CTAs of a frame from L2)
    for (int c = warp; c < C; c += nw) {
        const float* f = feat + ((int64_t)b * C + c) * HW;
        float a = 0.f;
        for (int i = lane; i < HW; i += 32) a = fmaf(w[i] * inv, f[i], a);
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) out[(int64_t)b * sob + (int64_t)c * soc + (int64_t)j * soj] = a;
    }
}

// out[n,o,j] = sum_c x[n,c,j] W[o,c,j] (+ bias[o,j]) (+ resid -> out2); all operands through element strides so that
// broadcast inputs (stride 0), shared weights (swj = 0) and transposed outputs need no copies.
__global__ void locally_connected_kernel(const float* __restrict__ x, int64_t sxn, int64_t sxc, int64_t sxj,
                                         const float* __restrict__ W, int64_t swo, int64_t swc, int64_t swj,
                                         const float* __restrict__ bias, int64_t sbo, int64_t sbj,
                                         float* __restrict__ out, int64_t son, int64_t soo, int64_t soj,
                                         const float* __restrict__ resid, float* __restrict__ out2, int64_t N, int C, int O, int J) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= N * O * J) return;
    const int j = (int)(idx % J), o = (int)((idx / J) % O);
    const int64_t n = idx / ((int64_t)J * O);
    const float* xp = x + n * sxn + j * sxj;
    const float* wp = W + o * swo + j * swj;
    float a = 0.f;
    for (int c = 0; c < C; ++c) a = fmaf(xp[c * sxc], wp[c * swc], a);
    if (bias) a += bias[o * sbo + j * sbj];
    const int64_t oi = n * son + o * soo + j * soj;
    out[oi] = a;
    if (out2) out2[oi] = a + resid[oi];
}

__global__ void activation_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, int kind, float slope) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    y[i] = kind == 0 ? (v > 0.f ? v : v * slope) : tanhf(v);
}

}  // namespace gait

using namespace gait;

extern "C" {

int gait_keypoint_attention(const float* feat, const float* heat, float scale, float* out, int64_t B, int C, int J, int HW,
                            int64_t sob, int64_t soc, int64_t soj, gait_stream_t stream) {
    GAIT_REQUIRE(B >= 0 && C > 0 && J > 0 && HW > 0, "keypoint_attention: bad sizes");
    if (B == 0) return GAIT_OK;
    GAIT_REQUIRE(feat && heat && out, "keypoint_attention: null pointer");
    GAIT_REQUIRE(B < 65536 && HW <= 48 * 1024 / 4, "keypoint_attention: at most 65535 frames and 12288 pixels per map");
    keypoint_attention_kernel<<<dim3((unsigned)J, (unsigned)B), 256, (size_t)HW * sizeof(float), as_stream(stream)>>>(
        feat, heat, scale, out, C, J, HW, sob, soc, soj);
    return check_launch("keypoint_attention");
}

int gait_locally_connected(const float* x, int64_t sxn, int64_t sxc, int64_t sxj, const float* W, int64_t swo, int64_t swc,
                           int64_t swj, const float* bias, int64_t sbo, int64_t sbj, float* out, int64_t son, int64_t soo,
                           int64_t soj, const float* resid, float* out2, int64_t N, int C, int O, int J, gait_stream_t stream) {
    GAIT_REQUIRE(N >= 0 && C > 0 && O > 0 && J > 0, "locally_connected: bad sizes");
    if (N == 0) return GAIT_OK;
    GAIT_REQUIRE(x && W && out && ((resid == nullptr) == (out2 == nullptr)), "locally_connected: null pointer (resid and out2 go together)");
    locally_connected_kernel<<<(unsigned)ceil_div(N * O * J, 256), 256, 0, as_stream(stream)>>>(
        x, sxn, sxc, sxj, W, swo, swc, swj, bias, sbo, sbj, out, son, soo, soj, resid, out2, N, C, O, J);
    return check_launch("locally_connected");
}

int gait_activation(const float* x, float* y, int64_t n, int kind, float slope, gait_stream_t stream) {
    GAIT_REQUIRE(n >= 0 && (kind == 0 || kind == 1), "activation: kind 0 (leaky_relu) or 1 (tanh)");
    if (n == 0) return GAIT_OK;
    GAIT_REQUIRE(x && y, "activation: null pointer");
    activation_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(x, y, n, kind, slope);
    return check_launch("activation");
}

}  // extern "C"
