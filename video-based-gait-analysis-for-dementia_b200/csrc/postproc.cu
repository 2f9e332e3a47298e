// Post-processing next to the regression head:
//   gait_one_euro_filter            lib/utils/one_euro_filter.py:5-46 as lib/utils/smooth_pose.py:51-56,84-88 drives it
//   gait_crop_cam_to_orig_img       lib/utils/demo_utils.py:176-193
//   gait_crop_coords_to_orig_img    lib/utils/demo_utils.py:196-209
// The reference does this in numpy; every operation below is the correctly rounded single operation numpy performs
// (no FMA contraction, numpy's scalar/array type promotion), so results are bit-identical.
#include <math.h>

#include "common.cuh"

namespace gait {

// One thread per channel, T sequential steps (the filter is a recurrence over frames); channels are contiguous, so
// each step is a coalesced row access.  t_e = 1 at every step (frames are sampled at t = 0, 1, 2, ...).
__global__ void one_euro_kernel(const float* __restrict__ x, float* __restrict__ xh, int64_t T, int64_t C, float mc, float beta,
                                float rd, float two_pi) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= C) return;
    float x_prev = x[c], dx_prev = 0.f;
    xh[c] = x_prev;
    const float a_d = __fdiv_rn(rd, __fadd_rn(rd, 1.f));                    // smoothing_factor(t_e, d_cutoff)
    for (int64_t t = 1; t < T; ++t) {
        const float xi = x[t * C + c];
        const float dx = __fsub_rn(xi, x_prev);                              // (x - x_prev) / t_e, t_e = 1
        const float dx_hat = __fadd_rn(__fmul_rn(a_d, dx), __fmul_rn(__fsub_rn(1.f, a_d), dx_prev));
        const float cutoff = __fadd_rn(mc, __fmul_rn(beta, fabsf(dx_hat)));
        const float r = __fmul_rn(two_pi, cutoff);
        const float a = __fdiv_rn(r, __fadd_rn(r, 1.f));
        const float x_hat = __fadd_rn(__fmul_rn(a, xi), __fmul_rn(__fsub_rn(1.f, a), x_prev));
        xh[t * C + c] = x_hat;
        x_prev = x_hat;
        dx_prev = dx_hat;
    }
}

template <typename T> struct Arith;
template <> struct Arith<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Arith<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

// P = dtype of bbox = dtype of the result (numpy promotes float32 cam against float64 boxes to float64)
template <typename P>
__global__ void crop_cam_kernel(const float* __restrict__ cam, const P* __restrict__ bbox, int64_t ldb, double img_w, double img_h,
                                P* __restrict__ out, int64_t N) {
    using A = Arith<P>;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const P cx = bbox[i * ldb], cy = bbox[i * ldb + 1], h = bbox[i * ldb + 2];
    const P hw = (P)(img_w / 2.), hh = (P)(img_h / 2.);
    const P s = (P)cam[i * 3], c1 = (P)cam[i * 3 + 1], c2 = (P)cam[i * 3 + 2];
    const P sx = A::mul(s, A::div((P)1, A::div((P)img_w, h)));
    const P sy = A::mul(s, A::div((P)1, A::div((P)img_h, h)));
    const P tx = A::add(A::div(A::div(A::sub(cx, hw), hw), sx), c1);
    const P ty = A::add(A::div(A::div(A::sub(cy, hh), hh), sy), c2);
    out[i * 4] = sx; out[i * 4 + 1] = sy; out[i * 4 + 2] = tx; out[i * 4 + 3] = ty;
}

// keypoints stay float32 (the reference updates the float32 array in place; mixed operations run in P and are cast back)
template <typename P>
__global__ void crop_coords_kernel(const P* __restrict__ bbox, int64_t ldb, const float* __restrict__ kp, float* __restrict__ out,
                                   int64_t N, int J, int D, double crop) {
    using A = Arith<P>;
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= N * J * D) return;
    const int64_t n = idx / ((int64_t)J * D);
    const int d = (int)(idx % D);
    const P cx = bbox[n * ldb], cy = bbox[n * ldb + 1], h = bbox[n * ldb + 2];
    float k = __fmul_rn((float)(0.5 * crop), __fadd_rn(kp[idx], 1.0f));         // 0.5 * crop_size * (keypoints + 1.0)
    k = (float)A::mul((P)k, A::div(h, (P)crop));                                 // keypoints *= h / crop_size
    if (d == 0) k = (float)A::add(A::sub(cx, A::div(h, (P)2)), (P)k);            // + (cx - h/2)
    else if (d == 1) k = (float)A::add(A::sub(cy, A::div(h, (P)2)), (P)k);       // + (cy - h/2)
    out[idx] = k;
}

}  // namespace gait

using namespace gait;

extern "C" {

int gait_one_euro_filter(const float* x, float* x_hat, int64_t T, int64_t C, double min_cutoff, double beta, double d_cutoff,
                         gait_stream_t stream) {
    GAIT_REQUIRE(T >= 0 && C >= 0, "one_euro_filter: negative size");
    if (T == 0 || C == 0) return GAIT_OK;
    GAIT_REQUIRE(x && x_hat, "one_euro_filter: null pointer");
    // numpy turns the Python-float products into float32 when they meet the float32 arrays
    const float rd = (float)(2 * M_PI * d_cutoff), two_pi = (float)(2 * M_PI);
    one_euro_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, as_stream(stream)>>>(x, x_hat, T, C, (float)min_cutoff, (float)beta, rd, two_pi);
    return check_launch("one_euro_filter");
}

int gait_crop_cam_to_orig_img(const float* cam, const void* bbox, int bbox_is_f64, int64_t ldb, double img_width,
                              double img_height, void* out, int64_t N, gait_stream_t stream) {
    GAIT_REQUIRE(N >= 0, "crop_cam_to_orig_img: negative size");
    if (N == 0) return GAIT_OK;
    GAIT_REQUIRE(cam && bbox && out && ldb >= 3, "crop_cam_to_orig_img: null pointer or bbox rows shorter than 3");
    const unsigned grid = (unsigned)ceil_div(N, 128);
    if (bbox_is_f64)
        crop_cam_kernel<double><<<grid, 128, 0, as_stream(stream)>>>(cam, static_cast<const double*>(bbox), ldb, img_width, img_height,
                                                                       static_cast<double*>(out), N);
    else
        crop_cam_kernel<float><<<grid, 128, 0, as_stream(stream)>>>(cam, static_cast<const float*>(bbox), ldb, img_width, img_height,
                                                                      static_cast<float*>(out), N);
    return check_launch("crop_cam_to_orig_img");
}

int gait_crop_coords_to_orig_img(const void* bbox, int bbox_is_f64, int64_t ldb, const float* keypoints, float* out, int64_t N,
                                 int J, int D, double crop_size, gait_stream_t stream) {
    GAIT_REQUIRE(N >= 0 && J >= 0 && D >= 2, "crop_coords_to_orig_img: bad sizes (need at least x and y)");
    if (N == 0 || J == 0) return GAIT_OK;
    GAIT_REQUIRE(bbox && keypoints && out && ldb >= 3, "crop_coords_to_orig_img: null pointer or bbox rows shorter than 3");
    const unsigned grid = (unsigned)ceil_div(N * J * D, 256);
    if (bbox_is_f64)
        crop_coords_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const double*>(bbox), ldb, keypoints, out, N, J, D, crop_size);
    else
        crop_coords_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(bbox), ldb, keypoints, out, N, J, D, crop_size);
    return check_launch("crop_coords_to_orig_img");
}

}  // extern "C"
