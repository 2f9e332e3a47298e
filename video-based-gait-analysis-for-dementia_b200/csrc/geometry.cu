// Elementwise rotation / camera kernels: one thread per rotation (or per point).
// Replaces the ~8 (rot6d) and ~40 (R -> axis-angle) ATen launches of lib/utils/geometry.py.
#include "common.cuh"
#include "geometry.cuh"

namespace gait {

constexpr int kThreads = 256;

__global__ void rot6d_to_rotmat_kernel(const float* __restrict__ x6, int group, int64_t in_group_stride,
                                       float* __restrict__ R, int64_t n, float eps) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float in[6], out[9];
    // 24-byte records on an even stride: 8-byte aligned
    const float2* p = reinterpret_cast<const float2*>(x6 + (i / group) * in_group_stride + (i % group) * 6);
    float2 a = p[0], b = p[1], c = p[2];
    in[0] = a.x; in[1] = a.y; in[2] = b.x; in[3] = b.y; in[4] = c.x; in[5] = c.y;
    rot6d_to_rotmat_dev(in, eps, out);
#pragma unroll
    for (int k = 0; k < 9; ++k) R[i * 9 + k] = out[k];
}

__global__ void rotmat_to_rot6d_kernel(const float* __restrict__ R, float* __restrict__ x6, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = R + i * 9;
    float* o = x6 + i * 6;
    o[0] = r[0]; o[1] = r[1]; o[2] = r[3]; o[3] = r[4]; o[4] = r[6]; o[5] = r[7];
}

__global__ void rotmat_to_quat_kernel(const float* __restrict__ R, int rs, float* __restrict__ q, int64_t n, float eps) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m[12], out[4];
    const int len = 3 * rs;
    for (int k = 0; k < len; ++k) m[k] = R[i * len + k];
    rotmat_to_quat_dev(m, rs, eps, out);
    reinterpret_cast<float4*>(q)[i] = make_float4(out[0], out[1], out[2], out[3]);
}

__global__ void quat_to_axis_angle_kernel(const float* __restrict__ q, float* __restrict__ aa, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = reinterpret_cast<const float4*>(q)[i];
    float in[4] = {v.x, v.y, v.z, v.w}, out[3];
    quat_to_axis_angle_dev(in, out);
    aa[i * 3 + 0] = out[0]; aa[i * 3 + 1] = out[1]; aa[i * 3 + 2] = out[2];
}

__global__ void rotmat_to_axis_angle_kernel(const float* __restrict__ R, int rs, float* __restrict__ aa, int64_t n,
                                            int group, int64_t out_group_stride, int out_offset) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m[12], out[3];
    const int len = 3 * rs;
    for (int k = 0; k < len; ++k) m[k] = R[i * len + k];
    rotmat_to_axis_angle_dev(m, rs, out);
    float* o = aa + (i / group) * out_group_stride + out_offset + (i % group) * 3;
    o[0] = out[0]; o[1] = out[1]; o[2] = out[2];
}

__global__ void quat2mat_kernel(const float* __restrict__ q, float* __restrict__ R, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = reinterpret_cast<const float4*>(q)[i];
    float in[4] = {v.x, v.y, v.z, v.w}, out[9];
    quat2mat_dev(in, out);
#pragma unroll
    for (int k = 0; k < 9; ++k) R[i * 9 + k] = out[k];
}

__global__ void rodrigues_kernel(const float* __restrict__ aa, float* __restrict__ R, int64_t n, int variant) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float in[3] = {aa[i * 3], aa[i * 3 + 1], aa[i * 3 + 2]}, out[9];
    if (variant == 0) rodrigues_smplx_dev(in, out);
    else rodrigues_quat_dev(in, out);
#pragma unroll
    for (int k = 0; k < 9; ++k) R[i * 9 + k] = out[k];
}

__global__ void weak_persp_kernel(const float* __restrict__ cam, float* __restrict__ t, int64_t n, float focal, float res) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    t[i * 3 + 0] = cam[i * 3 + 1];
    t[i * 3 + 1] = cam[i * 3 + 2];
    t[i * 3 + 2] = weak_persp_tz(cam[i * 3], focal, res);
}

__global__ void perspective_projection_kernel(const float* __restrict__ pts, const float* __restrict__ rot,
                                              const float* __restrict__ trans, const float* __restrict__ center,
                                              float focal, float divisor, float* __restrict__ out, int64_t b, int j) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= b * j) return;
    const int64_t bi = i / j;
    float X = pts[i * 3], Y = pts[i * 3 + 1], Z = pts[i * 3 + 2];
    if (rot != nullptr) {
        const float* r = rot + bi * 9;
        const float x2 = r[0] * X + r[1] * Y + r[2] * Z;
        const float y2 = r[3] * X + r[4] * Y + r[5] * Z;
        const float z2 = r[6] * X + r[7] * Y + r[8] * Z;
        X = x2; Y = y2; Z = z2;
    }
    const float cx = center ? center[bi * 2] : 0.f, cy = center ? center[bi * 2 + 1] : 0.f;
    float o[2];
    project_point(X, Y, Z, trans[bi * 3], trans[bi * 3 + 1], trans[bi * 3 + 2], focal, cx, cy, divisor, o);
    reinterpret_cast<float2*>(out)[i] = make_float2(o[0], o[1]);
}

static inline unsigned grid_for(int64_t n) { return (unsigned)ceil_div(n, kThreads); }

}  // namespace gait

using namespace gait;

extern "C" {

int gait_rot6d_to_rotmat(const float* x6, int group, int64_t in_group_stride, float* R, int64_t n, float eps,
                         gait_stream_t stream) {
    GAIT_REQUIRE(n >= 0 && (n == 0 || (x6 && R)), "rot6d_to_rotmat: null pointer or negative n");
    if (n == 0) return GAIT_OK;
    GAIT_REQUIRE(group >= 1 && in_group_stride >= 6 * (int64_t)group && (in_group_stride & 1) == 0,
                 "rot6d_to_rotmat: group stride must be even and >= 6*group");
    GAIT_REQUIRE(aligned8(x6), "rot6d_to_rotmat: x6 must be 8-byte aligned");
    rot6d_to_rotmat_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(x6, group, in_group_stride, R, n, eps);
    return check_launch("rot6d_to_rotmat");
}

int gait_rotmat_to_rot6d(const float* R, float* x6, int64_t n, gait_stream_t stream) {
    GAIT_REQUIRE(n >= 0 && (n == 0 || (x6 && R)), "rotmat_to_rot6d: null pointer or negative n");
    if (n == 0) return GAIT_OK;
    rotmat_to_rot6d_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(R, x6, n);
    return check_launch("rotmat_to_rot6d");
}

int gait_rotmat_to_quaternion(const float* R, int row_stride, float* quat, int64_t n, float eps,
                              gait_stream_t stream) {
    GAIT_REQUIRE(n >= 0 && (n == 0 || (quat && R)), "rotmat_to_quaternion: null pointer or negative n");
    GAIT_REQUIRE(row_stride == 3 || row_stride == 4, "rotmat_to_quaternion: row_stride must be 3 or 4");
    if (n == 0) return GAIT_OK;
    GAIT_REQUIRE(aligned16(quat), "rotmat_to_quaternion: quat must be 16-byte aligned");
    rotmat_to_quat_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(R, row_stride, quat, n, eps);
    return check_launch("rotmat_to_quaternion");
}

int gait_quaternion_to_axis_angle(const float* quat, float* aa, int64_t n, gait_stream_t stream) {
    GAIT_REQUIRE(n >= 0 && (n == 0 || (quat && aa)), "quaternion_to_axis_angle: null pointer or negative n");
    if (n == 0) return GAIT_OK;
    GAIT_REQUIRE(aligned16(quat), "quaternion_to_axis_angle: quat must be 16-byte aligned");
    quat_to_axis_angle_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(quat, aa, n);
    return check_launch("quaternion_to_axis_angle");
}

int gait_rotmat_to_axis_angle(const float* R, int row_stride, float* aa, int64_t n, int group,
                              int64_t out_group_stride, int out_offset, gait_stream_t stream) {
    GAIT_REQUIRE(n >= 0 && (n == 0 || (R && aa)), "rotmat_to_axis_angle: null pointer or negative n");
    GAIT_REQUIRE(row_stride == 3 || row_stride == 4, "rotmat_to_axis_angle: row_stride must be 3 or 4");
    GAIT_REQUIRE(group >= 1 && out_offset >= 0 && out_group_stride >= 3 * (int64_t)group + out_offset,
                 "rotmat_to_axis_angle: bad output packing");
    if (n == 0) return GAIT_OK;
    rotmat_to_axis_angle_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(R, row_stride, aa, n, group,
                                                                                 out_group_stride, out_offset);
    return check_launch("rotmat_to_axis_angle");
}

int gait_quat2mat(const float* quat, float* R, int64_t n, gait_stream_t stream) {
    GAIT_REQUIRE(n >= 0 && (n == 0 || (quat && R)), "quat2mat: null pointer or negative n");
    if (n == 0) return GAIT_OK;
    GAIT_REQUIRE(aligned16(quat), "quat2mat: quat must be 16-byte aligned");
    quat2mat_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(quat, R, n);
    return check_launch("quat2mat");
}

int gait_batch_rodrigues(const float* aa, float* R, int64_t n, int variant, gait_stream_t stream) {
    GAIT_REQUIRE(n >= 0 && (n == 0 || (aa && R)), "batch_rodrigues: null pointer or negative n");
    GAIT_REQUIRE(variant == 0 || variant == 1, "batch_rodrigues: variant must be 0 (smplx) or 1 (geometry.py)");
    if (n == 0) return GAIT_OK;
    rodrigues_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(aa, R, n, variant);
    return check_launch("batch_rodrigues");
}

int gait_weak_perspective_to_translation(const float* cam, float* trans, int64_t n, float focal_length,
                                         float img_res, gait_stream_t stream) {
    GAIT_REQUIRE(n >= 0 && (n == 0 || (cam && trans)), "weak_perspective_to_translation: null pointer or negative n");
    if (n == 0) return GAIT_OK;
    weak_persp_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(cam, trans, n, focal_length, img_res);
    return check_launch("weak_perspective_to_translation");
}

int gait_perspective_projection(const float* points, const float* rotation, const float* translation,
                                const float* center, float focal_length, float out_divisor, float* out,
                                int64_t b, int j, gait_stream_t stream) {
    GAIT_REQUIRE(b >= 0 && j >= 0, "perspective_projection: negative size");
    if (b == 0 || j == 0) return GAIT_OK;
    GAIT_REQUIRE(points && translation && out, "perspective_projection: null pointer");
    GAIT_REQUIRE(aligned8(out), "perspective_projection: out must be 8-byte aligned");
    perspective_projection_kernel<<<grid_for(b * j), kThreads, 0, as_stream(stream)>>>(
        points, rotation, translation, center, focal_length, out_divisor, out, b, j);
    return check_launch("perspective_projection");
}

}  // extern "C"
