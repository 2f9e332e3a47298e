// Device functions for the rotation/camera arithmetic of lib/utils/geometry.py.
// Written for FP32 parity with the PyTorch reference: no fast-math intrinsics, divisions and
// square roots are IEEE (nvcc default), operations follow the reference's order.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace gait {

// F.normalize(v, eps): v / max(||v||_2, eps)
__device__ __forceinline__ void normalize3(float& x, float& y, float& z, float eps) {
    float n = sqrtf(x * x + y * y + z * z);
    n = fmaxf(n, eps);
    x = x / n;
    y = y / n;
    z = z / n;
}

// geometry.py:395-410.  in: 6 floats viewed (3,2): a1 = (x0,x2,x4), a2 = (x1,x3,x5).
// out: R[r*3+c], columns b1,b2,b3.
__device__ __forceinline__ void rot6d_to_rotmat_dev(const float* __restrict__ x, float eps, float* R) {
    float a1x = x[0], a1y = x[2], a1z = x[4];
    float a2x = x[1], a2y = x[3], a2z = x[5];
    normalize3(a1x, a1y, a1z, eps);
    float d = a1x * a2x + a1y * a2y + a1z * a2z;
    float ux = a2x - d * a1x, uy = a2y - d * a1y, uz = a2z - d * a1z;
    normalize3(ux, uy, uz, eps);
    float cx = a1y * uz - a1z * uy;
    float cy = a1z * ux - a1x * uz;
    float cz = a1x * uy - a1y * ux;
    R[0] = a1x; R[1] = ux; R[2] = cx;
    R[3] = a1y; R[4] = uy; R[5] = cy;
    R[6] = a1z; R[7] = uz; R[8] = cz;
}

// geometry.py:213-293 on R given as R[i*rs + j]; quaternion (w,x,y,z).
__device__ __forceinline__ void rotmat_to_quat_dev(const float* __restrict__ R, int rs, float eps, float* q) {
    // the reference works on the transposed matrix: m[i][j] = R[j][i]
    const float m00 = R[0], m01 = R[rs], m02 = R[2 * rs];
    const float m10 = R[1], m11 = R[rs + 1], m12 = R[2 * rs + 1];
    const float m20 = R[2], m21 = R[rs + 2], m22 = R[2 * rs + 2];
    const bool d2 = m22 < eps;
    const bool d0_gt_d1 = m00 > m11;
    const bool d0_lt_nd1 = m00 < -m11;
    float t, q0, q1, q2, q3;
    if (d2 && d0_gt_d1) {
        t = 1.f + m00 - m11 - m22;
        q0 = m12 - m21; q1 = t; q2 = m01 + m10; q3 = m20 + m02;
    } else if (d2) {
        t = 1.f - m00 + m11 - m22;
        q0 = m20 - m02; q1 = m01 + m10; q2 = t; q3 = m12 + m21;
    } else if (d0_lt_nd1) {
        t = 1.f - m00 - m11 + m22;
        q0 = m01 - m10; q1 = m20 + m02; q2 = m12 + m21; q3 = t;
    } else {
        t = 1.f + m00 + m11 + m22;
        q0 = t; q1 = m12 - m21; q2 = m20 - m02; q3 = m01 - m10;
    }
    const float s = sqrtf(t);
    q[0] = (q0 / s) * 0.5f;
    q[1] = (q1 / s) * 0.5f;
    q[2] = (q2 / s) * 0.5f;
    q[3] = (q3 / s) * 0.5f;
}

// geometry.py:159-210
__device__ __forceinline__ void quat_to_axis_angle_dev(const float* q, float* aa) {
    const float q1 = q[1], q2 = q[2], q3 = q[3];
    const float s2 = q1 * q1 + q2 * q2 + q3 * q3;
    const float s = sqrtf(s2);
    const float c = q[0];
    const float two_theta = 2.0f * ((c < 0.0f) ? atan2f(-s, -c) : atan2f(s, c));
    const float k = (s2 > 0.0f) ? (two_theta / s) : 2.0f;
    aa[0] = q1 * k;
    aa[1] = q2 * k;
    aa[2] = q3 * k;
}

// geometry.py:68-97 (NaN -> 0)
__device__ __forceinline__ void rotmat_to_axis_angle_dev(const float* __restrict__ R, int rs, float* aa) {
    float q[4];
    rotmat_to_quat_dev(R, rs, 1e-6f, q);
    quat_to_axis_angle_dev(q, aa);
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (isnan(aa[i])) aa[i] = 0.f;
}

// geometry.py:38-65
__device__ __forceinline__ void quat2mat_dev(const float* q, float* R) {
    const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const float w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
    const float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
    const float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
    R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;     R[2] = 2 * wy + 2 * xz;
    R[3] = 2 * wz + 2 * xy;   R[4] = w2 - x2 + y2 - z2;   R[5] = 2 * yz - 2 * wx;
    R[6] = 2 * xz - 2 * wy;   R[7] = 2 * wx + 2 * yz;     R[8] = w2 - x2 - y2 + z2;
}

// smplx lbs.batch_rodrigues: angle = ||a + 1e-8||, R = I + sin K + (1-cos) K K
__device__ __forceinline__ void rodrigues_smplx_dev(const float* a, float* R) {
    const float bx = a[0] + 1e-8f, by = a[1] + 1e-8f, bz = a[2] + 1e-8f;
    const float angle = sqrtf(bx * bx + by * by + bz * bz);
    const float rx = a[0] / angle, ry = a[1] / angle, rz = a[2] / angle;
    const float c = cosf(angle), s = sinf(angle);
    const float K[9] = {0.f, -rz, ry, rz, 0.f, -rx, -ry, rx, 0.f};
    const float omc = 1.f - c;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float kk = K[i * 3 + 0] * K[0 * 3 + k] + K[i * 3 + 1] * K[1 * 3 + k] + K[i * 3 + 2] * K[2 * 3 + k];
            R[i * 3 + k] = ((i == k) ? 1.f : 0.f) + s * K[i * 3 + k] + omc * kk;
        }
}

// geometry.py:23-35
__device__ __forceinline__ void rodrigues_quat_dev(const float* a, float* R) {
    const float bx = a[0] + 1e-8f, by = a[1] + 1e-8f, bz = a[2] + 1e-8f;
    const float angle = sqrtf(bx * bx + by * by + bz * bz);
    const float ux = a[0] / angle, uy = a[1] / angle, uz = a[2] / angle;
    const float half = angle * 0.5f;
    const float c = cosf(half), s = sinf(half);
    const float q[4] = {c, s * ux, s * uy, s * uz};
    quat2mat_dev(q, R);
}

// geometry.py:427-446 third component
__device__ __forceinline__ float weak_persp_tz(float s, float focal, float res) {
    return 2.f * focal / (res * s + 1e-9f);
}

// geometry.py:448-479 with identity rotation and zero camera centre: K [(X+t)/z]
__device__ __forceinline__ void project_point(float X, float Y, float Z, float tx, float ty, float tz,
                                              float focal, float cx, float cy, float divisor, float* o) {
    float px = X + tx, py = Y + ty, pz = Z + tz;
    const float zn = pz / pz;      // 1, or NaN exactly where the reference produces NaN
    px = px / pz;
    py = py / pz;
    o[0] = (focal * px + cx * zn) / divisor;
    o[1] = (focal * py + cy * zn) / divisor;
}

}  // namespace gait
