// SMPL stage kernels: rest joints + kinematic chain, blend shapes, linear blend skinning,
// joint regression, joint assembly / projection, theta packing.
// Arithmetic restated from smplx==0.1.26 lbs.py (the reference calls it at
// lib/models/smpl.py:111,113) and lib/models/smpl.py:108-130,149-191.
#include "common.cuh"
#include "geometry.cuh"

namespace gait {

constexpr int NJ = GAIT_NUM_JOINTS;      // 24
constexpr int NB = GAIT_NUM_BETAS;       // 10

// ------------------------------------------------------------------------------------------
// Kinematic chain: one warp per frame, lane j = joint j.  Each lane builds its local transform
// [R_j | J_j - J_parent], then the world transforms are composed level by level down the tree:
// at level d every lane of depth d pulls its parent's 3x4 transform with __shfl_sync and
// multiplies.  The SMPL tree has 9 levels, so the 23 serial 4x4 matmul launches of smplx
// batch_rigid_transform become 8 shuffle rounds inside one warp.
// ------------------------------------------------------------------------------------------
constexpr int kChainWarps = 4;

// Optional fusions (the regression head runs them as ONE launch instead of rot6d + chain + theta):
//   x6 != NULL : the rotations come from the 6-D representation (geometry.py:395-410), row f at x6 + f * ldx6, joint j at
//                + 6 j, and are also written to R_out (F,24,3,3);
//   theta != NULL : theta (F,85) = [cam | axis-angle(72) | betas] (spin.py:288, pare.py:79) with the geometry.py:68-97 route.
__global__ void __launch_bounds__(kChainWarps * 32)
smpl_pose_chain_kernel(const float* __restrict__ R, const float* __restrict__ x6, int64_t ldx6, float eps6,
                       float* __restrict__ R_out, const float* __restrict__ betas, int64_t ldb,
                       const float* __restrict__ cam, int64_t ldcam, float* __restrict__ theta,
                       const float* __restrict__ J_template, const float* __restrict__ J_shapedirs,
                       const int32_t* __restrict__ parents, float* __restrict__ A, float* __restrict__ J_posed,
                       float* __restrict__ coef, float* __restrict__ Aop, int64_t F) {
    __shared__ int s_parent[32];
    __shared__ int s_depth[32];
    pdl_wait_cta();
    pdl_trigger();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 32) s_parent[threadIdx.x] = (threadIdx.x < NJ) ? parents[threadIdx.x] : -1;
    __syncthreads();
    if (threadIdx.x < 32) {
        int d = 0, p = (threadIdx.x < NJ) ? s_parent[threadIdx.x] : -1;
        while (p >= 0 && d < NJ) { ++d; p = s_parent[p]; }
        s_depth[threadIdx.x] = (threadIdx.x < NJ) ? d : -1;
    }
    __syncthreads();
    const int64_t f = blockIdx.x * (int64_t)kChainWarps + warp;
    if (f >= F) return;
    const bool active = lane < NJ;
    const int j = active ? lane : 0;
    const int parent = active ? s_parent[j] : -1;
    const int depth = s_depth[lane];
    const int max_depth = __reduce_max_sync(0xffffffffu, depth);

    // rotation of this joint
    float r[9];
    if (x6) {
        const float2* p = reinterpret_cast<const float2*>(x6 + f * ldx6 + j * 6);      // 24-byte records, 8-byte aligned
        const float2 a0 = p[0], a1 = p[1], a2 = p[2];
        const float in[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y};
        rot6d_to_rotmat_dev(in, eps6, r);
        if (active) {
#pragma unroll
            for (int k = 0; k < 9; ++k) R_out[(f * NJ + j) * 9 + k] = r[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) r[k] = R[(f * NJ + j) * 9 + k];
    }
    if (theta) {
        // lane k < 24 converts its rotation, lanes 24..26 copy the camera, lanes 27..31 two betas each
        float* th = theta + f * 85;
        if (lane < NJ) {
            float aa[3];
            rotmat_to_axis_angle_dev(r, 3, aa);
            th[3 + lane * 3] = aa[0]; th[3 + lane * 3 + 1] = aa[1]; th[3 + lane * 3 + 2] = aa[2];
        } else if (lane < NJ + 3) {
            th[lane - NJ] = cam[f * ldcam + (lane - NJ)];
        } else {
            const int q = (lane - NJ - 3) * 2;
            th[75 + q] = betas[f * ldb + q];
            th[75 + q + 1] = betas[f * ldb + q + 1];
        }
    }

    // rest joint: J = J_template + J_shapedirs . beta
    float b[NB];
#pragma unroll
    for (int l = 0; l < NB; ++l) b[l] = betas[f * ldb + l];
    float Jr[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float acc = J_template[j * 3 + c];
#pragma unroll
        for (int l = 0; l < NB; ++l) acc = fmaf(J_shapedirs[(j * 3 + c) * NB + l], b[l], acc);
        Jr[c] = acc;
    }
    // relative joint position
    const int src = parent >= 0 ? parent : 0;
    float rel[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float pj = __shfl_sync(0xffffffffu, Jr[c], src);
        rel[c] = (parent >= 0) ? (Jr[c] - pj) : Jr[c];
    }
    // world transform G = [g (3x3) | t (3)], starts as the local transform
    float g[9], t[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) g[k] = r[k];
#pragma unroll
    for (int c = 0; c < 3; ++c) t[c] = rel[c];

    for (int level = 1; level <= max_depth; ++level) {
        float pg[9], pt[3];
#pragma unroll
        for (int k = 0; k < 9; ++k) pg[k] = __shfl_sync(0xffffffffu, g[k], src);
#pragma unroll
        for (int c = 0; c < 3; ++c) pt[c] = __shfl_sync(0xffffffffu, t[c], src);
        if (depth == level) {
            float ng[9], nt[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    ng[i * 3 + k] = pg[i * 3 + 0] * r[0 * 3 + k] + pg[i * 3 + 1] * r[1 * 3 + k] + pg[i * 3 + 2] * r[2 * 3 + k];
                nt[i] = pg[i * 3 + 0] * rel[0] + pg[i * 3 + 1] * rel[1] + pg[i * 3 + 2] * rel[2] + pt[i];
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) g[k] = ng[k];
#pragma unroll
            for (int c = 0; c < 3; ++c) t[c] = nt[c];
        }
    }
    if (!active) return;
    // A = G - pad(G @ [J;0]): translation column loses the rotated rest joint
    float av[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float gj = g[i * 3 + 0] * Jr[0] + g[i * 3 + 1] * Jr[1] + g[i * 3 + 2] * Jr[2];
        av[i * 4 + 0] = g[i * 3 + 0]; av[i * 4 + 1] = g[i * 3 + 1]; av[i * 4 + 2] = g[i * 3 + 2]; av[i * 4 + 3] = t[i] - gj;
    }
    if (A) {
        float* a = A + (f * NJ + j) * 12;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            reinterpret_cast<float4*>(a)[i] = make_float4(av[i * 4], av[i * 4 + 1], av[i * 4 + 2], av[i * 4 + 3]);
    }
    if (Aop) {
        // tensor-core LBS operand (lbs_tc.cu): per 8-frame group a blob [hi|lo][kchunk][rowgroup][8][4] with
        // row n = (f % 8) * 12 + c, k = joint; values pre-split into TF32 hi/lo.
        float* blob = Aop + (f >> 3) * (2 * 6 * 96 * 4);
        const int fl = (int)(f & 7);
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            const int n = fl * 12 + c;
            const int idx = ((j >> 2) * 12 + (n >> 3)) * 32 + (n & 7) * 4 + (j & 3);
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(av[c]));
            const float hi = __uint_as_float(hb);
            uint32_t lb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(av[c] - hi));
            blob[idx] = hi;
            blob[6 * 96 * 4 + idx] = __uint_as_float(lb);
        }
    }
    float* jp = J_posed + (f * NJ + j) * 3;
    jp[0] = t[0]; jp[1] = t[1]; jp[2] = t[2];
    if (coef) {
        float* cf = coef + f * GAIT_BLEND_LD;
        if (j >= 1) {
#pragma unroll
            for (int k = 0; k < 9; ++k) cf[(j - 1) * 9 + k] = r[k] - ((k == 0 || k == 4 || k == 8) ? 1.f : 0.f);
        } else {
#pragma unroll
            for (int l = 0; l < NB; ++l) cf[GAIT_POSE_BASIS + l] = b[l];
            cf[GAIT_POSE_BASIS + NB] = 1.f;
#pragma unroll
            for (int k = GAIT_BLEND_K; k < GAIT_BLEND_LD; ++k) cf[k] = 0.f;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Linear blend skinning, SIMT FP32.  CTA = 128 vertices x FT frames.
//   T[f,v] = sum_j W[v,j] A[f,j]   (3x4, dense W exactly as smplx does)
//   verts[f,v] = T[f,v] . [v_posed[f,v]; 1]
// The W tile is read once per CTA (padded rows in smem, conflict-free), A for the frame tile
// sits in smem and is read as broadcast float4, v_posed / verts move through smem so every
// global access is a coalesced 8-byte stream.
// ------------------------------------------------------------------------------------------
constexpr int LBS_VT = 128;     // vertices per CTA (= threads)
constexpr int LBS_FT = 8;       // frames per CTA

__global__ void __launch_bounds__(LBS_VT)
smpl_lbs_kernel(const float* __restrict__ v_posed, int64_t ldv, const float* __restrict__ A, const float* __restrict__ W,
                float* __restrict__ verts, int F, int V) {
    __shared__ __align__(16) float sA[LBS_FT][NJ * 12];
    __shared__ float sW[LBS_VT][NJ + 1];
    __shared__ __align__(16) float sV[LBS_FT][LBS_VT * 3];
    const int tid = threadIdx.x;
    const int v0 = blockIdx.x * LBS_VT;
    const int f0 = blockIdx.y * LBS_FT;
    const int nv = min(LBS_VT, V - v0);
    const int nf = min(LBS_FT, F - f0);

    // stage A (nf x 288 floats, contiguous in global)
    {
        const float4* src = reinterpret_cast<const float4*>(A + (int64_t)f0 * NJ * 12);
        float4* dst = reinterpret_cast<float4*>(&sA[0][0]);
        for (int i = tid; i < nf * NJ * 3; i += LBS_VT) dst[i] = src[i];
    }
    // stage W tile (nv x 24 floats, contiguous)
    for (int i = tid; i < nv * NJ; i += LBS_VT) sW[i / NJ][i % NJ] = W[(int64_t)v0 * NJ + i];
    // stage v_posed tile: per frame nv*3 contiguous floats (8-byte aligned: V*3*4 and v0*12 are multiples of 8)
    for (int f = 0; f < nf; ++f) {
        const float2* src = reinterpret_cast<const float2*>(v_posed + (int64_t)(f0 + f) * ldv + (int64_t)v0 * 3);
        float2* dst = reinterpret_cast<float2*>(&sV[f][0]);
        const int n2 = (nv * 3) >> 1;
        for (int i = tid; i < n2; i += LBS_VT) dst[i] = src[i];
        if (((nv * 3) & 1) && tid == 0) sV[f][nv * 3 - 1] = v_posed[(int64_t)(f0 + f) * ldv + (int64_t)v0 * 3 + nv * 3 - 1];
    }
    __syncthreads();

    float w[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) w[j] = (tid < nv) ? sW[tid][j] : 0.f;

    for (int f = 0; f < nf; f += 2) {
        const int f1 = (f + 1 < nf) ? f + 1 : f;
        float t0[12], t1[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) { t0[k] = 0.f; t1[k] = 0.f; }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float4* a0 = reinterpret_cast<const float4*>(&sA[f][j * 12]);
            const float4* a1 = reinterpret_cast<const float4*>(&sA[f1][j * 12]);
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float4 x = a0[q], y = a1[q];
                t0[q * 4 + 0] = fmaf(w[j], x.x, t0[q * 4 + 0]); t0[q * 4 + 1] = fmaf(w[j], x.y, t0[q * 4 + 1]);
                t0[q * 4 + 2] = fmaf(w[j], x.z, t0[q * 4 + 2]); t0[q * 4 + 3] = fmaf(w[j], x.w, t0[q * 4 + 3]);
                t1[q * 4 + 0] = fmaf(w[j], y.x, t1[q * 4 + 0]); t1[q * 4 + 1] = fmaf(w[j], y.y, t1[q * 4 + 1]);
                t1[q * 4 + 2] = fmaf(w[j], y.z, t1[q * 4 + 2]); t1[q * 4 + 3] = fmaf(w[j], y.w, t1[q * 4 + 3]);
            }
        }
        if (tid < nv) {
            {
                const float x = sV[f][tid * 3], y = sV[f][tid * 3 + 1], z = sV[f][tid * 3 + 2];
                const float ox = t0[0] * x + t0[1] * y + t0[2] * z + t0[3];
                const float oy = t0[4] * x + t0[5] * y + t0[6] * z + t0[7];
                const float oz = t0[8] * x + t0[9] * y + t0[10] * z + t0[11];
                sV[f][tid * 3] = ox; sV[f][tid * 3 + 1] = oy; sV[f][tid * 3 + 2] = oz;
            }
            if (f1 != f) {
                const float x = sV[f1][tid * 3], y = sV[f1][tid * 3 + 1], z = sV[f1][tid * 3 + 2];
                const float ox = t1[0] * x + t1[1] * y + t1[2] * z + t1[3];
                const float oy = t1[4] * x + t1[5] * y + t1[6] * z + t1[7];
                const float oz = t1[8] * x + t1[9] * y + t1[10] * z + t1[11];
                sV[f1][tid * 3] = ox; sV[f1][tid * 3 + 1] = oy; sV[f1][tid * 3 + 2] = oz;
            }
        }
    }
    __syncthreads();
    for (int f = 0; f < nf; ++f) {
        float2* dst = reinterpret_cast<float2*>(verts + ((int64_t)(f0 + f) * V + v0) * 3);
        const float2* src = reinterpret_cast<const float2*>(&sV[f][0]);
        const int n2 = (nv * 3) >> 1;
        for (int i = tid; i < n2; i += LBS_VT) dst[i] = src[i];
        if (((nv * 3) & 1) && tid == 0) verts[((int64_t)(f0 + f) * V + v0) * 3 + nv * 3 - 1] = sV[f][nv * 3 - 1];
    }
}

// ------------------------------------------------------------------------------------------
// vertices2joints: out[f,r,:] = sum_v Jreg[r,v] verts[f,v,:].  CTA = JR_FT frames x up to JR_RT
// regressor rows; threads stride over vertices (coalesced), block-reduce at the end.
// ------------------------------------------------------------------------------------------
constexpr int JR_THREADS = 256;
constexpr int JR_FT = 2;
constexpr int JR_RT = 9;

__global__ void __launch_bounds__(JR_THREADS)
joint_regress_kernel(const float* __restrict__ verts, const float* __restrict__ Jreg, float* __restrict__ out,
                     int F, int V, int Rj) {
    const int f0 = blockIdx.x * JR_FT;
    const int r0 = blockIdx.y * JR_RT;
    const int nr = min(JR_RT, Rj - r0);
    const int nf = min(JR_FT, F - f0);
    float acc[JR_FT][JR_RT][3];
#pragma unroll
    for (int f = 0; f < JR_FT; ++f)
#pragma unroll
        for (int r = 0; r < JR_RT; ++r) { acc[f][r][0] = 0.f; acc[f][r][1] = 0.f; acc[f][r][2] = 0.f; }
    for (int v = threadIdx.x; v < V; v += JR_THREADS) {
        float p[JR_FT][3];
#pragma unroll
        for (int f = 0; f < JR_FT; ++f) {
            const int ff = (f < nf) ? f0 + f : f0;
            const float* q = verts + ((int64_t)ff * V + v) * 3;
            p[f][0] = q[0]; p[f][1] = q[1]; p[f][2] = q[2];
        }
#pragma unroll
        for (int r = 0; r < JR_RT; ++r) {
            if (r < nr) {
                const float w = Jreg[(int64_t)(r0 + r) * V + v];
#pragma unroll
                for (int f = 0; f < JR_FT; ++f) {
                    acc[f][r][0] = fmaf(w, p[f][0], acc[f][r][0]);
                    acc[f][r][1] = fmaf(w, p[f][1], acc[f][r][1]);
                    acc[f][r][2] = fmaf(w, p[f][2], acc[f][r][2]);
                }
            }
        }
    }
    __shared__ float red[JR_THREADS / 32][JR_FT * JR_RT * 3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int f = 0; f < JR_FT; ++f)
#pragma unroll
        for (int r = 0; r < JR_RT; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v = acc[f][r][c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) red[warp][(f * JR_RT + r) * 3 + c] = v;
            }
    __syncthreads();
    if (threadIdx.x < JR_FT * JR_RT * 3) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < JR_THREADS / 32; ++w) s += red[w][threadIdx.x];
        const int f = threadIdx.x / (JR_RT * 3), r = (threadIdx.x / 3) % JR_RT, c = threadIdx.x % 3;
        if (f < nf && r < nr) out[((int64_t)(f0 + f) * Rj + r0 + r) * 3 + c] = s;
    }
}

// ------------------------------------------------------------------------------------------
// Joint assembly + projection + Kinect-25 gather.  A block handles JA_FB frames: first the extra joints' partial sums
// (one per vertex tile when the regressor row was fused into skinning: 54 strided values per coordinate) are reduced by
// groups of 8 lanes into shared memory, then one thread per (frame, output joint) gathers / projects.
// ------------------------------------------------------------------------------------------
constexpr int JA_FB = 8;             // frames per block
constexpr int JA_MAX_EXTRA = 9;      // J_regressor_extra rows
struct ExtraJoints {
    const float* data;       // (parts, F, n_extra, 3): partial sums over vertex tiles, or one complete part
    int n_extra, parts;
    int64_t part_stride;
};

__device__ __forceinline__ void fetch_virtual_joint(int v, int64_t f, int fl, const float* __restrict__ J_posed,
                                                    const float* __restrict__ verts, int64_t V,
                                                    const int32_t* __restrict__ landmarks, int n_landmarks,
                                                    const float (*ex_s)[JA_MAX_EXTRA * 3], float* o) {
    if (v >= NJ + n_landmarks) {
        const float* p = ex_s[fl] + (v - NJ - n_landmarks) * 3;
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
        return;
    }
    const float* p = (v < NJ) ? J_posed + (f * NJ + v) * 3 : verts + (f * V + landmarks[v - NJ]) * 3;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
}

__global__ void __launch_bounds__(256)
joints_assemble_kernel(const float* __restrict__ J_posed, const float* __restrict__ verts, int64_t V,
                       const int32_t* __restrict__ landmarks, int n_landmarks, ExtraJoints ex,
                       const int32_t* __restrict__ joint_map, int J, float* __restrict__ joints,
                       const float* __restrict__ cam, int64_t ldcam, float focal, float res,
                       float divisor, float* __restrict__ kp2d, const int32_t* __restrict__ gather,
                       int n_gather, float* __restrict__ gathered, int64_t F) {
    __shared__ float ex_s[JA_FB][JA_MAX_EXTRA * 3];
    pdl_wait_cta();
    pdl_trigger();
    const int64_t f0 = (int64_t)blockIdx.x * JA_FB;
    const int nf = (int)min((int64_t)JA_FB, F - f0);
    // phase 1: sum the partial extra joints; 8 consecutive lanes share one (frame, joint, coordinate)
    const int n_sums = nf * ex.n_extra * 3;
    const int sub = threadIdx.x & 7;
    for (int base = 0; base < n_sums; base += 32) {                      // 256 threads = 32 groups of 8 lanes
        const int q = base + (threadIdx.x >> 3);
        float a = 0.f;
        if (q < n_sums) {
            const int fl = q / (ex.n_extra * 3), r = q % (ex.n_extra * 3);
            const float* p = ex.data + ((f0 + fl) * ex.n_extra) * 3 + r;
            for (int part = sub; part < ex.parts; part += 8) a += p[part * ex.part_stride];
        }
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        if (q < n_sums && sub == 0) ex_s[q / (ex.n_extra * 3)][q % (ex.n_extra * 3)] = a;
    }
    __syncthreads();
    // phase 2: one thread per (frame, output slot)
    const int per = J + n_gather;
    for (int i = threadIdx.x; i < nf * per; i += blockDim.x) {
        const int fl = i / per, k = i % per;
        const int64_t f = f0 + fl;
        float p[3];
        if (k < J) {
            fetch_virtual_joint(joint_map[k], f, fl, J_posed, verts, V, landmarks, n_landmarks, ex_s, p);
            float* o = joints + (f * J + k) * 3;
            o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
            if (kp2d) {
                const float* c = cam + f * ldcam;
                float q[2];
                project_point(p[0], p[1], p[2], c[1], c[2], weak_persp_tz(c[0], focal, res), focal, 0.f, 0.f, divisor, q);
                reinterpret_cast<float2*>(kp2d)[f * J + k] = make_float2(q[0], q[1]);
            }
        } else {
            const int g = gather[k - J];
            if (g >= 0) fetch_virtual_joint(joint_map[g], f, fl, J_posed, verts, V, landmarks, n_landmarks, ex_s, p);
            else { p[0] = 0.f; p[1] = 0.f; p[2] = 0.f; }
            float* o = gathered + (f * n_gather + (k - J)) * 3;
            o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
        }
    }
}

// theta = [cam | axis-angle(72) | betas]: thread (f, k) with k<24 converts one rotation, k in 24..26 copy.
__global__ void pack_theta_kernel(const float* __restrict__ R, const float* __restrict__ cam, int64_t ldcam,
                                  const float* __restrict__ betas, int64_t ldb, float* __restrict__ theta, int64_t F) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= F * 32) return;
    const int64_t f = i >> 5;
    const int k = (int)(i & 31);
    float* th = theta + f * 85;
    if (k < NJ) {
        float m[9], aa[3];
#pragma unroll
        for (int q = 0; q < 9; ++q) m[q] = R[(f * NJ + k) * 9 + q];
        rotmat_to_axis_angle_dev(m, 3, aa);
        th[3 + k * 3] = aa[0]; th[3 + k * 3 + 1] = aa[1]; th[3 + k * 3 + 2] = aa[2];
    } else if (k < NJ + 3) {
        th[k - NJ] = cam[f * ldcam + (k - NJ)];
    } else if (k < NJ + 3 + 5) {
        const int q = (k - NJ - 3) * 2;
        th[75 + q] = betas[f * ldb + q];
        th[75 + q + 1] = betas[f * ldb + q + 1];
    }
}

// ------------------------------------------------------------------------------------------
// Joints-only skinning (BASELINE configs[4]) without the mesh.  The Kinect-25 set needs only the n_lm landmark vertices
// and ONE regressor row (thorax) over all vertices.  Skinning is linear in v_posed for fixed transforms,
//     thorax[f] = sum_v jx[v] T[f,v] [v_posed[f,v]; 1] = sum_j A[f,j] [ P_j coef[f] ; s_j ],
//     P_j (3 x 224) = sum_v jx[v] W[v,j] basis[3v..3v+2, :],   s_j = sum_v jx[v] W[v,j]      (once per model),
// so a (F x (3 n_lm + 72) x 224) GEMM u = coef . [basis rows of the landmarks ; P]^T replaces the 20 670-column blend GEMM and
// this kernel replaces the skinning pass: one thread per (frame, landmark) applies the dense 24-joint blend of its vertex,
// one thread per frame combines the 24 weighted centroids.  No vertex other than the landmarks is ever formed.
// A (F,24,12); u (F, ldu): [landmark v_posed (n_lm x 3) | centroids (24 x 3)]; lm_weights (n_lm, 24); s (24).
// ------------------------------------------------------------------------------------------
__global__ void smpl_reduced_joints_kernel(const float* __restrict__ A, const float* __restrict__ u, int64_t ldu,
                                           const float* __restrict__ lm_weights, const float* __restrict__ s,
                                           float* __restrict__ lm_out, float* __restrict__ thorax, int64_t F, int n_lm) {
    pdl_wait_cta();
    pdl_trigger();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int per = n_lm + 1;
    if (i >= F * per) return;
    const int64_t f = i / per;
    const int l = (int)(i % per);
    const float* a = A + f * NJ * 12;
    const float* uf = u + f * ldu;
    if (l < n_lm) {
        float T[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) T[k] = 0.f;
        const float* w = lm_weights + l * NJ;
        for (int j = 0; j < NJ; ++j) {
            const float wj = w[j];
#pragma unroll
            for (int k = 0; k < 12; ++k) T[k] = fmaf(wj, a[j * 12 + k], T[k]);
        }
        const float x = uf[l * 3], y = uf[l * 3 + 1], z = uf[l * 3 + 2];
        float* o = lm_out + (f * n_lm + l) * 3;
        o[0] = T[0] * x + T[1] * y + T[2] * z + T[3];
        o[1] = T[4] * x + T[5] * y + T[6] * z + T[7];
        o[2] = T[8] * x + T[9] * y + T[10] * z + T[11];
    } else {
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
        const float* p = uf + n_lm * 3;
        for (int j = 0; j < NJ; ++j) {
            const float x = p[j * 3], y = p[j * 3 + 1], z = p[j * 3 + 2], sj = s[j];
            const float* t = a + j * 12;
            o0 += t[0] * x + t[1] * y + t[2] * z + t[3] * sj;
            o1 += t[4] * x + t[5] * y + t[6] * z + t[7] * sj;
            o2 += t[8] * x + t[9] * y + t[10] * z + t[11] * sj;
        }
        thorax[f * 3] = o0; thorax[f * 3 + 1] = o1; thorax[f * 3 + 2] = o2;
    }
}

}  // namespace gait



using namespace gait;

extern "C" {

int gait_smpl_pose_chain(const float* R, const float* betas, int64_t ldb, const float* J_template,
                         const float* J_shapedirs, const int32_t* parents, float* A, float* J_posed,
                         float* coef, float* Aop, int64_t F, gait_stream_t stream) {
    GAIT_REQUIRE(F >= 0, "smpl_pose_chain: negative F");
    if (F == 0) return GAIT_OK;
    GAIT_REQUIRE(R && betas && J_template && J_shapedirs && parents && J_posed && (A || Aop), "smpl_pose_chain: null pointer");
    GAIT_REQUIRE(ldb >= NB, "smpl_pose_chain: ldb < 10");
    GAIT_REQUIRE(A == nullptr || aligned16(A), "smpl_pose_chain: A must be 16-byte aligned");
    launch_pdl(4, smpl_pose_chain_kernel, dim3((unsigned)ceil_div(F, kChainWarps)), dim3(kChainWarps * 32), 0, as_stream(stream),
        R, nullptr, 0, 0.f, nullptr, betas, ldb, nullptr, 0, nullptr, J_template, J_shapedirs, parents, A, J_posed, coef, Aop, F);
    return check_launch("smpl_pose_chain");
}

int gait_smpl_pose_chain_rot6d(const float* x6, int64_t ldx6, float eps, const float* betas, int64_t ldb, const float* cam,
                               int64_t ldcam, const float* J_template, const float* J_shapedirs, const int32_t* parents,
                               float* R_out, float* A, float* J_posed, float* coef, float* Aop, float* theta, int64_t F,
                               gait_stream_t stream) {
    GAIT_REQUIRE(F >= 0, "smpl_pose_chain_rot6d: negative F");
    if (F == 0) return GAIT_OK;
    GAIT_REQUIRE(x6 && R_out && betas && J_template && J_shapedirs && parents && J_posed && (A || Aop),
                 "smpl_pose_chain_rot6d: null pointer");
    GAIT_REQUIRE(ldx6 >= 6 * NJ && (ldx6 & 1) == 0 && aligned8(x6) && ldb >= NB, "smpl_pose_chain_rot6d: bad stride or alignment");
    GAIT_REQUIRE(theta == nullptr || (cam && ldcam >= 3), "smpl_pose_chain_rot6d: theta needs cam");
    GAIT_REQUIRE(A == nullptr || aligned16(A), "smpl_pose_chain_rot6d: A must be 16-byte aligned");
    launch_pdl(4, smpl_pose_chain_kernel, dim3((unsigned)ceil_div(F, kChainWarps)), dim3(kChainWarps * 32), 0, as_stream(stream),
        nullptr, x6, ldx6, eps, R_out, betas, ldb, cam, ldcam, theta, J_template, J_shapedirs, parents, A, J_posed, coef, Aop, F);
    return check_launch("smpl_pose_chain_rot6d");
}

int gait_smpl_blend(const float* coef, const float* basis_t, float* v_posed, int64_t ldv, int64_t F, int64_t V3,
                    gait_stream_t stream) {
    GAIT_REQUIRE(F >= 0 && V3 >= 0, "smpl_blend: negative size");
    if (F == 0 || V3 == 0) return GAIT_OK;
    GAIT_REQUIRE(coef && basis_t && v_posed && ldv >= V3, "smpl_blend: null pointer or ldv < 3V");
    return linear_launch(coef, GAIT_BLEND_LD, basis_t, GAIT_BLEND_LD, nullptr, nullptr, 0, v_posed, ldv, F, V3,
                         GAIT_BLEND_LD, as_stream(stream));
}

int gait_smpl_lbs(const float* v_posed, int64_t ldv, const float* A, const float* lbs_weights, float* verts,
                  int64_t F, int64_t V, gait_stream_t stream) {
    GAIT_REQUIRE(F >= 0 && V >= 0, "smpl_lbs: negative size");
    if (F == 0 || V == 0) return GAIT_OK;
    GAIT_REQUIRE(v_posed && A && lbs_weights && verts, "smpl_lbs: null pointer");
    GAIT_REQUIRE(aligned16(A) && aligned8(v_posed) && aligned8(verts), "smpl_lbs: misaligned pointer");
    GAIT_REQUIRE((V & 1) == 0 && (ldv & 1) == 0 && ldv >= 3 * V, "smpl_lbs: V and ldv must be even (8-byte row alignment), ldv >= 3V");
    GAIT_REQUIRE(F < (1ll << 31) && V < (1ll << 31) && ceil_div(F, LBS_FT) < 65536, "smpl_lbs: size too large");
    dim3 grid((unsigned)ceil_div(V, LBS_VT), (unsigned)ceil_div(F, LBS_FT));
    smpl_lbs_kernel<<<grid, LBS_VT, 0, as_stream(stream)>>>(v_posed, ldv, A, lbs_weights, verts, (int)F, (int)V);
    return check_launch("smpl_lbs");
}

int gait_joint_regress(const float* verts, const float* Jreg, float* out, int64_t F, int64_t V, int Rj,
                       gait_stream_t stream) {
    GAIT_REQUIRE(F >= 0 && V >= 0 && Rj >= 0, "joint_regress: negative size");
    if (F == 0 || Rj == 0) return GAIT_OK;
    GAIT_REQUIRE(verts && Jreg && out, "joint_regress: null pointer");
    GAIT_REQUIRE(F < (1ll << 31) && V < (1ll << 31), "joint_regress: size too large");
    dim3 grid((unsigned)ceil_div(F, JR_FT), (unsigned)ceil_div(Rj, JR_RT));
    joint_regress_kernel<<<grid, JR_THREADS, 0, as_stream(stream)>>>(verts, Jreg, out, (int)F, (int)V, Rj);
    return check_launch("joint_regress");
}

int gait_joints_assemble(const float* J_posed, const float* verts, int64_t V, const int32_t* landmarks,
                         int n_landmarks, const float* extra, int n_extra, int extra_parts, int64_t extra_part_stride,
                         const int32_t* joint_map, int J, float* joints, const float* cam, int64_t ldcam, float focal_length, float img_res,
                         float kp2d_divisor, float* kp2d, const int32_t* gather, int n_gather,
                         float* gathered, int64_t F, gait_stream_t stream) {
    GAIT_REQUIRE(F >= 0 && J >= 0 && n_gather >= 0 && n_landmarks >= 0 && n_extra >= 0, "joints_assemble: negative size");
    if (F == 0 || J + n_gather == 0) return GAIT_OK;
    GAIT_REQUIRE(J_posed && joint_map && joints, "joints_assemble: null pointer");
    GAIT_REQUIRE(n_landmarks == 0 || (verts && landmarks), "joints_assemble: landmarks need verts");
    GAIT_REQUIRE(n_extra == 0 || (extra && extra_parts >= 1), "joints_assemble: n_extra > 0 needs extra and extra_parts >= 1");
    GAIT_REQUIRE(kp2d == nullptr || (cam && ldcam >= 3 && aligned8(kp2d)), "joints_assemble: kp2d needs cam");
    GAIT_REQUIRE(n_gather == 0 || (gather && gathered), "joints_assemble: gather needs output");
    GAIT_REQUIRE(n_extra <= JA_MAX_EXTRA, "joints_assemble: at most 9 extra joints");
    launch_pdl(4, joints_assemble_kernel, dim3((unsigned)ceil_div(F, JA_FB)), dim3(256), 0, as_stream(stream),
        J_posed, verts, V, landmarks, n_landmarks, ExtraJoints{extra, n_extra, extra_parts, extra_part_stride}, joint_map, J,
        joints, cam, ldcam, focal_length,
        img_res, kp2d_divisor, kp2d, gather, n_gather, gathered, F);
    return check_launch("joints_assemble");
}

int gait_pack_theta(const float* R, const float* cam, int64_t ldcam, const float* betas, int64_t ldb,
                    float* theta, int64_t F, gait_stream_t stream) {
    GAIT_REQUIRE(F >= 0, "pack_theta: negative F");
    if (F == 0) return GAIT_OK;
    GAIT_REQUIRE(R && cam && betas && theta && ldcam >= 3 && ldb >= NB, "pack_theta: null pointer or bad stride");
    pack_theta_kernel<<<(unsigned)ceil_div(F * 32, 256), 256, 0, as_stream(stream)>>>(R, cam, ldcam, betas, ldb, theta, F);
    return check_launch("pack_theta");
}

int gait_smpl_reduced_joints(const float* A, const float* u, int64_t ldu, const float* lm_weights, const float* s,
                             float* lm_out, float* thorax, int64_t F, int n_lm, gait_stream_t stream) {
    GAIT_REQUIRE(F >= 0 && n_lm > 0 && n_lm <= 1024, "smpl_reduced_joints: bad sizes");
    if (F == 0) return GAIT_OK;
    GAIT_REQUIRE(A && u && lm_weights && s && lm_out && thorax, "smpl_reduced_joints: null pointer");
    GAIT_REQUIRE(ldu >= 3 * n_lm + 3 * GAIT_NUM_JOINTS, "smpl_reduced_joints: ldu < 3 n_lm + 72");
    const int64_t n = F * (n_lm + 1);
    launch_pdl(4, smpl_reduced_joints_kernel, dim3((unsigned)ceil_div(n, 128)), dim3(128), 0, as_stream(stream), A, u, ldu, lm_weights, s,
               lm_out, thorax, F, n_lm);
    return check_launch("smpl_reduced_joints");
}

}  // extern "C"

// Generic joint gather (kp_utils.convert_kps on the device): dst[f,k] = src[f,idx[k]] or 0.
namespace gait {
__global__ void gather_joints_kernel(const float* __restrict__ src, int Js, const int32_t* __restrict__ idx, int Jd,
                                     float* __restrict__ dst, int64_t F) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= F * Jd) return;
    const int64_t f = i / Jd;
    const int g = idx[i % Jd];
    float x = 0.f, y = 0.f, z = 0.f;
    if (g >= 0 && g < Js) {
        const float* p = src + (f * Js + g) * 3;
        x = p[0]; y = p[1]; z = p[2];
    }
    dst[i * 3] = x; dst[i * 3 + 1] = y; dst[i * 3 + 2] = z;
}
}  // namespace gait

extern "C" int gait_gather_joints(const float* src, int Js, const int32_t* idx, int Jd, float* dst, int64_t F,
                                  gait_stream_t stream) {
    GAIT_REQUIRE(F >= 0 && Js >= 0 && Jd >= 0, "gather_joints: negative size");
    if (F == 0 || Jd == 0) return GAIT_OK;
    GAIT_REQUIRE(src && idx && dst, "gather_joints: null pointer");
    gait::gather_joints_kernel<<<(unsigned)gait::ceil_div(F * Jd, 256), 256, 0, gait::as_stream(stream)>>>(src, Js, idx, Jd, dst, F);
    return gait::check_launch("gather_joints");
}
