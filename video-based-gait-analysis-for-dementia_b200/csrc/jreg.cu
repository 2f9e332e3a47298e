// Joint regression out[f,r,:] = sum_v Jreg[r,v] verts[f,v,:] (smplx vertices2joints; lib/models/pare.py:70-76 and
// lib/models/spin.py:279-282 with the (17,6890) H36M regressor; lib/models/smpl.py:113 with J_regressor_extra) as ONE
// streaming pass over the mesh: 82 680 B read per frame against 12*Rj B written - an HBM stream.
//
// It is also 17 FMA per mesh float, i.e. ~53 % of the nominal FP32 FMA rate at the HBM roofline, so the inner loop must
// spend its issue slots on FMAs only (measured on B200, scripts/jreg_time.py with the GAIT_JREG_EXP_* builds, 1024 frames:
// whole kernel 34 us; without vertex loads 31 us; with 1 row of FMAs instead of 17 24 us; with neither 12 us - the FMA
// phase, not the HBM stream, bounds it.  Row pairs on packed fma.rn.f32x2 were tried: 39 us, the operand duplication and
// 168 registers cost more than the halved FMA count buys):
//   * lane = frame.  A warp owns 32 frames, so the regressor weights it multiplies with are warp-uniform: they are read
//     from shared memory as broadcast 128-bit loads (one wavefront per 4 weights per 32 frames) and every frame's 3*Rj
//     running sums stay in registers for the whole pass - no cross-lane reduction in the loop.
//   * the weights are pre-packed once per regressor (gait_joint_regress_pack: [4-vertex group][row][4], zero padded) and
//     staged per pipeline stage with ONE TMA bulk copy (cp.async.bulk + mbarrier complete_tx).
//   * the vertices are streamed with cp.async (LDGSTS) straight into a padded shared-memory tile - frame rows are 82 680 B
//     apart, i.e. only 8-byte aligned for odd frames, which rules out TMA (16-byte alignment of address and stride) and
//     128-bit loads; a warp's 32 lanes copy 256 contiguous bytes per instruction.  Row pitch 148 floats makes the transposed
//     reads (lane = row) conflict-free 128-bit loads.  4 stages x 2 CTAs per SM keep ~110 KB of loads in flight per SM.
//   * the vertex range is split over a cluster of 8 CTAs (256 CTAs at 1024 frames: one wave at 2 CTAs / SM); the eight
//     partial sums are combined through distributed shared memory in a fixed order (deterministic), every rank one eighth
//     of the values; no second kernel, no workspace, no atomics.
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gait {
namespace jreg {
using namespace tcu;

constexpr int FB = 32;                 // frames per cluster (lane = frame)
constexpr int CL = 8;                  // CTAs per cluster = vertex-range splits
constexpr int GS = 12;                 // 4-vertex groups per stage (48 vertices, 576 B per frame row)
#ifndef GAIT_JREG_NW
#define GAIT_JREG_NW 6
#endif
constexpr int NW = GAIT_JREG_NW;       // warps per CTA, GS / NW groups each per stage
constexpr int GPW = GS / NW;
#ifndef GAIT_JREG_D
#define GAIT_JREG_D 4
#endif
constexpr int D = GAIT_JREG_D;         // pipeline stages
constexpr int THREADS = NW * 32;
constexpr int XROW = GS * 12;          // floats per frame row per stage
constexpr int XPITCH = XROW + 4;       // 148: (pitch / 4) odd -> 8 rows hit 8 distinct 16-byte bank groups
constexpr int X_BYTES = FB * XPITCH * 4;
static_assert(GS % NW == 0 && (XPITCH / 4) % 2 == 1, "tile shape");

template <int JT> struct Cfg {
    static constexpr int W_BYTES = GS * JT * 16;
    static constexpr int STAGE = X_BYTES + W_BYTES;                    // 22 208 B for JT = 17
    static constexpr int RED_BYTES = NW * JT * 3 * 32 * 4;             // cross-warp reduction scratch (aliases the stages)
    static constexpr int PART_BYTES = JT * 3 * 32 * 4;                 // this CTA's partial, read by rank 0 through DSMEM
    static constexpr int OFF_PART = (D * STAGE > RED_BYTES ? D * STAGE : RED_BYTES);
    static constexpr int OFF_BAR = OFF_PART + PART_BYTES;
    static constexpr int SMEM = OFF_BAR + 64;
};

__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 8 : 0;                                       // src-size 0: the 8 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ float ld_cluster_f32(uint32_t cluster_addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
    return v;
}

// packed (Gpad, JT, 4) with packed[g][j][k] = Jreg[r0 + j][4 g + k], zero outside (rows >= Rj, vertices >= V)
__global__ void jreg_pack_kernel(const float* __restrict__ Jreg, float* __restrict__ packed, int V, int Rj, int JT, int Gpad,
                                 int blocks) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t per_block = (int64_t)Gpad * JT * 4;
    if (i >= per_block * blocks) return;
    const int rb = (int)(i / per_block);
    const int rem = (int)(i % per_block);
    const int g = rem / (JT * 4), j = (rem / 4) % JT, k = rem & 3;
    const int r = rb * JT + j, v = 4 * g + k;
    packed[i] = (r < Rj && v < V) ? Jreg[(int64_t)r * V + v] : 0.f;
}

template <int JT>
__global__ void __launch_bounds__(THREADS, 2)
joint_regress_stream_kernel(const float* __restrict__ verts, const float* __restrict__ packed, float* __restrict__ out,
                            int F, int V, int Rj, int r0, int gpc) {
    using C = Cfg<JT>;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + C::OFF_BAR;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int f0 = (blockIdx.x / CL) * FB;
    const int nf = min(FB, F - f0);
    const int V3 = V * 3;
    const int g_lo = (int)rank * gpc;                                  // first 4-vertex group of this CTA
    const int nst = gpc / GS;

    if (tid == 0) {
        for (int s = 0; s < D; ++s) mbar_init(bar0 + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // one stage: thread 0 bulk-copies the packed weights of GS groups; all threads copy the 32 x 576 B vertex tile
    auto issue = [&](int it) {
        const int slot = it % D;
        const uint32_t st = sbase + slot * C::STAGE;
        const int g0 = g_lo + it * GS;
        if (tid == 0) {
            mbar_arrive_expect_tx(bar0 + 8u * slot, C::W_BYTES);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(st + X_BYTES), "l"(packed + (int64_t)g0 * JT * 4), "r"((uint32_t)C::W_BYTES), "r"(bar0 + 8u * slot)
                         : "memory");
        }
        const int e0 = g0 * 12;                                        // first float of the tile inside a frame row
        constexpr int C8 = XROW / 2;                                   // 8-byte chunks per row = 72
#pragma unroll
        for (int k = 0; k < (FB * C8) / THREADS; ++k) {
            const int i = tid + k * THREADS;
            const int row = i / C8, c8 = i % C8;
            const int e = e0 + 2 * c8;
            const bool ok = row < nf && e < V3;
            const float* src = ok ? verts + ((int64_t)(f0 + row) * V3 + e) : verts;
            cp_async8(st + (uint32_t)(row * XPITCH + 2 * c8) * 4u, src, ok);
        }
    };
    static_assert((FB * (XROW / 2)) % THREADS == 0, "copy loop");

#pragma unroll
    for (int s = 0; s < D - 1; ++s) {
        if (s < nst) issue(s);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    float acc[JT][3];
#pragma unroll
    for (int j = 0; j < JT; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; }

    for (int it = 0; it < nst; ++it) {
        const int slot = it % D;
        asm volatile("cp.async.wait_group %0;" ::"n"(D - 2) : "memory");      // this thread's copies of stage `it` have landed
#ifdef GAIT_JREG_EXP_NOLOAD
        if (it < D - 1)
#endif
        mbar_wait(bar0 + 8u * slot, (it / D) & 1);                            // ... and the weights
        __syncthreads();                                                      // everyone's copies; everyone is done with stage it-1
#ifdef GAIT_JREG_EXP_NOLOAD      // timing experiment only (wrong results): no vertex / weight traffic after the prologue
        if (false)
#endif
        if (it + D - 1 < nst) issue(it + D - 1);                              // refills the slot of stage it-1
        asm volatile("cp.async.commit_group;" ::: "memory");
        const float* sX = reinterpret_cast<const float*>(smem + slot * C::STAGE) + lane * XPITCH;
        const float4* sW = reinterpret_cast<const float4*>(smem + slot * C::STAGE + X_BYTES);
#pragma unroll
        for (int q = 0; q < GPW; ++q) {
            const int gi = warp * GPW + q;
            const float4 x0 = *reinterpret_cast<const float4*>(sX + gi * 12);
            const float4 x1 = *reinterpret_cast<const float4*>(sX + gi * 12 + 4);
            const float4 x2 = *reinterpret_cast<const float4*>(sX + gi * 12 + 8);
            // vertices: (x0.x x0.y x0.z) (x0.w x1.x x1.y) (x1.z x1.w x2.x) (x2.y x2.z x2.w)
#ifdef GAIT_JREG_EXP_ROWS        // timing experiment only (wrong results): FMAs for the first rows only
            constexpr int JN = GAIT_JREG_EXP_ROWS;
#else
            constexpr int JN = JT;
#endif
#pragma unroll
            for (int j = 0; j < JN; ++j) {
                const float4 w = sW[gi * JT + j];                             // warp-uniform address: broadcast
                acc[j][0] = fmaf(w.x, x0.x, fmaf(w.y, x0.w, fmaf(w.z, x1.z, fmaf(w.w, x2.y, acc[j][0]))));
                acc[j][1] = fmaf(w.x, x0.y, fmaf(w.y, x1.x, fmaf(w.z, x1.w, fmaf(w.w, x2.z, acc[j][1]))));
                acc[j][2] = fmaf(w.x, x0.z, fmaf(w.y, x1.y, fmaf(w.z, x2.x, fmaf(w.w, x2.w, acc[j][2]))));
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                                   // the stage memory becomes the reduction scratch

    // cross-warp sum (fixed order) -> this CTA's partial [JT*3][32]
    float* red = reinterpret_cast<float*>(smem);
    float* part = reinterpret_cast<float*>(smem + C::OFF_PART);
#pragma unroll
    for (int j = 0; j < JT; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) red[(warp * JT * 3 + j * 3 + c) * 32 + lane] = acc[j][c];
    __syncthreads();
    for (int i = tid; i < JT * 3 * 32; i += THREADS) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += red[w * JT * 3 * 32 + i];
        part[i] = s;
    }
    cluster_sync_all();                                                // partials of all 8 CTAs are complete and visible
    {
        // every rank combines one eighth of the values (same fixed order over the ranks as before: deterministic) - with rank 0
        // alone each of its threads made 9 dependent trips through distributed shared memory (~3 us of a 34 us kernel)
        constexpr int N = JT * 3 * 32, PER = (N + CL - 1) / CL;
        const uint32_t pa = smem_u32(part);
        const int i_end = min(N, ((int)rank + 1) * PER);
        for (int i = (int)rank * PER + tid; i < i_end; i += THREADS) {
            const int jc = i >> 5, l = i & 31;
            const int j = jc / 3, c = jc % 3;
            float s = 0.f;
#pragma unroll
            for (uint32_t r = 0; r < (uint32_t)CL; ++r) s += ld_cluster_f32(map_to_rank(pa + 4u * i, r));
            if (l < nf && r0 + j < Rj) out[((int64_t)(f0 + l) * Rj + r0 + j) * 3 + c] = s;
        }
    }
    cluster_sync_all();                                                // nobody leaves while a peer may still read its partial
}

inline int rows_per_pass(int Rj) { return Rj <= 9 ? 9 : 17; }
inline int64_t groups_per_cta(int64_t V) { return ceil_div(ceil_div(ceil_div(V, 4), CL), GS) * GS; }

template <int JT>
static int launch(const float* verts, const float* packed, float* out, int64_t F, int64_t V, int Rj, cudaStream_t stream) {
    using C = Cfg<JT>;
    static PerDeviceOnce attr_once;
    int dev = 0;
    if (attr_once.needed(&dev)) {
        GAIT_CUDA(cudaFuncSetAttribute(joint_regress_stream_kernel<JT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        attr_once.mark(dev);
    }
    const int64_t gpc = groups_per_cta(V);
    const int blocks = (int)ceil_div(Rj, JT);
    for (int rb = 0; rb < blocks; ++rb) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(CL * ceil_div(F, FB)));
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = C::SMEM;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, joint_regress_stream_kernel<JT>, verts, packed + (int64_t)rb * CL * gpc * JT * 4, out,
                                           (int)F, (int)V, Rj, rb * JT, (int)gpc);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("joint_regress_packed: launch failed: %s", cudaGetErrorString(e));
            return GAIT_ERR_CUDA;
        }
        GAIT_TRY(check_launch("joint_regress_packed"));
    }
    return GAIT_OK;
}

}  // namespace jreg
}  // namespace gait

using namespace gait;

extern "C" {

size_t gait_joint_regress_pack_bytes(int64_t V, int Rj) {
    if (V <= 0 || Rj <= 0) return 0;
    const int JT = jreg::rows_per_pass(Rj);
    return (size_t)ceil_div(Rj, JT) * jreg::CL * jreg::groups_per_cta(V) * JT * 4 * sizeof(float);
}

int gait_joint_regress_pack(const float* Jreg, float* packed, int64_t V, int Rj, gait_stream_t stream) {
    GAIT_REQUIRE(V >= 0 && Rj >= 0, "joint_regress_pack: negative size");
    if (V == 0 || Rj == 0) return GAIT_OK;
    GAIT_REQUIRE(Jreg && packed && aligned16(packed), "joint_regress_pack: null or misaligned pointer");
    GAIT_REQUIRE(V < (1ll << 28), "joint_regress_pack: V too large");
    const int JT = jreg::rows_per_pass(Rj);
    const int blocks = (int)ceil_div(Rj, JT);
    const int Gpad = (int)(jreg::CL * jreg::groups_per_cta(V));
    const int64_t n = (int64_t)blocks * Gpad * JT * 4;
    jreg::jreg_pack_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(Jreg, packed, (int)V, Rj, JT, Gpad, blocks);
    return check_launch("joint_regress_pack");
}

int gait_joint_regress_packed(const float* verts, const float* packed, float* out, int64_t F, int64_t V, int Rj,
                              gait_stream_t stream) {
    GAIT_REQUIRE(F >= 0 && V >= 0 && Rj >= 0, "joint_regress_packed: negative size");
    if (F == 0 || Rj == 0) return GAIT_OK;
    GAIT_REQUIRE(verts && packed && out, "joint_regress_packed: null pointer");
    GAIT_REQUIRE(aligned8(verts) && (V % 2) == 0 && aligned16(packed),
                 "joint_regress_packed: verts must be 8-byte aligned with an even vertex count, packed 16-byte aligned");
    GAIT_REQUIRE(F < (1ll << 31) - 64 && V < (1ll << 28) && ceil_div(F, jreg::FB) * jreg::CL < (1ll << 31), "joint_regress_packed: size too large");
    if (jreg::rows_per_pass(Rj) == 9) return jreg::launch<9>(verts, packed, out, F, V, Rj, as_stream(stream));
    return jreg::launch<17>(verts, packed, out, F, V, Rj, as_stream(stream));
}

}  // extern "C"
