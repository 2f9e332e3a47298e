// Linear blend skinning on tcgen05: the dense skinning contraction of smplx lbs()
//
//     T[f,v] (3x4) = sum_j W[v,j] A[f,j]         (M = vertices, N = 12 x frames, K = 24 joints)
//     verts[f,v]   = T[f,v] . [v_posed[f,v]; 1]
//
// is 288 FMA per (vertex, frame) against 24 bytes of HBM traffic - above the FP32 SIMT ridge, so
// the SIMT kernel (smpl.cu) is issue-bound at ~11 % of HBM peak.  Here W.A runs on the tensor
// cores in FP32-accurate split-TF32 form and the CUDA cores only apply T to the vertex, so the
// kernel is bound by the v_posed / verts streams.
//
// One work item = 128 vertices x 8 frames (a CTA loops over up to 4 items of one vertex tile):
//   * operands arrive as three kinds of contiguous blobs via cp.async.bulk (TMA, one mbarrier):
//       - skin weights of the vertex tile, pre-split hi/lo and pre-arranged in the UMMA no-swizzle
//         K-major core-matrix layout (gait_smpl_lbs_pack, once per model):          24 576 B
//       - the 8 frames' skinning transforms, likewise hi/lo + core-matrix layout, written in that
//         form by the kinematic-chain kernel (N = 8 frames x 12 = 96 rows, K = 24):  18 432 B
//       - 8 rows of v_posed (128 vertices x 12 B each):                              8 x 1 536 B
//   * one elected thread issues 9 tcgen05.mma.kind::tf32 (3 k-steps x {lo.hi, hi.lo, hi.hi}),
//     M = 128, N = 96, accumulator in 128 TMEM columns
//   * the 128 threads (thread = vertex = TMEM lane) read T per frame with tcgen05.ld, apply it to
//     the vertex from smem, write the result back to smem; optional fused partial of one
//     J_regressor_extra row (thorax) by warp shuffles; then coalesced 8-byte stores.
//   The weight blob stays in smem for the CTA's lifetime; the per-group operands and the TMEM
//   accumulator are double-buffered, so TMA loads and MMAs of group g+1 run under the epilogue and
//   stores of group g.  86 KB smem + 256 TMEM columns per CTA: 2 CTAs per SM.
#include <algorithm>

#include <cuda.h>

#include "common.cuh"

namespace gait {
namespace lbs {

constexpr int NJ = 24;
constexpr int VT = 128;                    // vertices per CTA (UMMA M)
constexpr int FT = 8;                      // frames per CTA
constexpr int NCOL = FT * 12;              // UMMA N = 96
constexpr int KC = NJ / 4;                 // 16-byte K chunks = 6
constexpr int W_PART = KC * VT * 16;       // 12 288 B (hi or lo)
constexpr int A_PART = KC * NCOL * 16;     //  9 216 B
constexpr int W_BLOB = 2 * W_PART;         // 24 576
constexpr int A_BLOB = 2 * A_PART;         // 18 432
constexpr int V_ROW = VT * 3 * 4;          //  1 536 B per frame

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
// Plain wait for the consumer warps (every register counts there).  They cannot hang alone: a stuck consumer stops releasing
// accumulators and slots, the producer / MMA warps then block in the guarded wait below and trap.
__device__ __forceinline__ void mbar_wait_raw(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: a waiting warp sleeps instead of taking issue slots
    } while (!ok);
}
// Bounded spin (producer, MMA and weight-loader warps): a protocol bug becomes a launch failure (trap) after ~3 s instead of a
// hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: a waiting warp sleeps instead of taking issue slots
    if (ok) return;
    uint32_t t0;
    asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(t0));
    for (uint32_t spins = 1;; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
        if (ok) return;
        if ((spins & 63u) == 0) {
            uint32_t now;
            asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(now));
            if (now - t0 > 3000000000u) __trap();          // 3 s (the low word wraps every 4.29 s)
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// K-major, no swizzle: 8-row x 16-byte core matrices; SBO between 8-row groups, LBO between K chunks.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// same with the M-side operand read from tensor memory (lane = row, one 32-bit column per k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}

// Offset (in floats) of element (row, k) inside one hi/lo part of a core-matrix blob with `rows` rows.
__host__ __device__ __forceinline__ int blob_index(int rows, int row, int k) {
    return ((k >> 2) * (rows >> 3) + (row >> 3)) * 32 + (row & 7) * 4 + (k & 3);
}

// Pack lbs_weights (V,24) into per-tile blobs [tile][hi|lo][kchunk][rowgroup][8][4]; rows >= V are zero.
__global__ void lbs_pack_weights_kernel(const float* __restrict__ W, float* __restrict__ out, int V, int tiles) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= tiles * VT * NJ) return;
    const int tile = i / (VT * NJ), rem = i % (VT * NJ), row = rem / NJ, k = rem % NJ;
    const int v = tile * VT + row;
    const float w = (v < V) ? W[(int64_t)v * NJ + k] : 0.f;
    const float hi = rna_tf32(w), lo = rna_tf32(w - hi);
    float* blob = out + (int64_t)tile * (W_BLOB / 4);
    blob[blob_index(VT, row, k)] = hi;
    blob[W_PART / 4 + blob_index(VT, row, k)] = lo;
}

// Persistent, warp-specialised kernel: one CTA per SM walks a contiguous range of work items
// (vertex tile, 8-frame group) in tile-major order.
//   producer warp (1 thread) : TMA loads into a 4-stage ring (transform blob + 8 v_posed rows per item),
//                              weight blob reloaded when the range crosses into the next vertex tile
//   MMA warp (1 thread)      : 9 tcgen05.mma per item into TMEM buffer = stage
//   NG consumer groups of 4 warps, items dealt round-robin: warp q of a group owns TMEM lanes 32q..
//                              (= vertices): tcgen05.ld T for 8 frames, apply to v_posed in smem, fused
//                              regressor-row partial, coalesced stores, then release the stage
// Loads run up to 3 items ahead and two epilogues are in flight, which hides the TMA/HBM/TMEM latencies.
// Two shared-memory rings, decoupled because their latencies differ by an order of magnitude: the transform blobs come
// from L2 (2.4 MB per 1024 frames, re-read by all 54 vertex tiles) and are released as soon as the MMAs of the item have
// retired; the v_posed rows come from HBM and their slot doubles as the output staging buffer, so it is held until the
// consumer group has stored the skinned vertices.  With ONE ring of 6 combined stages only ~3 HBM loads were in flight per
// SM (the other 3 slots being processed by the 3 consumer groups): the kernel sat at 4.0-4.7 TB/s with nothing saturated.
#ifndef GAIT_LBS_NA
#define GAIT_LBS_NA 3
#endif
#ifndef GAIT_LBS_NV
#define GAIT_LBS_NV 6
#endif
constexpr int NA = GAIT_LBS_NA;                          // transform-blob stages (L2 latency)
constexpr int NV = GAIT_LBS_NV;                          // v_posed stages (HBM latency); a slot is released as soon as its rows are in registers
constexpr int NACC = 4;                                  // TMEM accumulator buffers (4 x 96 = 384 columns)
#ifndef GAIT_LBS_NG
#define GAIT_LBS_NG 3
#endif
constexpr int NG = GAIT_LBS_NG;                          // consumer groups
static_assert(NV % NG == 0, "a consumer group must meet every phase of the v_posed barriers it waits on");
constexpr int V_STAGE = FT * V_ROW;                      // 12 288 B
constexpr int OFF_A = 0;
constexpr int OFF_V = OFF_A + NA * A_BLOB;
constexpr int OFF_BAR = OFF_V + NV * V_STAGE;
constexpr int N_BARS = 2 * NA + 2 * NV + 2 * NACC + 4;
constexpr int SMEM3 = OFF_BAR + ((N_BARS * 8 + 8 + 127) / 128) * 128;
static_assert(SMEM3 <= 232448, "shared memory budget");
constexpr int TMEM_COLS3 = 512;                          // 384 accumulator columns + 2 x 48 weight columns -> 512
constexpr int TMEM_W = NACC * NCOL;                      // first column of the weight operand: 2 buffers x [hi 24 | lo 24]
constexpr int W_COLS = 2 * NJ;
constexpr int NCOMPUTE = NG * 128;
constexpr int W_PROD_A = NCOMPUTE / 32;                  // warp roles after the consumer warps
constexpr int W_MMA = NCOMPUTE / 32 + 1;
constexpr int W_PROD_V = NCOMPUTE / 32 + 2;
constexpr int W_LOADER = NCOMPUTE / 32 + 4;              // 4 warps (aligned to a warpgroup: TMEM lane quarter = warp & 3)
constexpr int THREADS3 = NCOMPUTE + 128 + 128;           // + {A producer, MMA, V producer, idle} + weight-loader warpgroup

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
// one lane of a converged warp (warp-uniform control flow keeps descriptors in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Work-item order.  Items are walked tile-major inside super-blocks of GB frame groups: (super-block, vertex tile, group).
// A CTA's contiguous item range then stays inside one vertex tile for ~47 items (the tile's weights are loaded into tensor
// memory once), and the transform blobs of a super-block (GB x 18 KB = 9.4 MB) stay L2-resident while the 54 tiles sweep
// over them; with one block spanning all frames the 37.7 MB of blobs of a 16 384-frame launch were evicted by the
// v_posed / verts streams between sweeps (4.7 -> 4.1 TB/s).  Up to GB groups (4096 frames) the order is plain tile-major.
constexpr int GB = 512;
struct ItemCursor {
    int tile, g, g_begin, g_end, tiles, groups;
    __device__ __forceinline__ void init(int item, int tiles_, int groups_) {
        tiles = tiles_; groups = groups_;
        const int per_block = tiles * GB;
        const int sb = item / per_block, rem = item - sb * per_block;
        g_begin = sb * GB;
        g_end = min(g_begin + GB, groups);
        const int gc = g_end - g_begin;
        tile = rem / gc;
        g = g_begin + rem % gc;
    }
    __device__ __forceinline__ void next() {
        if (++g == g_end) {
            g = g_begin;
            if (++tile == tiles) {
                tile = 0;
                g_begin = g_end;
                g_end = min(g_begin + GB, groups);
                g = g_begin;
            }
        }
    }
    // last item of this CTA's walk that uses the current tile's weights
    __device__ __forceinline__ bool last_of_tile() const { return g + 1 == g_end; }
};

// MESH = false is the joints-only variant (BASELINE config 5): the skinned vertices never leave the SM; only the `n_lm`
// landmark vertices lm_idx[] the joint sets need are written, to lm_out (F, n_lm, 3), next to the fused regressor row.
template <bool HAS_JX, bool MESH>
__global__ void __launch_bounds__(THREADS3, 1)
smpl_lbs_tc_kernel(const __grid_constant__ CUtensorMap tmV, const float* __restrict__ Aop,
                   const float* __restrict__ Wpack, const float* __restrict__ jx, float* __restrict__ verts,
                   float* __restrict__ jx_partial, const int32_t* __restrict__ lm_idx, int n_lm,
                   float* __restrict__ lm_out, int F, int V, int groups, int tiles, int n_items) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t bar0 = smem_u32(smem + OFF_BAR);
    auto FULL_A = [&](int s) { return bar0 + 8u * s; };             // transform blob landed (tx count)
    auto EMPTY_A = [&](int s) { return bar0 + 8u * (NA + s); };     // the MMAs that read it have retired
    auto FULL_V = [&](int s) { return bar0 + 8u * (2 * NA + s); };  // v_posed rows landed (tx count)
    auto EMPTY_V = [&](int s) { return bar0 + 8u * (2 * NA + NV + s); };     // consumer group done with the slot
    auto MMAD = [&](int a) { return bar0 + 8u * (2 * NA + 2 * NV + a); };    // accumulator a ready
    auto ACCFREE = [&](int a) { return bar0 + 8u * (2 * NA + 2 * NV + NACC + a); };   // accumulator a has been read out
    auto WFULL = [&](int b) { return bar0 + 8u * (2 * NA + 2 * NV + 2 * NACC + b); };     // weight tile b is in tensor memory
    auto WFREE = [&](int b) { return bar0 + 8u * (2 * NA + 2 * NV + 2 * NACC + 2 + b); }; // the MMAs that read weight buffer b have retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 8 * N_BARS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // contiguous, balanced item range of this CTA; item = tile * groups + group
    const int item_lo = (int)(((int64_t)n_items * blockIdx.x) / gridDim.x);
    const int item_hi = (int)(((int64_t)n_items * (blockIdx.x + 1)) / gridDim.x);
    ItemCursor cur0;
    cur0.init(item_lo, tiles, groups);

    if (tid == 0) {
        for (int s = 0; s < NA; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(FULL_A(s)) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(EMPTY_A(s)) : "memory");           // tcgen05.commit
        }
        for (int s = 0; s < NV; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(FULL_V(s)) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(EMPTY_V(s)), "r"(4) : "memory");  // one arrival per consumer warp
        }
        for (int a = 0; a < NACC; ++a) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(MMAD(a)) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ACCFREE(a)), "r"(4) : "memory");
        }
        for (int b = 0; b < 2; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(WFULL(b)) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(WFREE(b)) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS3) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;
    // programmatic dependent launch (common.cuh): the prologue above ran under the tail of the previous kernel (in a step: the
    // blend GEMM, whose CTAs finish at different times); v_posed, the transform blobs and the weights are read after this
    pdl_wait_cta();
    pdl_trigger();

#ifndef GAIT_LBS_REGS_CONSUMER
#define GAIT_LBS_REGS_CONSUMER (GAIT_LBS_NG == 4 ? 96 : 128)   // 0: no setmaxnreg (every warp keeps the kernel's registers)
#endif
#define GAIT_LBS_REGS_OTHER (GAIT_LBS_NG == 4 ? 40 : 48)
    if (warp >= W_PROD_A && warp < W_LOADER) {
    // one setmaxnreg per warpgroup (all four warps must execute the same instruction): the producers and the MMA issuer are
    // small; what they give up lets a consumer thread keep its 96 accumulator columns and 24 v_posed values in registers
#if GAIT_LBS_REGS_CONSUMER
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GAIT_LBS_REGS_OTHER));
#endif
    if (warp == W_PROD_A) {
        // ---------------------------------------------------------------- transform-blob producer (one elected lane)
        ItemCursor c = cur0;
        for (int n = 0; n < item_hi - item_lo; ++n, c.next()) {
            const int s = n % NA;
            if (n >= NA) mbar_wait(EMPTY_A(s), ((n / NA) - 1) & 1);
            if (lane == 0) {
#ifdef GAIT_LBS_EXP_NOA      // timing experiment only (wrong results): the transform blobs are loaded for the first NA items only
                if (n >= NA) { mbar_arrive(FULL_A(s)); } else
#endif
                {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(FULL_A(s)), "r"((uint32_t)A_BLOB) : "memory");
                bulk_g2s(smem_u32(smem + OFF_A + s * A_BLOB), Aop + (int64_t)c.g * (A_BLOB / 4), A_BLOB, FULL_A(s));
                }
            }
            __syncwarp();
        }
    } else if (warp == W_PROD_V) {
        // ---------------------------------------------------------------- v_posed producer: runs up to NV items ahead
        ItemCursor c = cur0;
        for (int n = 0; n < item_hi - item_lo; ++n, c.next()) {
            const int s = n % NV;
            if (n >= NV) mbar_wait(EMPTY_V(s), ((n / NV) - 1) & 1);
            if (lane == 0) {
#ifdef GAIT_LBS_EXP_NOV      // timing experiment only (wrong results): v_posed rows are loaded for the first NV items only
                if (n >= NV) { mbar_arrive(FULL_V(s)); } else
#endif
                // 8 rows x 1536 B of v_posed as one 2D tensor copy (64-bit elements; rows past F read as zero)
                {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(FULL_V(s)), "r"((uint32_t)V_STAGE) : "memory");
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                    ::"r"(smem_u32(smem + OFF_V + s * V_STAGE)), "l"(reinterpret_cast<uint64_t>(&tmV)), "r"(c.tile * (VT * 3 / 2)),
                      "r"(c.g * FT), "r"(FULL_V(s)) : "memory");
                }
            }
            __syncwarp();
        }
    } else if (warp == W_MMA) {
        // ---------------------------------------------------------------- MMA issuer (warp-uniform loop, one lane issues)
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NCOL >> 3) << 17) | ((uint32_t)(VT >> 4) << 24);
        constexpr uint32_t A_LBO = NCOL * 16, SBO = 128;
        int cur_tile = -1, widx = -1;
        ItemCursor c = cur0;
        const int n_total = item_hi - item_lo;
        for (int n = 0; n < n_total; ++n) {
            const int s = n % NA;
            const int tile = c.tile;
            if (tile != cur_tile) {
                ++widx;
                mbar_wait(WFULL(widx & 1), (widx >> 1) & 1);
                cur_tile = tile;
            }
            const int a = n % NACC;
            if (n >= NACC) mbar_wait(ACCFREE(a), ((n / NACC) - 1) & 1);
            mbar_wait(FULL_A(s), (n / NA) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_u32(smem + OFF_A + s * A_BLOB), a_lo = a_hi + A_PART;
            const uint32_t acc = tmem_d + (uint32_t)(a * NCOL);
            const uint32_t w_hi = tmem_d + (uint32_t)(TMEM_W + (widx & 1) * W_COLS), w_lo = w_hi + NJ;
            ItemCursor nx = c;
            nx.next();
            const bool last_of_tile = (n + 1 == n_total) || nx.tile != tile;
            if (elect_one()) {
                // D[128 x 96] = W(128 x 24, tensor memory) . Aop(96 x 24, shared memory)^T, split-TF32: small cross terms first
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
                    for (int ks = 0; ks < NJ / 8; ++ks) {
                        const uint32_t wo = (pass == 0 ? w_lo : w_hi) + 8 * ks;
                        const uint32_t ao = (pass == 1 ? a_lo : a_hi) + 2 * ks * A_LBO;
                        umma_tf32_ts(acc, wo, make_desc(ao, A_LBO, SBO), idesc, (pass | ks) ? 1u : 0u);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(MMAD(a)) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(EMPTY_A(s)) : "memory");
                if (last_of_tile)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(WFREE(widx & 1)) : "memory");
            }
            __syncwarp();
            c = nx;
        }
    }
    } else if (warp >= W_LOADER) {
#if GAIT_LBS_REGS_CONSUMER
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GAIT_LBS_REGS_OTHER));
#endif
        // ---------------------------------------------------------------- weight loader: vertex tile -> tensor memory
        // The skinning weights are the M-side operand of every MMA of a vertex tile (~46 work items per CTA), so
        // they live in tensor memory (lane = vertex, columns [hi 24 | lo 24]) instead of being re-read from shared
        // memory by each of the 9 MMAs of each item.  Two buffers: the next tile is loaded while the current one is used.
        const int row = (warp & 3) * 32 + lane;                    // TMEM lane = vertex within the tile
        ItemCursor c = cur0;
        int cur_tile = -1, i = -1;
        for (int n = 0; n < item_hi - item_lo; ++n, c.next()) {
            if (c.tile == cur_tile) continue;          // same rule as the MMA issuer: one load per run of items of a tile
            cur_tile = c.tile;
            ++i;
            const int tile = c.tile;
            const int b = i & 1;
            if (i >= 2) mbar_wait(WFREE(b), ((i >> 1) - 1) & 1);
            const float* blob = Wpack + (int64_t)tile * (W_BLOB / 4);
            const uint32_t ta = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(TMEM_W + b * W_COLS);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int part = 0; part < 2; ++part) {
#pragma unroll
                for (int c = 0; c < NJ / 8; ++c) {
                    float v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = blob[part * (W_PART / 4) + blob_index(VT, row, c * 8 + k)];
                    tmem_st8(ta + part * NJ + c * 8, v);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(WFULL(b));
        }
    } else if (warp < NCOMPUTE / 32) {
#if GAIT_LBS_REGS_CONSUMER
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GAIT_LBS_REGS_CONSUMER));
#endif
        // ---------------------------------------------------------------- consumer groups
        // Thread = vertex (TMEM lane).  Everything an item needs is pulled into registers at once - the 96 accumulator
        // columns (T for 8 frames) and the vertex's 8 v_posed entries - after which the accumulator and the v_posed slot
        // are released immediately; the skinned vertex is stored straight from registers (three 4-byte stores per frame,
        // a warp covering 384 contiguous bytes: the sectors are completed in L2).  The earlier version staged the result
        // in shared memory for 8-byte stores and read it back for the regressor row: ~430 instead of ~220 warp
        // instructions per item.  What bounds the kernel (measured, scripts/lbs_sweep.py with the GAIT_LBS_EXP_* builds, 1024
        // frames): all loads removed 36 us, consumers removed 33 us, complete 41 us - neither the HBM streams nor the
        // instruction count, but the MMA -> tensor-memory -> register chain: K = 24 in split TF32 is 9 accumulating
        // MMAs of depth 8, each reading and writing the whole 128 x 96 accumulator (0.9 MB of tensor-memory traffic per
        // 24 KB of mesh), against which the consumers' tcgen05.ld compete.  Tried and measured slower or equal: v_posed
        // ring of 9 / 12 slots, two MMA-issuing warps (breaks the in-order argument of the parity waits), 16-frame items
        // with N = 192 (MMA side 33 -> 27 us, consumer side slower), TMA tensor stores (box starts must be 16-byte
        // aligned: impossible for odd frames of a (F, 6890, 3) array, scripts/microbench/tma_store.cu).
        const int grp = warp >> 2, quad = warp & 3;                // group, TMEM lane quarter
        const int gt = tid & 127;                                  // thread within the group = vertex within the tile
        int cur_tile = -1;
        float wjx = 0.f;
        bool any_jx = false;
        int lm_slot = -1, lm_count = 0;
        ItemCursor c = cur0;
        for (int k = 0; k < grp; ++k) c.next();
        for (int n = grp; n < item_hi - item_lo; n += NG) {
            const int s = n % NV;
            const int tile = c.tile, g = c.g;
            const uint32_t ph = (n / NV) & 1;
            const int v0 = tile * VT, nv = min(VT, V - v0);
            const int f0 = g * FT, nf = min(FT, F - f0);
            const float* sV = reinterpret_cast<const float*>(smem + OFF_V + s * V_STAGE);
            if (tile != cur_tile) {
                cur_tile = tile;
                if (HAS_JX) {
                    // regressor-row weight of this thread's vertex; rows of a real SMPL regressor are sparse, and a warp
                    // whose 32 vertices all have weight 0 skips the reduction (it contributes exact zeros)
                    wjx = (gt < nv) ? jx[v0 + gt] : 0.f;
                    any_jx = __any_sync(0xffffffffu, wjx != 0.f);
                }
                if (!MESH || lm_out) {
                    lm_slot = -1; lm_count = 0;
                    if (gt < nv)
                        for (int l = 0; l < n_lm; ++l)
                            if (lm_idx[l] == v0 + gt) { if (lm_count == 0) lm_slot = l; ++lm_count; }
                }
            }
            const int a = n % NACC;
            mbar_wait_raw(FULL_V(s), ph);                          // v_posed rows visible
            mbar_wait_raw(MMAD(a), (n / NACC) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t t[NCOL];
            const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * NCOL);
            tmem_ld32_nowait(taddr, t);
            tmem_ld32_nowait(taddr + 32, t + 32);
            tmem_ld32_nowait(taddr + 64, t + 64);
            float vx[FT], vy[FT], vz[FT];
            const float* p = sV + gt * 3;
#pragma unroll
            for (int f = 0; f < FT; ++f) { vx[f] = p[f * (VT * 3)]; vy[f] = p[f * (VT * 3) + 1]; vz[f] = p[f * (VT * 3) + 2]; }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            // the smem values must have arrived before the slot is handed back to the TMA producer
            asm volatile("" ::"f"(vx[FT - 1]), "f"(vy[FT - 1]), "f"(vz[FT - 1]) : "memory");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(ACCFREE(a));                           // the accumulator can take the MMAs of item n + NACC
                mbar_arrive(EMPTY_V(s));                           // the slot can take the rows of item n + NV
            }
            float ox[FT], oy[FT], oz[FT];
#pragma unroll
            for (int f = 0; f < FT; ++f) {
                const uint32_t* q = t + f * 12;
                const float x = vx[f], y = vy[f], z = vz[f];
                ox[f] = __uint_as_float(q[0]) * x + __uint_as_float(q[1]) * y + __uint_as_float(q[2]) * z + __uint_as_float(q[3]);
                oy[f] = __uint_as_float(q[4]) * x + __uint_as_float(q[5]) * y + __uint_as_float(q[6]) * z + __uint_as_float(q[7]);
                oz[f] = __uint_as_float(q[8]) * x + __uint_as_float(q[9]) * y + __uint_as_float(q[10]) * z + __uint_as_float(q[11]);
            }
            if (MESH && gt < nv) {
                float* dst = verts + ((int64_t)f0 * V + v0 + gt) * 3;
#pragma unroll
                for (int f = 0; f < FT; ++f)
                    if (f < nf) {
                        float* d = dst + (int64_t)f * V * 3;
                        d[0] = ox[f]; d[1] = oy[f]; d[2] = oz[f];
                    }
            }
            if ((!MESH || lm_out) && lm_count > 0) {
                // landmark vertices (the only output in joints-only mode; next to a mesh that goes to peer memory they keep
                // the joint assembly's reads local)
                for (int l = lm_slot; l < n_lm; ++l) {
                    if (l != lm_slot && lm_idx[l] != v0 + gt) continue;
#pragma unroll
                    for (int f = 0; f < FT; ++f)
                        if (f < nf) {
                            float* o = lm_out + ((int64_t)(f0 + f) * n_lm + l) * 3;
                            o[0] = ox[f]; o[1] = oy[f]; o[2] = oz[f];
                        }
                    if (lm_count == 1) break;
                }
            }
            if (HAS_JX) {
                // partial of the regressor row over this warp's 32 vertices: 24 sums (8 frames x 3) reduced with a transposed
                // butterfly - at every level a lane keeps one half of its values and sends the other - 24 shuffles in all;
                // afterwards lane l holds the sum for frame 4*b4 + 2*b3 + b2 (b = bits of l), component (l & 3), lanes with (l & 3) == 3 idle
                float r = 0.f;
                if (any_jx) {
                    float P[24];
#pragma unroll
                    for (int f = 0; f < FT; ++f) { P[f * 3] = wjx * ox[f]; P[f * 3 + 1] = wjx * oy[f]; P[f * 3 + 2] = wjx * oz[f]; }
                    const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2, b1 = lane & 1;
                    float Q[12], R6[6], S[4], U[2];
#pragma unroll
                    for (int i = 0; i < 12; ++i) {
                        const float send = b16 ? P[i] : P[i + 12], keep = b16 ? P[i + 12] : P[i];
                        Q[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const float send = b8 ? Q[i] : Q[i + 6], keep = b8 ? Q[i + 6] : Q[i];
                        R6[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const float send = b4 ? R6[i] : R6[i + 3], keep = b4 ? R6[i + 3] : R6[i];
                        S[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                    S[3] = 0.f;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float send = b2 ? S[i] : S[i + 2], keep = b2 ? S[i + 2] : S[i];
                        U[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                    }
                    {
                        const float send = b1 ? U[0] : U[1], keep = b1 ? U[1] : U[0];
                        r = keep + __shfl_xor_sync(0xffffffffu, send, 1);
                    }
                }
                const int pf = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1), pc = lane & 3;
                if (pc < 3 && pf < nf) jx_partial[((int64_t)(tile * 4 + quad) * F + f0 + pf) * 3 + pc] = r;
            }
#pragma unroll
            for (int k = 0; k < NG; ++k) c.next();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == W_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS3) : "memory");
    }
}

}  // namespace lbs
}  // namespace gait

using namespace gait;

extern "C" {

size_t gait_smpl_lbs_pack_bytes(int64_t V) {
    return V <= 0 ? 0 : (size_t)ceil_div(V, lbs::VT) * lbs::W_BLOB;
}

int gait_smpl_lbs_pack(const float* lbs_weights, float* packed, int64_t V, gait_stream_t stream) {
    GAIT_REQUIRE(V >= 0, "smpl_lbs_pack: negative V");
    if (V == 0) return GAIT_OK;
    GAIT_REQUIRE(lbs_weights && packed, "smpl_lbs_pack: null pointer");
    const int tiles = (int)ceil_div(V, lbs::VT);
    const int n = tiles * lbs::VT * lbs::NJ;
    lbs::lbs_pack_weights_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(lbs_weights, packed, (int)V, tiles);
    return check_launch("smpl_lbs_pack");
}

int64_t gait_smpl_lbs_jx_parts(int64_t V) { return V <= 0 ? 0 : 4 * ceil_div(V, lbs::VT); }

size_t gait_smpl_lbs_aop_bytes(int64_t F) {
    return F <= 0 ? 0 : (size_t)ceil_div(F, lbs::FT) * lbs::A_BLOB;
}

static int lbs_tc_launch(const float* v_posed, int64_t ldv, const float* Aop, const float* Wpack, const float* jx,
                         float* verts, float* jx_partial, const int32_t* lm_idx, int n_lm, float* lm_out, int64_t F,
                         int64_t V, cudaStream_t stream) {
    GAIT_REQUIRE(F >= 0 && V >= 0, "smpl_lbs_tc: negative size");
    if (F == 0 || V == 0) return GAIT_OK;
    const bool mesh = verts != nullptr;
    GAIT_REQUIRE(v_posed && Aop && Wpack, "smpl_lbs_tc: null pointer");
    GAIT_REQUIRE(mesh || lm_out, "smpl_lbs_tc: neither a mesh nor a landmark output given");
    GAIT_REQUIRE(!lm_out || (lm_idx && n_lm > 0 && n_lm <= 1024), "smpl_lbs_tc: landmark output needs 1..1024 landmark ids");
    GAIT_REQUIRE((jx == nullptr) == (jx_partial == nullptr), "smpl_lbs_tc: jx and jx_partial go together");
    GAIT_REQUIRE((V & 1) == 0 && (!mesh || aligned8(verts)), "smpl_lbs_tc: V must be even and verts 8-byte aligned");
    const int64_t tiles = ceil_div(V, lbs::VT);
    GAIT_REQUIRE(ldv >= tiles * lbs::VT * 3 && (ldv & 3) == 0 && aligned16(v_posed) && F < (1ll << 31),
                 "smpl_lbs_tc: v_posed rows must be padded to 384*ceil(V/128) floats (ldv %% 4 == 0, 16-byte aligned)");
    GAIT_REQUIRE(aligned16(Aop) && aligned16(Wpack), "smpl_lbs_tc: operand blobs must be 16-byte aligned");
    GAIT_REQUIRE(F < (1ll << 31) && V < (1ll << 31) && ceil_div(F, lbs::FT) < 65536, "smpl_lbs_tc: size too large");
    static PerDeviceOnce attr_once;
    int dev = 0;
    if (attr_once.needed(&dev)) {
        GAIT_CUDA(cudaFuncSetAttribute(lbs::smpl_lbs_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lbs::SMEM3));
        GAIT_CUDA(cudaFuncSetAttribute(lbs::smpl_lbs_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lbs::SMEM3));
        GAIT_CUDA(cudaFuncSetAttribute(lbs::smpl_lbs_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lbs::SMEM3));
        GAIT_CUDA(cudaFuncSetAttribute(lbs::smpl_lbs_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lbs::SMEM3));
#if GAIT_LBS_REGS_CONSUMER
        // setmaxnreg.inc waits until the CTA's register pool (threads x the kernel's register count) can supply the consumers'
        // budget; a build whose register count came out lower would hang, so refuse it here
        cudaFuncAttributes fa;
        GAIT_CUDA(cudaFuncGetAttributes(&fa, lbs::smpl_lbs_tc_kernel<true, true>));
        if ((int64_t)fa.numRegs * lbs::THREADS3 < (int64_t)lbs::NCOMPUTE * GAIT_LBS_REGS_CONSUMER + (lbs::THREADS3 - lbs::NCOMPUTE) * GAIT_LBS_REGS_OTHER) {
            set_error("smpl_lbs_tc: kernel compiled to %d registers, the setmaxnreg budget needs %d", fa.numRegs,
                      (int)((lbs::NCOMPUTE * GAIT_LBS_REGS_CONSUMER + (lbs::THREADS3 - lbs::NCOMPUTE) * GAIT_LBS_REGS_OTHER) / lbs::THREADS3));
            return GAIT_ERR_UNSUPPORTED;
        }
#endif
        attr_once.mark(dev);
    }
    const int n_sms = device_sm_count();
    const int64_t groups = ceil_div(F, lbs::FT);
    const int64_t n_items = tiles * groups;
    GAIT_REQUIRE(n_items < (1ll << 31), "smpl_lbs_tc: too many work items");
    const unsigned grid = (unsigned)std::min<int64_t>(n_items, n_sms);          // persistent: one CTA per SM
    CUtensorMap tmV;
    // inner extent = the 3V written floats (8-byte elements, V even): the padding of the last tile is zero-filled by TMA
    // instead of being read from the (never written) tail of the v_posed rows
    GAIT_TRY(make_tensor_map_2d(&tmV, 8, v_posed, (uint64_t)(V * 3 / 2), (uint64_t)F, (uint64_t)ldv * sizeof(float),
                                lbs::VT * 3 / 2, lbs::FT, false));
#define GAIT_LBS_LAUNCH(JX, MESH)                                                                                       \
    launch_pdl(2, lbs::smpl_lbs_tc_kernel<JX, MESH>, dim3(grid), dim3(lbs::THREADS3), lbs::SMEM3, stream,                    \
               tmV, Aop, Wpack, jx, verts, jx_partial, lm_idx, n_lm, lm_out, (int)F, (int)V, (int)groups, (int)tiles, (int)n_items)
    if (jx && mesh) GAIT_LBS_LAUNCH(true, true);
    else if (jx) GAIT_LBS_LAUNCH(true, false);
    else if (mesh) GAIT_LBS_LAUNCH(false, true);
    else GAIT_LBS_LAUNCH(false, false);
#undef GAIT_LBS_LAUNCH
    return check_launch("smpl_lbs_tc");
}

int gait_smpl_lbs_tc(const float* v_posed, int64_t ldv, const float* Aop, const float* Wpack, const float* jx,
                     float* verts, float* jx_partial, int64_t F, int64_t V, gait_stream_t stream) {
    GAIT_REQUIRE(verts != nullptr || (F == 0 || V == 0), "smpl_lbs_tc: null pointer");
    return lbs_tc_launch(v_posed, ldv, Aop, Wpack, jx, verts, jx_partial, nullptr, 0, nullptr, F, V, as_stream(stream));
}

int gait_smpl_lbs_tc_ex(const float* v_posed, int64_t ldv, const float* Aop, const float* Wpack, const float* jx,
                        float* verts, float* jx_partial, const int32_t* lm_idx, int n_lm, float* lm_out, int64_t F,
                        int64_t V, gait_stream_t stream) {
    return lbs_tc_launch(v_posed, ldv, Aop, Wpack, jx, verts, jx_partial, lm_idx, n_lm, lm_out, F, V, as_stream(stream));
}

int gait_smpl_lbs_tc_joints(const float* v_posed, int64_t ldv, const float* Aop, const float* Wpack, const float* jx,
                            float* jx_partial, const int32_t* lm_idx, int n_lm, float* lm_out, int64_t F, int64_t V,
                            gait_stream_t stream) {
    GAIT_REQUIRE(n_lm > 0 && n_lm <= 1024, "smpl_lbs_tc_joints: 1..1024 landmark vertices");
    return lbs_tc_launch(v_posed, ldv, Aop, Wpack, jx, nullptr, jx_partial, lm_idx, n_lm, lm_out, F, V, as_stream(stream));
}

}  // extern "C"
