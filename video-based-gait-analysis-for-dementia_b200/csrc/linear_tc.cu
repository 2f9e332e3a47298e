// Tensor-core path of gait_linear: FP32-accurate split-TF32 ("3xTF32") GEMM on tcgen05.
//
//   D[p,q] = sum_k P[p,k] Q[q,k]      P: 128-row tile (UMMA M = 128 TMEM lanes), Q: BN-row tile
//
// Both operands are K-major FP32 in global memory (torch Linear / GRU layout).  Per k-block of 32
// floats (= one 128-byte swizzle span) the pipeline is
//
//   TMA producers (2 warps)      cp.async.bulk.tensor 2D, SWIZZLE_128B, 64-row boxes, one lane per box, even / odd k-blocks
//                                -> raw P / Q tiles (for prepared weights: the hi tile and the lo tile of Q)
//   converters (2 groups of 4 warps, even / odd k-blocks)
//                                P (the 128-row operand): each thread moves its row from smem into TENSOR MEMORY, split
//                                round-to-nearest hi = RN_tf32(x), lo = RN_tf32(x - hi) (tcgen05.st: columns [hi 32 | lo 32]
//                                of the stage), so the MMAs read it from TMEM and neither P_hi nor P_lo is ever written to
//                                or re-read from shared memory;
//                                Q: nothing when the weight was prepared (QLO: both tiles arrive by TMA); otherwise only the
//                                lo tile is written, lo = RN_tf32(x - trunc(x)) with hi = the raw FP32 word (the tensor core
//                                ignores the low 13 mantissa bits), elementwise, so the TMA swizzle pattern is preserved.
//                                Shared-memory bandwidth (LDS/STS + UMMA operand reads share one 128 B/clk pipe,
//                                ncu: 97 % busy before the operand moved to TMEM) bounded the first version.
//   MMA issuers (2 warps, even / odd accumulator chunks; one lane issues)
//                                per 8-wide k-step three tcgen05.mma.kind::tf32 (A operand in TMEM) into one TMEM
//                                accumulator: P_lo.Q_hi + P_hi.Q_lo + P_hi.Q_hi (cross terms first)
//   promotion warps (8 warps, 120 registers via setmaxnreg)
//                                every 1 or 2 k-blocks (call-site policy, LinearPromote): tcgen05.ld 32x32b of the TMEM
//                                partial sum, added (round-to-nearest) into FP32 registers; the TMEM buffers alternate.
//                                Final: +bias +Cin -> vectorised global stores
//
// with mbarrier landed / converted / empty rings (4 stages for BN = 128, 4-6 for BN = 64; even counts so that every warp of a
// pair always returns to the same stages and sees each barrier phase) and accumulator full / empty rings; CTAs are persistent
// (one per SM) and walk (tile, k-split) work items in weight-tile-major order, so the stores of one item overlap the loads
// and MMAs of the next.  Dropped term: P_lo.Q_lo ~ 2^-22 relative.  The same kernel serves C = A.W^T (P = A) and the
// transposed form (P = W, Q = A) that keeps 128 TMEM lanes busy when A has few rows; split-K work items write partial sums
// that the consumer adds.  Details and measurements: the kernel comment below and DESIGN.md 4.1 / 4.2.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace gait {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int MAX_STAGES = 6;
#ifndef GAIT_P_BOX_ROWS
#define GAIT_P_BOX_ROWS 64
#define GAIT_Q_BOX_ROWS 64
#endif
// Rows per TMA box.  The rows of a box are fetched one after the other and boxes proceed in parallel, but every box also
// costs its issuing warp ~140 cycles: 64-row boxes (4 to 6 per k-block instead of 8 to 12) measured 5 % faster than 32-row ones.
constexpr int P_BOX = GAIT_P_BOX_ROWS;
__host__ __device__ constexpr int q_box(int bn) { return bn < GAIT_Q_BOX_ROWS ? bn : GAIT_Q_BOX_ROWS; }
constexpr int DRAIN_KB_LONG_K = 1;        // k-blocks accumulated in TMEM between promotions to FP32 registers when K >= 512 ...
constexpr int DRAIN_KB_SHORT_K = 2;       // ... and for short contractions (few truncating accumulations anyway; the blend GEMM)
constexpr int THREADS = 640;              // 20 warps, see the kernel comment

constexpr int NDRAIN = 256;
// register budget after the setup (setmaxnreg): 128 x 48 + 256 x 120 + 256 x 96 (converters, unchanged) = 640 x 96
constexpr int REGS_ISSUE = 48, REGS_PROMOTE = 120;

template <int BN>
struct Cfg {
    static constexpr int P_TILE = BM * BK * 4;
    static constexpr int Q_TILE = BN * BK * 4;
    static constexpr int STAGE = P_TILE + 2 * Q_TILE;         // [P raw][Q raw = hi][Q lo]
#ifndef GAIT_BN64_STAGES
#define GAIT_BN64_STAGES 4
#define GAIT_BN64_NBUF 4
#endif
    static constexpr int STAGES = (BN <= 64) ? GAIT_BN64_STAGES : 4;   // x 32 KB (BN 64) or 4 x 48 KB (BN 128)
    static constexpr int STAGING = 8 * 32 * 20 * 4;          // 8 promotion warps x (32 rows x 16 columns, row stride 20)
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM = STAGES * STAGE + STAGING + BAR_BYTES + 1024;   // +1024 alignment slack
    static constexpr int NBUF = (BN <= 64) ? GAIT_BN64_NBUF : 2;   // accumulator buffers (ring)
    static constexpr int TMEM_P0 = NBUF * BN;                 // first column of the P operand ring
    static constexpr int P_COLS = 2 * BK;                     // per stage: [hi 32 | lo 32]
    static constexpr int TMEM_COLS = 512;                     // NBUF * BN + STAGES * P_COLS = 512
    static_assert(NBUF * BN + STAGES * P_COLS <= 512, "tensor memory budget");
    // producers, converter groups and MMA issuers each come in pairs that alternate k-blocks / chunks; with an even number of
    // stages and accumulator buffers every warp always returns to the same stages / buffers and sees each barrier phase
    static_assert(STAGES % 2 == 0 && NBUF % 2 == 0, "stage and accumulator rings must be even");
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: a waiting warp sleeps instead of taking issue slots
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 accumulator columns in one instruction (no wait: the caller batches loads, then tcgen05.wait::ld once)
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
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// one lane of a converged warp (warp-uniform control flow keeps descriptors in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// round-to-nearest TF32 (10 explicit mantissa bits) with integer ops; the remainder x - hi is exact in FP32
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// lo part of the split whose hi part is the raw word (read truncated by the tensor core)
#ifndef GAIT_LO_ROUND
#define GAIT_LO_ROUND 1       // 1: lo = RN_tf32(x - trunc(x)) (|x - hi - lo| <= 2^-22 |x|); 0: lo = x - trunc(x), read truncated (2^-20)
#endif
__device__ __forceinline__ float tf32_lo(float x) {
#if GAIT_LO_ROUND
    return tf32_hi(x - tf32_trunc(x));
#else
    return x - tf32_trunc(x);
#endif
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (8-row x 128-byte atoms, 1024 B apart).
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

// barrier slots (8 B each)
constexpr int NBUF = 4;                    // max TMEM accumulator buffers (barrier slots)
enum : int { B_FULL = 0, B_CONV = MAX_STAGES, B_EMPTY = 2 * MAX_STAGES, B_ACC_FULL = 3 * MAX_STAGES, B_ACC_EMPTY = 3 * MAX_STAGES + NBUF };

// Persistent kernel: each CTA walks work items (p tile, q tile, k split) item = blockIdx.x + i*gridDim.x.
// The smem stage ring and the TMEM accumulator ring run continuously across items, so the TMA / convert /
// MMA roles work on item i+1 while the promotion warps are still storing item i.
//
// Warp roles (20 warps).  No role touches every k-block: the k-block period is otherwise bounded below by the serial
// latency chain of ONE warp's loop body (barrier probe -> work -> fence -> arrive, ~1000 cycles), whatever the
// throughput of the units behind it.
//   0, 2   TMA producers, even / odd k-blocks (one lane per 32-row box)
//   1, 3   MMA issuers, even / odd accumulator chunks (the MMAs of one accumulator stay in one warp, in order); 1 owns TMEM
//   4-11   promotion + epilogue: warp -> TMEM lane quadrant (warp & 3) x column half ((warp - 4) >> 2)
//   12-19  converters: two groups of four warps (all quadrants each), even / odd k-blocks
//
// QLO = true: the BN-row operand (constant weights) was split once (gait_prepare_weight: round-to-nearest hi and lo arrays);
// both tiles are loaded by TMA and the converters only handle the 128-row operand.
template <int BN, bool QLO>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ,
                   const __grid_constant__ CUtensorMap tmQlo, const float* __restrict__ bias, const float* Cin, int64_t ldcin, float* C, int64_t ldc,
                   int P_rows, int Q_rows, int K, int transposed, int kb_per_split, int64_t split_stride,
                   int tiles_p, int splits, int n_items, int mode, int drain, unsigned long long* trace) {
    using cfg = Cfg<BN>;
    constexpr int STAGES = cfg::STAGES;
    constexpr int HN = BN / 2;                                   // accumulator columns per promotion warp
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + STAGES * cfg::STAGE + cfg::STAGING;
    auto BAR = [&](int i) { return bars + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + STAGES * cfg::STAGE + cfg::STAGING + 240);
    float* staging = reinterpret_cast<float*>(gbase + STAGES * cfg::STAGE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb_total = (K + BK - 1) / BK;
    const int drain_kb = (mode == 2) ? (1 << 30) : drain;          // mode 2 (debug): never promote

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(BAR(B_FULL + s), BM / P_BOX + (QLO ? 2 : 1) * (BN / q_box(BN)));   // one arrival per TMA box
            mbar_init(BAR(B_CONV + s), 4);                      // one arrival per warp of the converter group
            mbar_init(BAR(B_EMPTY + s), 1);
        }
        for (int b = 0; b < cfg::NBUF; ++b) {
            mbar_init(BAR(B_ACC_FULL + b), 1);
            mbar_init(BAR(B_ACC_EMPTY + b), NDRAIN / 32);       // one arrival per promotion warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;
    // programmatic dependent launch: everything above overlapped the previous kernel's tail; nothing below may start before
    // that kernel has completed.  The successor may be scheduled from here on (it needs this CTA's SM to become free anyway).
    pdl_wait_cta();
    pdl_trigger();

    // Every role runs its own loop over the CTA's work items (it = k-blocks processed so far = position in the stage ring,
    // ch = accumulator chunks so far = position in the TMEM ring), so that the register budget can be re-divided per
    // warpgroup: the promotion warps hold a 128 x BN/2 FP32 tile plus 32 freshly loaded columns.
#define ITEM_LOOP_BEGIN                                                                                                   \
    for (int item = blockIdx.x, it = 0, ch = 0; item < n_items; item += gridDim.x) {                                     \
        /* item order: k-split fastest, then the 128-row tiles, then the BN-row (weight) tiles - CTAs that run            \
           concurrently share weight tiles, so a large weight matrix is streamed from HBM once, not once per row tile */ \
        const int z = item % splits;                                                                                      \
        const int pt = (item / splits) % tiles_p;                                                                         \
        const int qt = item / (splits * tiles_p);                                                                         \
        const int p0 = pt * BM, q0 = qt * BN;                                                                             \
        const int kb0 = z * kb_per_split;                                                                                 \
        const int nkb = min(kb_per_split, nkb_total - kb0);                                                               \
        const int nchunks = (nkb + drain_kb - 1) / drain_kb;                                                              \
        /* ring slots this item takes: with two k-blocks per chunk an odd count is padded by one EMPTY slot that only     \
           passes through the barriers, so that `it` stays 2 * ch and every issuer keeps its pair of stages */            \
        const int nslots = (drain_kb == 2) ? (nkb + 1) & ~1 : nkb;                                                        \
        (void)pt; (void)p0; (void)q0; (void)kb0; (void)nchunks; (void)ch; (void)nslots;
#define ITEM_LOOP_END                                                                                                     \
        it += nslots;                                                                                                     \
        ch += nchunks;                                                                                                    \
    }
    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_ISSUE));
        if (warp == 0 || warp == 2) {
        ITEM_LOOP_BEGIN
            // ------------------------------------------------------------ TMA producers
            // TMA issue laws measured on this machine (scripts/microbench/kblock_pipe.cu, tma_issue.cu): a TMA warp
            // instruction occupies its warp for ~450 cycles (+ ~50 per extra active lane), the rows of one box are
            // fetched one after the other, boxes issued by different lanes / warps proceed in parallel.
            const int par = warp == 0 ? 0 : 1;
            constexpr int QBOX = q_box(BN), PB = BM / P_BOX, QB = BN / QBOX;
            for (int kb = 0; kb < nslots; ++kb) {
                if (((it + kb) & 1) != par) continue;
                const int s = (it + kb) % STAGES;
                const uint32_t ph = ((it + kb) / STAGES) & 1;
                mbar_wait(BAR(B_EMPTY + s), ph ^ 1);
                const uint32_t st = base + s * cfg::STAGE;
                if (kb >= nkb) {                                            // padding slot: complete the phase, load nothing
                    if (lane < PB + (QLO ? 2 : 1) * QB) mbar_arrive(BAR(B_FULL + s));
                } else if (lane < PB + (QLO ? 2 : 1) * QB) {
                    if (trace && blockIdx.x == 0 && lane == 0 && it + kb < 64) trace[(it + kb) * 4 + 0] = clock64();
                    mbar_arrive_expect_tx(BAR(B_FULL + s), (lane < PB ? P_BOX : QBOX) * BK * 4);
                    if (lane < PB) tma_load_2d(st + lane * (P_BOX * BK * 4), &tmP, (kb0 + kb) * BK, p0 + lane * P_BOX, BAR(B_FULL + s));
                    else if (lane < PB + QB) tma_load_2d(st + cfg::P_TILE + (lane - PB) * (QBOX * BK * 4), &tmQ, (kb0 + kb) * BK, q0 + (lane - PB) * QBOX, BAR(B_FULL + s));
                    else tma_load_2d(st + cfg::P_TILE + cfg::Q_TILE + (lane - PB - QB) * (QBOX * BK * 4), &tmQlo, (kb0 + kb) * BK, q0 + (lane - PB - QB) * QBOX, BAR(B_FULL + s));
                }
                __syncwarp();
            }
        ITEM_LOOP_END
        } else {
        ITEM_LOOP_BEGIN
            // ------------------------------------------------------------ MMA issuers (warp-uniform loop, one lane issues)
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const int par = warp == 1 ? 0 : 1;
            for (int kb = 0; kb < nslots; ++kb) {
                const int cg = ch + kb / drain_kb, buf = cg % cfg::NBUF, use = cg / cfg::NBUF;
                if ((cg & 1) != par) continue;
                const int s = (it + kb) % STAGES;
                const uint32_t ph = ((it + kb) / STAGES) & 1;
                const bool chunk_start = (kb % drain_kb) == 0;
                if (kb >= nkb) {                                            // padding slot: hand the stage straight back
                    mbar_wait(BAR(B_CONV + s), ph);
                    if (elect_one()) umma_commit(BAR(B_EMPTY + s));
                    __syncwarp();
                    continue;
                }
                if (chunk_start && use >= 1) mbar_wait(BAR(B_ACC_EMPTY + buf), (use - 1) & 1);
                mbar_wait(BAR(B_CONV + s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_d + (uint32_t)(buf * BN);
                const uint32_t st = base + s * cfg::STAGE;
                const uint32_t p_hi = tmem_d + (uint32_t)(cfg::TMEM_P0 + s * cfg::P_COLS), p_lo = p_hi + BK;
                const uint64_t q_hi = make_sdesc(st + cfg::P_TILE), q_lo = make_sdesc(st + cfg::P_TILE + cfg::Q_TILE);
                const bool last_of_chunk = (kb % drain_kb) == drain_kb - 1 || kb == nkb - 1;
                if (elect_one()) {
                    if (trace && blockIdx.x == 0 && it + kb < 64) trace[(it + kb) * 4 + 2] = clock64();
                    // The tensor core truncates the FP32 accumulator after every MMA, an error proportional to the accumulator's
                    // magnitude: the small cross terms go first (into a still-small accumulator at the start of a chunk), the
                    // hi*hi products last.
                    if (mode != 1) {
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) {
                            const uint64_t adv = (uint64_t)(k * 32 >> 4);   // 8 floats = 32 bytes along K inside the swizzle span
                            umma_tf32_ts(acc, p_lo + 8 * k, q_hi + adv, idesc, (chunk_start && k == 0) ? 0u : 1u);
                            umma_tf32_ts(acc, p_hi + 8 * k, q_lo + adv, idesc, 1u);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint64_t adv = (uint64_t)(k * 32 >> 4);
                        umma_tf32_ts(acc, p_hi + 8 * k, q_hi + adv, idesc, (mode == 1 && chunk_start && k == 0) ? 0u : 1u);
                    }
                    umma_commit(BAR(B_EMPTY + s));                          // frees the stage when the MMAs retire
                    if (last_of_chunk) umma_commit(BAR(B_ACC_FULL + buf));
                    if (trace && blockIdx.x == 0 && it + kb < 64) trace[(it + kb) * 4 + 3] = clock64();
                }
                __syncwarp();
            }
        ITEM_LOOP_END
        }
    } else if (warp >= 12) {
        ITEM_LOOP_BEGIN
            // ------------------------------------------------------------ converters: FP32 -> (hi = raw word, lo)
            const int grp = (warp - 12) >> 2;
            const int quad = warp & 3;                               // TMEM lane quadrant this warp may write
            const int gt = quad * 32 + lane;                         // thread in the group = P tile row = TMEM lane
            for (int kb = 0; kb < nslots; ++kb) {
                if (((it + kb) & 1) != grp) continue;
                const int s = (it + kb) % STAGES;
                const uint32_t ph = ((it + kb) / STAGES) & 1;
                mbar_wait(BAR(B_FULL + s), ph);
                if (kb >= nkb) {                                            // padding slot
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(B_CONV + s));
                    continue;
                }
                if (trace && blockIdx.x == 0 && gt == 0 && it + kb < 64) trace[(it + kb) * 4 + 1] = clock64();
                uint8_t* st = gbase + s * cfg::STAGE;
                // P: this thread's row (8 x 16 bytes, un-swizzled while reading) -> tensor memory [hi 32 | lo 32]
                {
                    const float4* prow = reinterpret_cast<const float4*>(st + gt * (BK * 4));
                    uint32_t r[BK];
#pragma unroll
                    for (int c = 0; c < BK / 4; ++c) {
                        const float4 x = prow[c ^ (gt & 7)];
                        r[4 * c] = __float_as_uint(x.x); r[4 * c + 1] = __float_as_uint(x.y);
                        r[4 * c + 2] = __float_as_uint(x.z); r[4 * c + 3] = __float_as_uint(x.w);
                    }
                    // this operand goes through registers anyway, so it gets the round-to-nearest split
                    // (hi = RN_tf32(x), lo = RN_tf32(x - hi): |x - hi - lo| <= 2^-23 |x|, |lo| <= 2^-11 |x|)
                    const uint32_t ta = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cfg::TMEM_P0 + s * cfg::P_COLS);
                    uint32_t h[BK];
#pragma unroll
                    for (int j = 0; j < BK; ++j) h[j] = __float_as_uint(tf32_hi(__uint_as_float(r[j])));
                    tmem_st32(ta, h);
#pragma unroll
                    for (int j = 0; j < BK; ++j) r[j] = __float_as_uint(tf32_hi(__uint_as_float(r[j]) - __uint_as_float(h[j])));
                    tmem_st32(ta + BK, r);
                }
                // Q: lo tile only (the raw tile is the hi operand); nothing to do for prepared weights
                constexpr int NQ = QLO ? 0 : cfg::Q_TILE / 16 / 128;
                const float4* q_hi = reinterpret_cast<const float4*>(st + cfg::P_TILE) + gt;
                float4* q_lo = reinterpret_cast<float4*>(st + cfg::P_TILE + cfg::Q_TILE) + gt;
#pragma unroll
                for (int i0 = 0; i0 < NQ; i0 += 4) {
                    float4 v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = q_hi[(i0 + i) * 128];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        q_lo[(i0 + i) * 128] = make_float4(tf32_lo(v[i].x), tf32_lo(v[i].y), tf32_lo(v[i].z), tf32_lo(v[i].w));
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (!QLO) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(B_CONV + s));
            }
        ITEM_LOOP_END
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_PROMOTE));
        ITEM_LOOP_BEGIN
            // ------------------------------------------------------------ promotion + epilogue
            // The tensor core adds into its FP32 accumulator with truncation, so error grows linearly
            // with the number of MMAs per accumulator (measured: 7e-9 * K relative).  Every `drain_kb`
            // k-blocks the TMEM partial sum is therefore added (round-to-nearest) into FP32 registers
            // (measured mean error against FP64 at K = 2048: 1.9e-7 relative with 1 k-block, 3.7e-7 with 2;
            // cuBLAS FP32 7.1e-7).
            const int quad = warp & 3;                              // this warp owns TMEM lanes 32*quad .. +31
            const int half = (warp - 4) >> 2;                       // ... and accumulator columns half*HN .. +HN-1
            float accr[HN];
#pragma unroll
            for (int j = 0; j < HN; ++j) accr[j] = 0.f;
            for (int chunk = 0; chunk < nchunks; ++chunk) {
                const int cg = ch + chunk, buf = cg % cfg::NBUF, use = cg / cfg::NBUF;
                mbar_wait(BAR(B_ACC_FULL + buf), use & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // 32 columns per load instruction and one wait per load: with a promotion every k-block the time of this
                // loop bounds the accumulator ring.  The buffer is handed back as soon as its last values are in registers.
#pragma unroll
                for (int c = 0; c < HN / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32_nowait(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN + half * HN + c * 32), r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (c == HN / 32 - 1) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + buf));
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) accr[c * 32 + j] += __uint_as_float(r[j]);
                }
            }
            const int prow = p0 + quad * 32 + lane;
            const bool first_split = z == 0;
            float* Cz = C + z * split_stride;
            float* stg = staging + (warp - 4) * 32 * 20;
            const bool vec = !transposed && ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(Cz) & 15u) == 0) &&
                             (!(Cin && first_split) || (((ldcin & 3) == 0) && ((reinterpret_cast<uintptr_t>(Cin) & 15u) == 0))) &&
                             (!(bias && first_split) || ((reinterpret_cast<uintptr_t>(bias) & 15u) == 0));
#pragma unroll
            for (int c = 0; c < HN / 16; ++c) {
                const int colbase = q0 + half * HN + c * 16;
                if (transposed) {
                    // D row = output column: lanes run along the contiguous output dimension
                    if (prow < P_rows) {
                        const float bv = (bias && first_split) ? bias[prow] : 0.f;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int qrow = colbase + j;
                            if (qrow < Q_rows) {
                                float o = accr[c * 16 + j] + bv;
                                if (Cin && first_split) o += Cin[(int64_t)qrow * ldcin + prow];
                                Cz[(int64_t)qrow * ldc + prow] = o;
                            }
                        }
                    }
                } else {
                    // transpose 32 rows x 16 columns through smem (row stride 20 floats: conflict-free 16-byte accesses)
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(stg + lane * 20 + j) =
                            make_float4(accr[c * 16 + j], accr[c * 16 + j + 1], accr[c * 16 + j + 2], accr[c * 16 + j + 3]);
                    __syncwarp();
                    if (vec && colbase + 16 <= Q_rows) {
                        const int col = colbase + (lane & 3) * 4;
                        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (bias && first_split) bv = *reinterpret_cast<const float4*>(bias + col);
#pragma unroll
                        for (int r8 = 0; r8 < 32; r8 += 8) {
                            const int rr = r8 + (lane >> 2);
                            const int row = p0 + quad * 32 + rr;
                            if (row < P_rows) {
                                float4 o = *reinterpret_cast<const float4*>(stg + rr * 20 + (lane & 3) * 4);
                                o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                                if (Cin && first_split) {
                                    const float4 cv = *reinterpret_cast<const float4*>(Cin + (int64_t)row * ldcin + col);
                                    o.x += cv.x; o.y += cv.y; o.z += cv.z; o.w += cv.w;
                                }
                                *reinterpret_cast<float4*>(Cz + (int64_t)row * ldc + col) = o;
                            }
                        }
                    } else {
                        const int col = colbase + (lane & 15);
                        if (col < Q_rows) {
                            const float bv = (bias && first_split) ? bias[col] : 0.f;
#pragma unroll 4
                            for (int r2 = 0; r2 < 32; r2 += 2) {
                                const int rr = r2 + (lane >> 4);
                                const int row = p0 + quad * 32 + rr;
                                if (row < P_rows) {
                                    float o = stg[rr * 20 + (lane & 15)] + bv;
                                    if (Cin && first_split) o += Cin[(int64_t)row * ldcin + col];
                                    Cz[(int64_t)row * ldc + col] = o;
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        ITEM_LOOP_END
    }
#undef ITEM_LOOP_BEGIN
#undef ITEM_LOOP_END
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(cfg::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host
// box_rows: P_BOX for the 128-row operand, q_box(BN) for the other one (the kernel issues boxes of exactly these heights)
static int make_map(CUtensorMap* m, const float* ptr, int64_t rows, int64_t K, int64_t ld, int box_rows) {
    return make_tensor_map_2d(m, /*elem_bytes=*/4, ptr, (uint64_t)K, (uint64_t)rows, (uint64_t)ld * sizeof(float), BK,
                              box_rows, /*swizzle128=*/true);
}

unsigned long long* g_trace = nullptr;

template <int BN, bool QLO>
static int launch(const CUtensorMap& tmP, const CUtensorMap& tmQ, const CUtensorMap& tmQlo, const float* bias, const float* Cin,
                  int64_t ldcin, float* C, int64_t ldc, int P_rows, int Q_rows, int K, int transposed, int splits,
                  int64_t split_stride, cudaStream_t stream) {
    using cfg = Cfg<BN>;
    static PerDeviceOnce attr_once;                               // per kernel instantiation and per device
    int dev = 0;
    if (attr_once.needed(&dev)) {
        GAIT_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, QLO>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM));
        attr_once.mark(dev);
    }
    const int nkb = (K + BK - 1) / BK;
    const int kb_per_split = (nkb + splits - 1) / splits;
    const int real_splits = (nkb + kb_per_split - 1) / kb_per_split;
    if (real_splits != splits) {
        set_error("linear(tc): K=%d cannot be cut into %d non-empty splits", K, splits);
        return GAIT_ERR_INVALID;
    }
    const int tiles_p = (int)ceil_div(P_rows, BM), tiles_q = (int)ceil_div(Q_rows, BN);
    const int n_items = tiles_p * tiles_q * splits;
    const int n_sms = device_sm_count();
    // persistent CTAs, one per SM; balance the number of items per CTA
    const int per_cta = (int)ceil_div(n_items, n_sms);
    dim3 grid((unsigned)ceil_div(n_items, per_cta));
    static int max_ctas = -1;                                     // experiment knob: fewer CTAs (is the operand stream limited
    if (max_ctas < 0) {                                           // per SM or by the L2 as a whole?)
        const char* e = getenv("GAITB200_TC_MAXCTAS");
        max_ctas = e ? atoi(e) : 0;
    }
    if (max_ctas > 0 && (int)grid.x > max_ctas) grid.x = (unsigned)max_ctas;
    static int mode = -1, drain_override = 0;
    if (mode < 0) {
        const char* e = getenv("GAITB200_TC_MODE");
        mode = e ? atoi(e) : 0;
        const char* d = getenv("GAITB200_TC_DRAIN");          // experiment knob: k-blocks per promotion
        drain_override = d ? atoi(d) : 0;
    }
    // k-blocks per promotion.  The two MMA-issuing warps own alternate accumulator chunks; each must wait on every phase of
    // the "converted" barriers of the stages it uses (a parity wait on a barrier whose phases a warp skips could pass one
    // fill early), so a chunk always has to cover the same stages: with one k-block per chunk chunk and stage counters stay
    // in step by construction, with two the kernel pads an item with an odd number of k-blocks by one empty ring slot.
    int drain = g_linear_promote_kb > 0 ? g_linear_promote_kb : ((K >= 512) ? DRAIN_KB_LONG_K : DRAIN_KB_SHORT_K);
    if (drain_override > 0) drain = drain_override;
    if (drain != 1 && drain != 2) {
        set_error("linear(tc): GAITB200_TC_DRAIN must be 1 or 2");
        return GAIT_ERR_INVALID;
    }
    cudaError_t le = launch_pdl(1, gemm_tf32x3_kernel<BN, QLO>, grid, dim3(THREADS), cfg::SMEM, stream, tmP, tmQ, tmQlo, bias, Cin, ldcin,
                                C, ldc, P_rows, Q_rows, K, transposed, kb_per_split, split_stride, tiles_p, splits, n_items, mode,
                                drain, g_trace);
    if (le != cudaSuccess) {
        cudaGetLastError();
        set_error("linear(tf32x3 tcgen05): %s", cudaGetErrorString(le));
        return GAIT_ERR_CUDA;
    }
    return check_launch("linear(tf32x3 tcgen05)");
}

}  // namespace tc

thread_local int g_linear_promote_kb = 0;

// debug: per-k-block pipeline timestamps of CTA 0 (trace[kb*4 + {stage free, data landed, converted, MMAs issued}])
void linear_tc_set_trace(unsigned long long* p) { tc::g_trace = p; }

bool linear_tc_eligible(const float* A, int64_t lda, const float* W, int64_t ldw, int64_t M, int64_t N, int64_t K) {
    return ((lda & 3) == 0) && ((ldw & 3) == 0) && aligned16(A) && aligned16(W) && K >= 32 && N >= 64 &&
           (M >= 16 || N >= 512);
}

// ---- prepared weights: lo parts split off once per model (gait_prepare_weight) --------------------------------------------
namespace {
struct PreparedWeight { const float* base; const float* hilo; int64_t n; };     // hilo: [RN hi (n) | lo (n)]
std::mutex g_prep_mutex;
std::vector<PreparedWeight> g_prepared;
thread_local PreparedWeight g_override = {nullptr, nullptr, 0};          // explicit operand of gait_linear_prepared
}  // namespace

PreparedOverride::PreparedOverride(const float* W, const float* hilo, int64_t n)
    : prev_w(g_override.base), prev_hilo(g_override.hilo), prev_n(g_override.n) { g_override = {W, hilo, n}; }
PreparedOverride::~PreparedOverride() { g_override = {prev_w, prev_hilo, prev_n}; }

// prepared hi / lo pointers matching W (which may point inside a registered array); false when W is not prepared
static bool find_prepared(const float* W, int64_t n_needed, const float** hi, const float** lo) {
    if (g_override.base) {                                // explicit handle: no global lookup at all
        const auto& e = g_override;
        if (W >= e.base && W + n_needed <= e.base + e.n) {
            *hi = e.hilo + (W - e.base);
            *lo = e.hilo + e.n + (W - e.base);
            return true;
        }
        return false;
    }
    std::lock_guard<std::mutex> lock(g_prep_mutex);
    for (const auto& e : g_prepared)
        if (W >= e.base && W + n_needed <= e.base + e.n) {
            *hi = e.hilo + (W - e.base);
            *lo = e.hilo + e.n + (W - e.base);
            return true;
        }
    return false;
}

// round-to-nearest split, done once for constant weights: hi = RN_tf32(x), lo = RN_tf32(x - hi)
__global__ void split_hilo_kernel(const float* __restrict__ x, float* __restrict__ hilo, int64_t n) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i], h = tc::tf32_hi(v);
    hilo[i] = h;
    hilo[n + i] = tc::tf32_hi(v - h);
}

// splits > 1: C must hold `splits` partial results `split_stride` floats apart; bias/Cin go into split 0.
int linear_tc_launch(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, const float* Cin,
                     int64_t ldcin, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int splits,
                     int64_t split_stride, cudaStream_t stream) {
    using namespace tc;
    CUtensorMap tmP, tmQ, tmQlo;
    // few activation rows: put the weight rows on the 128 TMEM lanes, activations on the N side
    const bool transposed = M <= 64 || (M < 128 && N >= 128);
    if (transposed) {
        const int bn = (M <= 64) ? 64 : 128;
        GAIT_TRY(make_map(&tmP, W, N, K, ldw, P_BOX));
        GAIT_TRY(make_map(&tmQ, A, M, K, lda, q_box(bn)));
        if (bn == 64)
            return launch<64, false>(tmP, tmQ, tmQ, bias, Cin, ldcin, C, ldc, (int)N, (int)M, (int)K, 1, splits, split_stride, stream);
        return launch<128, false>(tmP, tmQ, tmQ, bias, Cin, ldcin, C, ldc, (int)N, (int)M, (int)K, 1, splits, split_stride, stream);
    }
    // tile width by a two-line cost model: waves of persistent CTAs x the measured k-block period of the tile shape (1300
    // cycles for 128 x 128, 875 for 128 x 64).  256 x 6144 (a recurrence step for 256 sequences): 96 wide tiles
    // in one wave beat 192 narrow ones in two (42 vs 57 us); 1024 x 1024 (regressor): 128 narrow tiles in one wave win.
    int bn = 128;
    if (N > 64) {
        const int64_t sms = device_sm_count(), rt = ceil_div(M, BM);
        const int64_t cost128 = ceil_div(rt * ceil_div(N, 128) * splits, sms) * 1300;
        const int64_t cost64 = ceil_div(rt * ceil_div(N, 64) * splits, sms) * 875;
        if (cost64 < cost128) bn = 64;
    }
    GAIT_TRY(make_map(&tmP, A, M, K, lda, P_BOX));
    const float *Whi = nullptr, *Wlo = nullptr;
    if (find_prepared(W, (N - 1) * ldw + K, &Whi, &Wlo) && aligned16(Whi) && aligned16(Wlo)) {
        GAIT_TRY(make_map(&tmQ, Whi, N, K, ldw, q_box(bn)));          // the prepared hi array replaces the raw weights
        GAIT_TRY(make_map(&tmQlo, Wlo, N, K, ldw, q_box(bn)));
        if (bn == 64)
            return launch<64, true>(tmP, tmQ, tmQlo, bias, Cin, ldcin, C, ldc, (int)M, (int)N, (int)K, 0, splits, split_stride, stream);
        return launch<128, true>(tmP, tmQ, tmQlo, bias, Cin, ldcin, C, ldc, (int)M, (int)N, (int)K, 0, splits, split_stride, stream);
    }
    GAIT_TRY(make_map(&tmQ, W, N, K, ldw, q_box(bn)));
    if (bn == 64)
        return launch<64, false>(tmP, tmQ, tmQ, bias, Cin, ldcin, C, ldc, (int)M, (int)N, (int)K, 0, splits, split_stride, stream);
    return launch<128, false>(tmP, tmQ, tmQ, bias, Cin, ldcin, C, ldc, (int)M, (int)N, (int)K, 0, splits, split_stride, stream);
}

}  // namespace gait

extern "C" {

int gait_prepare_weight(const float* W, float* W_hilo, int64_t n, gait_stream_t stream) {
    float* W_lo = W_hilo;
    GAIT_REQUIRE(n >= 0 && (n == 0 || (W && W_hilo)), "prepare_weight: null pointer or negative size");
    GAIT_REQUIRE((n & 3) == 0 && gait::aligned16(W_hilo), "prepare_weight: n must be a multiple of 4 and the split buffer 16-byte aligned");
    if (n == 0) return GAIT_OK;
    gait::split_hilo_kernel<<<(unsigned)gait::ceil_div(n, 256), 256, 0, gait::as_stream(stream)>>>(W, W_hilo, n);
    GAIT_TRY(gait::check_launch("prepare_weight"));
    std::lock_guard<std::mutex> lock(gait::g_prep_mutex);
    for (auto& e : gait::g_prepared)
        if (e.base == W) { e.hilo = W_lo; e.n = n; return GAIT_OK; }
    gait::g_prepared.push_back({W, W_lo, n});
    return GAIT_OK;
}

int gait_split_weight(const float* W, float* W_hilo, int64_t n, gait_stream_t stream) {
    GAIT_REQUIRE(n >= 0 && (n == 0 || (W && W_hilo)), "split_weight: null pointer or negative size");
    GAIT_REQUIRE((n & 3) == 0 && gait::aligned16(W_hilo), "split_weight: n must be a multiple of 4 and the split buffer 16-byte aligned");
    if (n == 0) return GAIT_OK;
    gait::split_hilo_kernel<<<(unsigned)gait::ceil_div(n, 256), 256, 0, gait::as_stream(stream)>>>(W, W_hilo, n);
    return gait::check_launch("split_weight");
}

int gait_linear_prepared(const float* A, int64_t lda, const float* W, int64_t ldw, const float* W_hilo, int64_t n_prepared,
                         const float* W_base, const float* bias, const float* Cin, int64_t ldcin, float* C, int64_t ldc,
                         int64_t M, int64_t N, int64_t K, gait_stream_t stream) {
    GAIT_REQUIRE(W_hilo && W_base && n_prepared > 0 && gait::aligned16(W_hilo), "linear_prepared: prepared operand missing or misaligned");
    GAIT_REQUIRE(W >= W_base && W + ((N > 0 ? N - 1 : 0) * ldw + K) <= W_base + n_prepared, "linear_prepared: W lies outside the prepared array");
    gait::PreparedOverride scope(W_base, W_hilo, n_prepared);
    return gait::linear_launch(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, gait::as_stream(stream));
}

int gait_release_weight(const float* W) {
    std::lock_guard<std::mutex> lock(gait::g_prep_mutex);
    for (size_t i = 0; i < gait::g_prepared.size(); ++i)
        if (gait::g_prepared[i].base == W) { gait::g_prepared.erase(gait::g_prepared.begin() + i); return GAIT_OK; }
    return GAIT_OK;
}

}  // extern "C"
