// The recurrence of one torch.nn.GRU layer/direction (TemporalEncoder; gait_feat_encoder.py:51-57,88) as ONE
// persistent kernel: all T steps, the hidden-state GEMM h_{t-1}.W_hh^T on tcgen05 in FP32-accurate split-TF32
// form, the gate math and the TemporalEncoder residual; steps are chained by per-CTA release/acquire flags.
//
// Decomposition (H = 2048: 128 CTAs = 64 clusters of 2, one CTA per SM, all co-resident):
//   * a cluster of KG = 2 CTAs owns 32 hidden units = 96 rows of W_hh (r, z and n gate rows of those units);
//     CTA `kr` of the cluster contracts the K-slice [kr*H/2, (kr+1)*H/2) of those rows against the same slice of
//     h_{t-1} for all (<= 64) sequences, so per step a CTA streams 96 x H/2 weights (L2-resident: W_hh is read
//     T times) and only 64 x H/2 of the hidden state.
//   * operands: A = [h_hi ; h_lo] (64 + 64 rows) lives in TENSOR MEMORY (TMA loads h and h_lo - the latter written
//     by the finalising threads of the previous step - and one thread of warp 20 hands the landed tile to the tensor
//     core's copy engine, tcgen05.cp shared -> tensor memory; until round 2 the converter warps moved the rows through
//     their registers with tcgen05.st; TS-mode MMAs read them), B = W_hi then W_lo (96 rows) in shared memory: two M128 N96 K8 MMAs per k-step
//     give h_hi.W + h_lo.W in TMEM lanes 0-63 / 64-127 (all four split products).  Split: hi = the raw FP32 word
//     (the tensor core ignores the low 13 mantissa bits), lo = RN_tf32(x - trunc(x)), so only W_lo is ever
//     written to shared memory by a thread (the first version was bound by the shared-memory pipe: LDS/STS +
//     UMMA operand reads).
//     The tensor core truncates when it adds into its FP32 accumulator, so every 64 k the partial sum is drained
//     (tcgen05.ld) and added round-to-nearest into FP32 registers (same scheme as linear_tc.cu).
//   * end of step: hi + lo rows are combined through shared memory, the two K-slice partials through
//     distributed shared memory (mbarrier handshake, no cluster-wide barrier); each CTA then finalises the 32 units
//     of the cluster for 32 of the 64 sequences: gates, h_t, h_t + residual, with h_{t-1} kept in registers.
//   * h_t goes to y (the API output) and is read back by TMA at the next step.  There is no grid-wide barrier:
//     k-block kb of a K-slice needs exactly the 32 units one cluster produces, so every CTA publishes a step
//     flag (st.release.gpu, one 128-byte line per CTA) and the h producer polls the flags of the clusters it
//     needs (relaxed loads + one acquire fence) before the TMA loads; the W tiles of the next step are prefetched
//     while those flags are awaited.
// Warp roles: 0 = W TMA producer, 1 and 3 = MMA issuers (even / odd accumulator chunks; 1 owns TMEM), 2 = h TMA producer (polls the flags),
// 4-11 = promotion + gates (256 threads), 12-19 = W_lo converters (two groups of 4 warps alternating k-blocks), 20 = A-operand copier.
// Registers are re-divided per warpgroup with setmaxnreg (40 / 120 / 56): the promotion warps hold a 48-float accumulator
// tile next to the gate math (they spilled at the 96 registers a 672-thread kernel gets; gate phase 4.2k -> 1.7k cycles).
// GRU_EXP_* macros are timing experiments (wrong results) kept for the measurements DESIGN.md section 4.2 quotes.
#include <cuda.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "tc_common.cuh"

namespace gait {
namespace grurec {
using namespace tcu;

constexpr int KG = 2;                       // CTAs per cluster = K slices
constexpr int UC = 16;                      // hidden units finalised per CTA
constexpr int UPC = UC * KG;                // hidden units per cluster
constexpr int NB = 3 * UPC;                 // W_hh rows per CTA (UMMA N) = 96
constexpr int SB = 64;                      // sequences (A rows: 64 hi + 64 lo = UMMA M 128)
constexpr int BK = 32;                      // floats per k-block = one 128-byte swizzle span
constexpr int STAGES = 5;
constexpr int H_TILE = SB * BK * 4;         //  8 192 B raw h tile

constexpr int G_TILE = UPC * BK * 4;        //  4 096 B: one gate's rows = one TMA box
constexpr int W_TILE = NB * BK * 4;         // 12 288 B
constexpr int A_TILE = 2 * H_TILE;          // 16 384 B: [h raw = hi (64 rows) | h lo (64 rows)], both loaded by TMA
constexpr int STAGE = A_TILE + 2 * W_TILE;  // 40 960 B: [h hi | h lo | W raw = hi | W lo]
constexpr int PLD = 100;                    // row stride (floats) of the partial-sum buffer: conflict-free float4 rows
constexpr int P_FLOATS = SB * PLD;
constexpr int NBUF = 3;                     // TMEM accumulator ring (3 x 96 columns)
constexpr int TMEM_A = NBUF * NB;           // A operand ring: STAGES x 32 columns, lanes 0-63 h_hi, 64-127 h_lo
constexpr int TMEM_COLS = 512;
static_assert(NBUF * NB + STAGES * BK <= 512, "tensor memory budget");
#ifndef GRU_DRAIN_KB
#define GRU_DRAIN_KB 2
#endif
constexpr int DRAIN_KB = GRU_DRAIN_KB;      // k-blocks per promotion
constexpr int NGRP = 2;                     // converter groups of 4 warps alternating k-blocks (a third group: 768 threads with
                                            // setmaxnreg, 0.475 vs 0.483 ms - not kept; it would also need its own barrier split)
constexpr int NPROM = 256, NCONV = 128 * NGRP;
#ifndef GRU_ACOPY
#define GRU_ACOPY 1                         // 1: the A operand goes shared -> tensor memory by tcgen05.cp (warp 20); 0: through the converters' registers
#endif
constexpr int THREADS = 128 + NPROM + NCONV + (GRU_ACOPY ? 32 : 0);
// Register budget (setmaxnreg, one instruction per warpgroup).  The promotion warps hold a 128 x 48 FP32 tile, four gate
// operands and the previous h per thread; the other roles are small.  setmaxnreg.inc blocks until the CTA's pool (THREADS x
// the kernel's register count R) holds the registers, so the launcher checks R against this budget (regs_ok) instead of
// trusting the compiler: 128 x 40 + 256 x 120 + 256 x 56 + 32 x R <= 672 x R  <=>  R >= 80.
constexpr int REGS_WG0 = 40, REGS_PROM = 120, REGS_CONV = 56;
constexpr int REGS_MIN_KERNEL = GRU_ACOPY ? (128 * REGS_WG0 + NPROM * REGS_PROM + NCONV * REGS_CONV + 639) / 640 : 0;
constexpr int OFF_P = STAGES * STAGE;
constexpr int OFF_BAR = OFF_P + P_FLOATS * 4;       // one partial-sum buffer (P_FREE handshake before it is rewritten)
constexpr int BAR_AREA = 384;               // barriers (8 B each) + the TMEM base slot in the last 8 bytes
constexpr int SMEM = OFF_BAR + BAR_AREA + 1024;  // + barriers + alignment slack
static_assert(SMEM <= 232448, "shared memory budget");

// "Data landed" barriers exist twice per stage, used by alternate fills of the stage (fill k -> barrier k & 1, phase
// k >> 1).  A converter group only handles every NGRP-th k-block; with 5 stages and 2 groups it sees every OTHER fill of a
// stage, and a parity wait on a barrier whose phases it skips would also pass while the fill in between is still
// pending (try_wait.parity cannot tell phase k from phase k - 2).  With the split each group waits on every phase of the
// barriers it uses.  All other barriers are waited on by warps that see each of their phases.
enum : int { B_FULL_W = 0, B_FULL_H = 2 * STAGES, B_CONV = 4 * STAGES, B_EMPTY = 5 * STAGES, B_ACC_FULL = 6 * STAGES,
             B_ACC_EMPTY = 6 * STAGES + NBUF, B_P_READY = 6 * STAGES + 2 * NBUF, B_P_FREE = 6 * STAGES + 2 * NBUF + 1, B_COUNT = 6 * STAGES + 2 * NBUF + 2 };
static_assert(B_COUNT * 8 <= BAR_AREA - 8, "barrier area");
// index of the landed-barrier and the parity to wait for, for the it-th k-block of the launch
__device__ __forceinline__ int full_slot(int it) { return it % STAGES + STAGES * ((it / STAGES) & 1); }
__device__ __forceinline__ uint32_t full_parity(int it) { return (uint32_t)((it / STAGES) >> 1) & 1u; }
constexpr int FLAG_STRIDE = 32;             // words between the step flags of consecutive CTAs (one 128-byte line each)
static_assert(KG == 2, "flag polling reads the two step flags of a cluster");
static_assert(UPC == BK, "one k-block of h = the units of exactly one cluster (flag indexing)");

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared memory -> tensor memory copy of a 128-row x 256-bit tile (8 FP32 columns per lane), source given by a UMMA
// shared-memory matrix descriptor; asynchronous, tracked by the tcgen05.commit the same thread issues after it
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// debug timestamps of CTA 0: trace[step*8 + i] (step < 32) and trace[256 + kb*8 + i] for the k-blocks of step 2
#define GRU_TRACE_STEP(i) do { if (TRACE && trace && blockIdx.x == 0 && step < 32) trace[step * 8 + (i)] = clock64(); } while (0)
#define GRU_TRACE_KB(i) do { if (TRACE && trace && blockIdx.x == 0 && step == 2 && kb < 64) trace[256 + kb * 8 + (i)] = clock64(); } while (0)

// TRACE = true is the instrumented build (gait_debug_gru_trace); the product launches TRACE = false
template <bool TRACE>
__global__ void __launch_bounds__(THREADS, 1)
gru_recurrent_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmY,
                     const __grid_constant__ CUtensorMap tmH0, const __grid_constant__ CUtensorMap tmL, float* hlo,
                     const float* __restrict__ gi,
                     const float* __restrict__ b_hh, const float* __restrict__ h0, float* y, int64_t ldy,
                     const float* __restrict__ resid, int64_t ldres, float* __restrict__ out, int64_t ldout,
                     float* __restrict__ hn, int S, int T, int H, int reverse, unsigned* flags,
                     unsigned long long* trace) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + OFF_BAR;
    auto BAR = [&](int i) { return bars + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + OFF_BAR + BAR_AREA - 8);
    float* P = reinterpret_cast<float*>(gbase + OFF_P);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t kr = cluster_ctarank();
    const int u0 = (blockIdx.x / KG) * UPC;            // first hidden unit of this cluster
    const int KS = H / KG, k0 = (int)kr * KS, NKB = KS / BK;
    const int nchunks = (NKB + DRAIN_KB - 1) / DRAIN_KB;
    const int first_gemm = h0 ? 0 : 1;                 // h_{-1} = 0: step 0 needs no GEMM

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            for (int j = 0; j < 2; ++j) {
                mbar_init(BAR(B_FULL_W + s + j * STAGES), 3);       // one 32-row box per gate, each issued by its own lane
                mbar_init(BAR(B_FULL_H + s + j * STAGES), 4);       // h and h_lo, two 32-sequence boxes each
            }
            mbar_init(BAR(B_CONV + s), 4 + GRU_ACOPY);    // one arrival per warp of the converter group that owns the k-block (+ the A copy's commit)
            mbar_init(BAR(B_EMPTY + s), 1);
        }
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(BAR(B_ACC_FULL + b), 1);
            mbar_init(BAR(B_ACC_EMPTY + b), NPROM / 32); // one arrival per promotion warp
        }
        mbar_init(BAR(B_P_READY), KG);
        mbar_init(BAR(B_P_FREE), KG);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync_all();                                // peers' barriers exist before anyone arrives on them remotely
    const uint32_t tmem_d = *tmem_slot;

    if (warp < 4) {
    // one setmaxnreg per warpgroup (all four warps execute the same instruction); the roles branch below it
#if GRU_ACOPY
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_WG0));
#endif
    if (warp == 0) {
        // ------------------------------------------------------------ W_hh tile producer (independent of h)
        // TMA issue laws measured on this machine (scripts/microbench/kblock_pipe.cu, tma_issue.cu): a TMA warp
        // instruction occupies its warp for ~450 cycles (+ ~50 per extra active lane), the rows of one box are fetched
        // one after the other (~20 cycles each), while boxes issued by different lanes proceed in parallel.
        // So: 32-row boxes (one per gate), one lane per box, two k-blocks (six lanes) per instruction.
        int it = 0;
        for (int step = first_gemm; step < T; ++step) {
            for (int kb0 = 0; kb0 < NKB; kb0 += 2) {
                const int kb = kb0 + lane / 3, g = lane % 3;
                if (lane < 6 && kb < NKB) {
                    const int my = it + lane / 3, s = my % STAGES;
                    const uint32_t ph = (my / STAGES) & 1;
                    mbar_wait(BAR(B_EMPTY + s), ph ^ 1);
                    if (g == 0) GRU_TRACE_KB(0);
                    const uint32_t fb = BAR(B_FULL_W + full_slot(my));
#ifdef GRU_EXP_WBULK   // timing experiment only (wrong results): the W tile of a k-block as ONE contiguous 12 KB bulk copy
                    if (g == 0) {
                        mbar_arrive_expect_tx(fb, W_TILE);
                        bulk_g2s(base + s * STAGE + A_TILE, gi + (((size_t)blockIdx.x * NKB + kb) % 2000) * (W_TILE / 4), W_TILE, fb);
                    } else mbar_arrive(fb);
#else
                    mbar_arrive_expect_tx(fb, G_TILE);
                    tma_load_2d(base + s * STAGE + A_TILE + g * G_TILE, &tmW, k0 + kb * BK, g * H + u0, fb);
#endif
                }
                __syncwarp();
                it += min(2, NKB - kb0);
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------ h_{t-1} tile producer
        int it = 0;
        for (int step = first_gemm; step < T; ++step) {
            const int t = reverse ? (T - 1 - step) : step;
            const int tp = reverse ? t + 1 : t - 1;
            if (step > 0) {
                // dataflow barrier: k-block kb of this K-slice = the 32 hidden units of cluster k0/32 + kb, whose KG CTAs
                // each publish flags[cta] = number of steps completed.  Lanes poll different k-blocks in parallel.
                if (lane == 0) GRU_TRACE_STEP(6);
                for (int kb = lane; kb < NKB; kb += 32) {
                    // relaxed polls, one acquire fence at the end; every CTA's flag has its own 128-byte line (the two
                    // release stores of a cluster and the polls of 64 readers otherwise meet in one L2 sector)
                    const unsigned* fl = flags + (size_t)(k0 / UPC + kb) * KG * FLAG_STRIDE;
                    SpinGuard guard;
                    for (;;) {
                        unsigned f0, f1;
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(f0) : "l"(fl) : "memory");
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(f1) : "l"(fl + FLAG_STRIDE) : "memory");
                        if (f0 >= (unsigned)step && f1 >= (unsigned)step) break;
                        guard.tick();
                    }
                }
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                __syncwarp();
                fence_proxy_async();                   // generic-proxy writes of h -> async-proxy (TMA) reads
                if (lane == 0) GRU_TRACE_STEP(0);
                if (TRACE && trace && lane == 0 && step == 4) trace[1024 + blockIdx.x * 2 + 1] = global_timer_ns();   // skew probe, all CTAs
            }
            // eight lanes issue the boxes of two k-blocks at a time: {h (= hi), h_lo} x {sequences 0-31, 32-63}.
            // h_lo of the previous step sits in slot (step-1)&1 of the scratch the finalising threads write.
            const int lslot = (step - 1) & 1;
            for (int kb0 = 0; kb0 < NKB; kb0 += 2) {
                const int kb = kb0 + (lane >> 2), part = (lane >> 1) & 1, half = lane & 1;
                if (lane < 8 && kb < NKB) {
                    const int my = it + (lane >> 2), s = my % STAGES;
                    const uint32_t ph = (my / STAGES) & 1;
                    mbar_wait(BAR(B_EMPTY + s), ph ^ 1);
                    if ((lane & 3) == 0) GRU_TRACE_KB(1);
                    const uint32_t fb = BAR(B_FULL_H + full_slot(my));
                    mbar_arrive_expect_tx(fb, H_TILE / 2);
                    const uint32_t dst = base + s * STAGE + part * H_TILE + half * (H_TILE / 2);
#ifdef GRU_EXP_HBULK   // timing experiment only (wrong results): every 32-sequence h box as one contiguous 4 KB bulk copy
                    bulk_g2s(dst, y + ((size_t)(tp < 0 ? 0 : tp) * NKB * 4 + kb * 4 + (lane & 3)) * (H_TILE / 8), H_TILE / 2, fb);
#else
                    if (part) tma_load_3d(dst, &tmL, k0 + kb * BK, lslot, half * (SB / 2), fb);
                    else if (step == 0) tma_load_3d(dst, &tmH0, k0 + kb * BK, 0, half * (SB / 2), fb);
                    else tma_load_3d(dst, &tmY, k0 + kb * BK, tp, half * (SB / 2), fb);
#endif
                }
                __syncwarp();
                it += min(2, NKB - kb0);
            }
        }
    } else {
        // ------------------------------------------------------------ MMA issuers (warp-uniform, one lane issues)
        // Every warp-specialised role used to touch every k-block, so the k-block period was bounded below by the serial
        // latency chain of ONE warp's loop body (barrier probe -> issue -> commit, ~900 cycles), not by any throughput.
        // Two issuing warps therefore alternate accumulator chunks (the MMAs of one accumulator stay in one warp, in
        // order), and two converter groups alternate k-blocks.
        constexpr uint32_t idesc = umma_idesc_tf32(128, NB);
        const int par = warp == 1 ? 0 : 1;
        int it = 0, ch = 0;
        for (int step = first_gemm; step < T; ++step) {
            for (int kb = 0; kb < NKB; ++kb, ++it) {
                const int cg = ch + kb / DRAIN_KB, buf = cg % NBUF, use = cg / NBUF;
                if ((cg & 1) != par) continue;
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const bool chunk_start = (kb % DRAIN_KB) == 0;
                if (chunk_start && use >= 1) mbar_wait(BAR(B_ACC_EMPTY + buf), (use - 1) & 1);
                if (lane == 0) GRU_TRACE_KB(2);
                mbar_wait(BAR(B_CONV + s), ph);
                if (lane == 0) GRU_TRACE_KB(3);
                tc_fence_after();
                const uint32_t acc = tmem_d + (uint32_t)(buf * NB);
                const uint32_t st = base + s * STAGE;
                const uint32_t a = tmem_d + (uint32_t)(TMEM_A + s * BK);
                const uint64_t b_hi = make_sdesc_sw128(st + A_TILE), b_lo = make_sdesc_sw128(st + A_TILE + W_TILE);
                const bool last_of_chunk = (kb % DRAIN_KB) == DRAIN_KB - 1 || kb == NKB - 1;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {
                        const uint64_t adv = (uint64_t)(k * 32 >> 4);
                        umma_tf32_ts(acc, a + 8 * k, b_hi + adv, idesc, (chunk_start && k == 0) ? 0u : 1u);
#ifndef GRU_EXP_HALFMMA   // timing experiment only (wrong results): W_hi products only
                        umma_tf32_ts(acc, a + 8 * k, b_lo + adv, idesc, 1u);
#endif
                    }
                    umma_commit(BAR(B_EMPTY + s));
                    if (last_of_chunk) umma_commit(BAR(B_ACC_FULL + buf));
                    GRU_TRACE_KB(7);
                    if (kb == NKB - 1) GRU_TRACE_STEP(2);
                }
                __syncwarp();
            }
            ch += nchunks;
        }
    }
    } else if (warp >= 12 && warp < 20) {
#if GRU_ACOPY
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CONV));
#endif
        // ------------------------------------------------------------ converters: two groups of 4 warps alternate k-blocks
        // W: lo tile only (the raw tile is the hi operand).  A: each thread moves one row of the stage's [h ; h_lo] tile
        // from shared memory into tensor memory (no arithmetic; TMEM lane = row, so a group needs all four warp quadrants).
        const int grp = (warp - 12) >> 2;
        const int quad = warp & 3;
        const int gt = quad * 32 + lane;                              // thread in the group = A row = TMEM lane
        int it = 0;
        for (int step = first_gemm; step < T; ++step) {
            for (int kb = 0; kb < NKB; ++kb, ++it) {
                if ((it % NGRP) != grp) continue;
                const int s = it % STAGES;
                uint8_t* st = gbase + s * STAGE;
                const float4* w_hi = reinterpret_cast<const float4*>(st + A_TILE) + gt;
                float4* w_lo = reinterpret_cast<float4*>(st + A_TILE + W_TILE) + gt;
                constexpr int NW = W_TILE / 16 / 128;                 // 6
                mbar_wait(BAR(B_FULL_W + full_slot(it)), full_parity(it));
                if (gt == 0) GRU_TRACE_KB(4);
#ifndef GRU_EXP_NOWCONV   // timing experiment only (wrong results): no W_lo conversion (24 KB less shared-memory traffic per k-block)
                float4 v[NW];
#pragma unroll
                for (int i = 0; i < NW; ++i) v[i] = w_hi[i * 128];
#pragma unroll
                for (int i = 0; i < NW; ++i)
                    w_lo[i * 128] = make_float4(tf32_lo(v[i].x), tf32_lo(v[i].y), tf32_lo(v[i].z), tf32_lo(v[i].w));
#else
                (void)w_hi; (void)w_lo;
#endif
#if !GRU_ACOPY
                mbar_wait(BAR(B_FULL_H + full_slot(it)), full_parity(it));
                if (gt == 0) { GRU_TRACE_KB(5); if (kb == 0) GRU_TRACE_STEP(1); }
                {
                    const float4* hrow = reinterpret_cast<const float4*>(st + gt * (BK * 4));
                    uint32_t r[32];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 x = hrow[c ^ (gt & 7)];                    // un-swizzle
                        r[4 * c] = __float_as_uint(x.x); r[4 * c + 1] = __float_as_uint(x.y);
                        r[4 * c + 2] = __float_as_uint(x.z); r[4 * c + 3] = __float_as_uint(x.w);
                    }
                    const uint32_t ta = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)(TMEM_A + s * BK);
                    tmem_st16(ta, r);
                    tmem_st16(ta + 16, r + 16);
                }
                tmem_st_wait();
                tc_fence_before();
#endif
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(B_CONV + s));
                if (gt == 0) GRU_TRACE_KB(6);
            }
        }
    } else if (GRU_ACOPY && warp == 20) {
        // ------------------------------------------------------------ A operand: [h ; h_lo] tile, shared -> tensor memory
        // One thread hands the stage's 128 x 32 tile to the tensor core's copy engine (no LSU / ALU work, no register pass)
        // and commits to the stage's "converted" barrier, next to the W_lo arrivals of the converter group.  The tensor-memory
        // slot is free: its previous readers (the MMAs of k-block it - STAGES) retired before the stage was refilled.
        int it = 0;
        for (int step = first_gemm; step < T; ++step) {
            for (int kb = 0; kb < NKB; ++kb, ++it) {
                const int s = it % STAGES;
                mbar_wait(BAR(B_FULL_H + full_slot(it)), full_parity(it));
                if (lane == 0) { GRU_TRACE_KB(5); if (kb == 0) GRU_TRACE_STEP(1); }
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t a_src = make_sdesc_sw128(base + s * STAGE);
                    const uint32_t a = tmem_d + (uint32_t)(TMEM_A + s * BK);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) tmem_cp_128x256b(a + 8 * k, a_src + (uint64_t)(k * 32 >> 4));
                    umma_commit(BAR(B_CONV + s));
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4 && warp < 12) {
#if GRU_ACOPY
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_PROM));
#endif
        // ------------------------------------------------------------ promotion + gates
        const int pt = threadIdx.x - 128;
        const int q = warp & 3;                            // TMEM lane quadrant of this warp
        const int hf = (warp - 4) >> 2;                    // which 48 of the 96 accumulator columns
        const int prow = (q * 32 + lane) & 63;             // sequence of this thread's TMEM lane (lanes 64+ = lo rows)
        // finalisation: thread -> (sequence, 4 consecutive hidden units).  CTA kr of the cluster finalises all 32 units of
        // the cluster for sequences 32*kr .. 32*kr+31, so every row it reads (gi, resid) or writes (y, out) is a whole
        // 128-byte line that no other CTA touches.
        const int seq = (int)kr * (SB / KG) + (pt >> 3), u4 = pt & 7;
        const int ucol = u4 * 4;                           // column inside a gate's 32-wide block of P
        const int unit = u0 + ucol;
        const bool act = seq < S;
        float4 hp = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act && h0) hp = *reinterpret_cast<const float4*>(h0 + (int64_t)seq * H + unit);
        const float4 br = *reinterpret_cast<const float4*>(b_hh + unit);
        const float4 bz = *reinterpret_cast<const float4*>(b_hh + H + unit);
        const float4 bn = *reinterpret_cast<const float4*>(b_hh + 2 * H + unit);
        int ch = 0, gstep = 0;
        for (int step = 0; step < T; ++step) {
            const int t = reverse ? (T - 1 - step) : step;
            const int64_t f = (int64_t)seq * T + t;
            float4 ar = br, az = bz, an = make_float4(0.f, 0.f, 0.f, 0.f), rs = an;
            if (act) {                                     // issued before the GEMM of this step completes
                const float* g = gi + f * 3 * H + unit;
                ar = f4_add(ar, *reinterpret_cast<const float4*>(g));
                az = f4_add(az, *reinterpret_cast<const float4*>(g + H));
                an = *reinterpret_cast<const float4*>(g + 2 * H);
                if (out) rs = *reinterpret_cast<const float4*>(resid + f * ldres + unit);
            }
            float4 sr = make_float4(0.f, 0.f, 0.f, 0.f), sz = sr, sn = sr;
            if (step >= first_gemm) {
                float acc[NB / 2];
#pragma unroll
                for (int j = 0; j < NB / 2; ++j) acc[j] = 0.f;
                for (int chunk = 0; chunk < nchunks; ++chunk) {
                    const int cg = ch + chunk, buf = cg % NBUF, use = cg / NBUF;
                    mbar_wait(BAR(B_ACC_FULL + buf), use & 1);
                    tc_fence_after();
                    const uint32_t ta = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * NB + hf * (NB / 2));
#pragma unroll
                    for (int c = 0; c < NB / 32; ++c) {
                        uint32_t r[16];
                        tmem_ld16_nowait(ta + c * 16, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[c * 16 + j] += __uint_as_float(r[j]);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + buf));
                }
                ch += nchunks;
                if (pt == 0) GRU_TRACE_STEP(3);
                // hi rows + lo rows -> this CTA's K-slice partial P[seq][96]
                float* Pb = P;
                float4* prow4 = reinterpret_cast<float4*>(Pb + prow * PLD + hf * (NB / 2));
                // every CTA of the cluster has finished reading the previous step's partials (its own and this one's)
                if (gstep >= 1) mbar_wait_cluster(BAR(B_P_FREE), (gstep - 1) & 1);
                if (q >= 2) {
#pragma unroll
                    for (int j = 0; j < NB / 8; ++j) prow4[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
                }
                named_bar_sync(1, NPROM);
                if (q < 2) {
#pragma unroll
                    for (int j = 0; j < NB / 8; ++j) {
                        const float4 o = prow4[j];
                        prow4[j] = make_float4(o.x + acc[4 * j], o.y + acc[4 * j + 1], o.z + acc[4 * j + 2], o.w + acc[4 * j + 3]);
                    }
                }
                named_bar_sync(1, NPROM);
                // tell every CTA of the cluster (incl. this one) that this partial is complete, wait for all
                const uint32_t pbar = BAR(B_P_READY);
                if (pt == 0) {
#pragma unroll
                    for (uint32_t rr = 0; rr < (uint32_t)KG; ++rr) mbar_arrive_remote_release(map_to_rank(pbar, rr));
                }
                mbar_wait_cluster(pbar, gstep & 1);
                if (pt == 0) GRU_TRACE_STEP(4);
                const uint32_t pa = smem_u32(Pb + seq * PLD + ucol);
#pragma unroll
                for (uint32_t rr = 0; rr < (uint32_t)KG; ++rr) {
                    const uint32_t ra = map_to_rank(pa, rr);
                    sr = f4_add(sr, ld_cluster_f4(ra));
                    sz = f4_add(sz, ld_cluster_f4(ra + UPC * 4));
                    sn = f4_add(sn, ld_cluster_f4(ra + 2 * UPC * 4));
                }
                ++gstep;
            }
            float4 h;
            {
                const float r0 = sigmoidf_(ar.x + sr.x), r1 = sigmoidf_(ar.y + sr.y), r2 = sigmoidf_(ar.z + sr.z), r3 = sigmoidf_(ar.w + sr.w);
                const float z0 = sigmoidf_(az.x + sz.x), z1 = sigmoidf_(az.y + sz.y), z2 = sigmoidf_(az.z + sz.z), z3 = sigmoidf_(az.w + sz.w);
                const float n0 = tanhf(an.x + r0 * (sn.x + bn.x)), n1 = tanhf(an.y + r1 * (sn.y + bn.y));
                const float n2 = tanhf(an.z + r2 * (sn.z + bn.z)), n3 = tanhf(an.w + r3 * (sn.w + bn.w));
                h = make_float4((1.f - z0) * n0 + z0 * hp.x, (1.f - z1) * n1 + z1 * hp.y, (1.f - z2) * n2 + z2 * hp.z,
                                (1.f - z3) * n3 + z3 * hp.w);
            }
            hp = h;
            if (pt == 0) GRU_TRACE_STEP(7);
            // what the next step of other CTAs reads (h_t and its lo part) goes out first and is published; the outputs nobody
            // inside the kernel reads (h_t + residual, h_n) are stored after the flag so that they are off the critical path
            if (act) {
                *reinterpret_cast<float4*>(y + f * ldy + unit) = h;
                if (step < T - 1)          // lo half of the next step's A operand (hi = h itself, read truncated)
                    *reinterpret_cast<float4*>(hlo + ((int64_t)seq * 2 + (step & 1)) * H + unit) =
                        make_float4(tf32_lo(h.x), tf32_lo(h.y), tf32_lo(h.z), tf32_lo(h.w));
            }
            if (step < T - 1) {
                fence_proxy_async_global();            // generic-proxy stores of h_t -> async-proxy (TMA) reads by other CTAs
                named_bar_sync(1, NPROM);
                if (pt == 0) {
                    st_release_gpu(flags + (size_t)blockIdx.x * FLAG_STRIDE, (unsigned)(step + 1));
                    if (step >= first_gemm) {           // the partial-sum buffers of this cluster may be rewritten
#pragma unroll
                        for (uint32_t rr = 0; rr < (uint32_t)KG; ++rr) mbar_arrive_remote_release(map_to_rank(BAR(B_P_FREE), rr));
                    }
                }
            }
            if (act) {
                if (out) *reinterpret_cast<float4*>(out + f * ldout + unit) = f4_add(h, rs);
                if (hn && step == T - 1) *reinterpret_cast<float4*>(hn + (int64_t)seq * H + unit) = h;
            }
            if (pt == 0) GRU_TRACE_STEP(5);
            if (TRACE && trace && pt == 0 && step == 3) trace[1024 + blockIdx.x * 2] = global_timer_ns();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS) : "memory");
    }
    cluster_sync_all();                                // nobody leaves while a peer may still read its partials
}

// h0 given by the caller: lo half of step 0's A operand into scratch slot 1 (= (0 - 1) & 1)
__global__ void split_h0_lo_kernel(const float* __restrict__ h0, float* __restrict__ hlo, int64_t S, int64_t H) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= S * H) return;
    const int64_t s = i / H, u = i % H;
    hlo[(s * 2 + 1) * H + u] = tf32_lo(h0[i]);
}

unsigned long long* g_trace = nullptr;

static int max_coresident_ctas() {
    static std::atomic<int> cache[64];                     // per device (0 = not queried yet; stored as value + 1)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64) {
        const int c = cache[dev].load(std::memory_order_relaxed);
        if (c > 0) return c - 1;
    }
    int n_clusters = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * KG);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = KG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaFuncSetAttribute(gru_recurrent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(gru_recurrent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess ||
        cudaOccupancyMaxActiveClusters(&n_clusters, gru_recurrent_kernel<false>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        n_clusters = 0;
    }
    if (REGS_MIN_KERNEL > 0) {   // the register pool must cover the setmaxnreg budget, or the promotion warps would wait forever
        cudaFuncAttributes fa0, fa1;
        if (cudaFuncGetAttributes(&fa0, gru_recurrent_kernel<false>) != cudaSuccess ||
            cudaFuncGetAttributes(&fa1, gru_recurrent_kernel<true>) != cudaSuccess || fa0.numRegs < REGS_MIN_KERNEL ||
            fa1.numRegs < REGS_MIN_KERNEL) {
            cudaGetLastError();
            n_clusters = 0;
        }
    }
    if (dev >= 0 && dev < 64) cache[dev].store(n_clusters * KG + 1, std::memory_order_relaxed);
    return n_clusters * KG;
}

}  // namespace grurec

bool gru_recurrent_eligible(const float* gi, const float* W_hh, const float* h0, const float* y, int64_t ldy,
                            const float* resid, int64_t ldres, const float* out, int64_t ldout, int64_t S, int64_t T,
                            int64_t H) {
    using namespace grurec;
    if (S < 1 || S > SB || T < 1 || T > 32768) return false;
    if (H % (BK * KG) != 0 || H % UPC != 0) return false;
    if (!aligned16(gi) || !aligned16(W_hh) || !aligned16(y) || (ldy & 3)) return false;
    if (h0 && !aligned16(h0)) return false;
    if (out && (!aligned16(out) || !aligned16(resid) || (ldout & 3) || (ldres & 3))) return false;
    return H / UC <= max_coresident_ctas();
}

int gru_recurrent_launch(const float* gi, const float* W_hh, const float* b_hh, const float* h0, float* y, int64_t ldy,
                         const float* resid, int64_t ldres, float* out, int64_t ldout, float* hn, int64_t S, int64_t T,
                         int64_t H, int reverse, unsigned int* counter, float* hlo, cudaStream_t stream) {
    using namespace grurec;
    GAIT_REQUIRE(aligned16(b_hh) && (!hn || aligned16(hn)), "gru(persistent): b_hh / hn must be 16-byte aligned");
    GAIT_REQUIRE(aligned16(hlo), "gru(persistent): scratch must be 16-byte aligned");
    CUtensorMap tmW, tmY, tmH0, tmL;
    // h_lo scratch (S, 2, H): slot = step parity
    GAIT_TRY(make_tensor_map_3d_f32(&tmL, hlo, (uint64_t)H, 2, (uint64_t)S, (uint64_t)H * 4, (uint64_t)(2 * H) * 4, BK, 1, SB / 2, true));
    if (h0) {
        split_h0_lo_kernel<<<(unsigned)ceil_div(S * H, 256), 256, 0, stream>>>(h0, hlo, S, H);
        GAIT_TRY(check_launch("gru(h0 split)"));
    }
    GAIT_TRY(make_tensor_map_2d(&tmW, 4, W_hh, (uint64_t)H, (uint64_t)(3 * H), (uint64_t)H * 4, BK, UPC, true));
    GAIT_TRY(make_tensor_map_3d_f32(&tmY, y, (uint64_t)H, (uint64_t)T, (uint64_t)S, (uint64_t)ldy * 4, (uint64_t)(T * ldy) * 4,
                                    BK, 1, SB / 2, true));
    if (h0) GAIT_TRY(make_tensor_map_3d_f32(&tmH0, h0, (uint64_t)H, 1, (uint64_t)S, (uint64_t)H * 4, (uint64_t)H * 4, BK, 1, SB / 2, true));
    else tmH0 = tmY;
    GAIT_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int) * (size_t)(H / UC) * FLAG_STRIDE, stream));   // per-CTA step flags
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(H / UC));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = KG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.attrs = at;
    // The kernel spin-waits on flags written by other CTAs, so all H/16 CTAs must be resident at once.  A cooperative launch
    // guarantees that (the runtime refuses it otherwise, or delays it until the device can hold the whole grid).  If the
    // cooperative + cluster launch is refused, the caller falls back to the per-step path (gru.cu), which has no residency
    // requirement; the refusal is not remembered (it may be transient).  GAITB200_GRU_COOP=0 - or an attached Nsight Compute,
    // which cannot replay cooperative launches - selects a plain cluster launch guarded by cudaOccupancyMaxActiveClusters;
    // that mode assumes the device is otherwise idle (profiling, smoke test).
    static int coop = -1;
    if (coop < 0) {
        const char* e = getenv("GAITB200_GRU_COOP");
        const bool profiler = getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("NV_NSIGHT_INJECTION_TRANSPORT_TYPE") ||
                              getenv("CUDA_INJECTION64_PATH") || getenv("NV_NSIGHT_INJECTION_PORT_BASE");
        coop = e ? (atoi(e) == 0 ? 0 : 1) : (profiler ? 0 : 1);
    }
    cfg.numAttrs = coop ? 2 : 1;
    auto kernel = g_trace ? gru_recurrent_kernel<true> : gru_recurrent_kernel<false>;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, tmW, tmY, tmH0, tmL, hlo, gi, b_hh, h0, y, ldy, resid, ldres, out,
                                       ldout, hn, (int)S, (int)T, (int)H, reverse, counter, g_trace);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("gru(persistent): %s launch refused: %s", coop ? "cooperative cluster" : "cluster", cudaGetErrorString(e));
        return GAIT_GRU_RETRY_PER_STEP;
    }
    return check_launch("gru(persistent recurrent)");
}

}  // namespace gait

// Debug hook: device buffer of 512 uint64 that receives CTA 0's per-step and per-k-block (step 2) clock64 stamps
// of subsequent persistent-GRU launches; NULL disables.
extern "C" int gait_debug_gru_trace(unsigned long long* device_buffer) {
    gait::grurec::g_trace = device_buffer;
    return GAIT_OK;
}
