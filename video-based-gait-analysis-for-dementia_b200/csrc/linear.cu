// gait_linear: C = A . W^T + bias + Cin, FP32.
//
// This file holds the SIMT FP32 path (exact FP32 FMA accumulation): a register-blocked,
// double-buffered shared-memory GEMM for "TN" operands (both A and W are K-major, which is how
// torch.nn.Linear / nn.GRU store weights and activations).  It serves the shapes the tensor-core
// path (linear_tc.cu: tcgen05 split-TF32) does not take: small M, odd K, unaligned strides.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace gait {

template <int BM, int BN, int BK, int TM, int TN, bool VEC>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_tn_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W, int64_t ldw,
                const float* __restrict__ bias, const float* Cin, int64_t ldcin, float* C, int64_t ldc,
                int M, int N, int K) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int KC = BK / 4;                       // float4 chunks per tile row
    constexpr int A_CH = (BM * KC) / NT;             // chunks per thread
    constexpr int W_CH = (BN * KC) / NT;
    static_assert((BM * KC) % NT == 0 && (BN * KC) % NT == 0, "tile/threads mismatch");
    static_assert(TM % 4 == 0 && TN % 4 == 0, "microtile must be a multiple of 4");
    constexpr int GM = TM / 4, GN = TN / 4;          // float4 groups per microtile side
    constexpr int SM_STRIDE = BM / GM, SN_STRIDE = BN / GN;

    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Ws[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[A_CH], rw[W_CH];

    auto load_tile = [&](const float* __restrict__ P, int64_t ld, int rows, int r0, int k0, int c) -> float4 {
        const int row = c / KC, kc = c % KC;
        const int gr = r0 + row, gk = k0 + kc * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < rows) {
            const float* p = P + (int64_t)gr * ld + gk;
            if (VEC && gk + 3 < K) {
                v = *reinterpret_cast<const float4*>(p);
            } else {
                if (gk + 0 < K) v.x = p[0];
                if (gk + 1 < K) v.y = p[1];
                if (gk + 2 < K) v.z = p[2];
                if (gk + 3 < K) v.w = p[3];
            }
        }
        return v;
    };
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_CH; ++i) ra[i] = load_tile(A, lda, M, m0, k0, tid + i * NT);
#pragma unroll
        for (int i = 0; i < W_CH; ++i) rw[i] = load_tile(W, ldw, N, n0, k0, tid + i * NT);
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_CH; ++i) {
            const int c = tid + i * NT, row = c / KC, kc = c % KC;
            As[buf][kc * 4 + 0][row] = ra[i].x; As[buf][kc * 4 + 1][row] = ra[i].y;
            As[buf][kc * 4 + 2][row] = ra[i].z; As[buf][kc * 4 + 3][row] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < W_CH; ++i) {
            const int c = tid + i * NT, row = c / KC, kc = c % KC;
            Ws[buf][kc * 4 + 0][row] = rw[i].x; Ws[buf][kc * 4 + 1][row] = rw[i].y;
            Ws[buf][kc * 4 + 2][row] = rw[i].z; Ws[buf][kc * 4 + 3][row] = rw[i].w;
        }
    };

    const int nk = (K + BK - 1) / BK;
    fetch(0);
    stash(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) fetch((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int g = 0; g < GM; ++g) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][k][g * SM_STRIDE + ty * 4]);
                a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int g = 0; g < GN; ++g) {
                float4 v = *reinterpret_cast<const float4*>(&Ws[buf][k][g * SN_STRIDE + tx * 4]);
                b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            stash(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue: bias + optional residual input, guarded scalar or float4 stores
    const bool vec_out = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0) &&
                         (Cin == nullptr || (((ldcin & 3) == 0) && ((reinterpret_cast<uintptr_t>(Cin) & 15u) == 0)));
#pragma unroll
    for (int gi = 0; gi < GM; ++gi)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int row = m0 + gi * SM_STRIDE + ty * 4 + ii;
            if (row >= M) continue;
#pragma unroll
            for (int gj = 0; gj < GN; ++gj) {
                const int col = n0 + gj * SN_STRIDE + tx * 4;
                if (col >= N) continue;
                float v[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) v[jj] = acc[gi * 4 + ii][gj * 4 + jj];
                if (vec_out && col + 3 < N) {
                    if (bias) {
                        const float4 bv = *reinterpret_cast<const float4*>(bias + col);
                        v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w;
                    }
                    if (Cin) {
                        const float4 cv = *reinterpret_cast<const float4*>(Cin + (int64_t)row * ldcin + col);
                        v[0] += cv.x; v[1] += cv.y; v[2] += cv.z; v[3] += cv.w;
                    }
                    *reinterpret_cast<float4*>(C + (int64_t)row * ldc + col) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        if (col + jj < N) {
                            float o = v[jj];
                            if (bias) o += bias[col + jj];
                            if (Cin) o += Cin[(int64_t)row * ldcin + col + jj];
                            C[(int64_t)row * ldc + col + jj] = o;
                        }
                    }
                }
            }
        }
}

template <int BM, int BN, int BK, int TM, int TN>
static int launch_sgemm(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                        const float* Cin, int64_t ldcin, float* C, int64_t ldc, int M, int N, int K,
                        bool vec, cudaStream_t stream) {
    dim3 grid((unsigned)ceil_div(N, BN), (unsigned)ceil_div(M, BM));
    constexpr int NT = (BM / TM) * (BN / TN);
    if (vec)
        sgemm_tn_kernel<BM, BN, BK, TM, TN, true><<<grid, NT, 0, stream>>>(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K);
    else
        sgemm_tn_kernel<BM, BN, BK, TM, TN, false><<<grid, NT, 0, stream>>>(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K);
    return check_launch("linear(simt)");
}

int linear_simt_launch(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                       const float* Cin, int64_t ldcin, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                       cudaStream_t stream) {
    const bool vec = ((lda & 3) == 0) && ((ldw & 3) == 0) && aligned16(A) && aligned16(W);
    const int64_t tiles128 = ceil_div(M, 128) * ceil_div(N, 128);
    if (tiles128 >= 148)
        return launch_sgemm<128, 128, 16, 8, 8>(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, (int)M, (int)N, (int)K, vec, stream);
    return launch_sgemm<64, 64, 16, 4, 4>(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, (int)M, (int)N, (int)K, vec, stream);
}

int linear_launch(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                  const float* Cin, int64_t ldcin, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                  cudaStream_t stream) {
    if (M == 0 || N == 0) return GAIT_OK;
    GAIT_REQUIRE(A && W && C, "linear: null pointer");
    GAIT_REQUIRE(M > 0 && N > 0 && K > 0, "linear: sizes must be positive");
    GAIT_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "linear: size exceeds int32");
    GAIT_REQUIRE(lda >= K && ldw >= K && ldc >= N && (Cin == nullptr || ldcin >= N), "linear: stride smaller than row");
    if (linear_path() != 1 && linear_tc_eligible(A, lda, W, ldw, M, N, K))
        return linear_tc_launch(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, 1, 0, stream);
    return linear_simt_launch(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, stream);
}

// 0 = choose by shape (default), 1 = force the SIMT FP32 kernel (GAITB200_LINEAR=simt; A/B testing)
int linear_path() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("GAITB200_LINEAR");
        mode = (e && strcmp(e, "simt") == 0) ? 1 : 0;
    }
    return mode;
}

}  // namespace gait

extern "C" int gait_linear(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                           const float* Cin, int64_t ldcin, float* C, int64_t ldc, int64_t M, int64_t N,
                           int64_t K, gait_stream_t stream) {
    return gait::linear_launch(A, lda, W, ldw, bias, Cin, ldcin, C, ldc, M, N, K, gait::as_stream(stream));
}

namespace gait { void linear_tc_set_trace(unsigned long long* p); }
// Debug hook (not part of the reference-facing surface): device buffer of 64*4 uint64 receiving pipeline
// timestamps of CTA 0 of the next tensor-core GEMM launches; NULL disables.
extern "C" int gait_debug_linear_trace(unsigned long long* device_buffer) {
    gait::linear_tc_set_trace(device_buffer);
    return GAIT_OK;
}
