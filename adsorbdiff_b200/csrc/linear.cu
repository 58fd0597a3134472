// K2 (fp32 SIMT variant): C[M][N] = act(A[M][K] . W[N][K]^T + bias).
// Exact-fp32 dense contraction used for every node-wise linear of PaiNN (x_proj, vec_proj,
// xvec_proj, output heads; reference: models/painn/painn_denoising.py:508-512, 580-587,
// 667-676).  128x128x16 CTA tile, 8x8 register tile per thread, double-buffered shared memory,
// A and W both K-contiguous (torch.nn.Linear layout), so tiles are transposed on the way in.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, LT = 256, PAD = 4;

__device__ __forceinline__ float apply_act(float v, int act) { return act == ADK_ACT_SSILU ? adk::ssilu(v) : v; }

__global__ void __launch_bounds__(LT, 2) linear_kernel(const float* __restrict__ A, int64_t lda,
                                                    const float* __restrict__ W, const float* __restrict__ bias,
                                                    int M, int N, int K, int act, float* __restrict__ C,
                                                    int64_t ldc) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    // global -> register staging: each thread moves 2 float4 of A and 2 of W per k-tile
    const int lrow = tid >> 2, lk = (tid & 3) * 4;  // rows lrow and lrow+64, k offset lk
    float4 ra[2], rb[2];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int m = m0 + lrow + 64 * h, n = n0 + lrow + 64 * h;
            ra[h] = (m < M) ? *reinterpret_cast<const float4*>(A + (int64_t)m * lda + k0 + lk)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
            rb[h] = (n < N) ? *reinterpret_cast<const float4*>(W + (int64_t)n * K + k0 + lk)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int r = lrow + 64 * h;
            As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y;
            As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
            Bs[buf][lk + 0][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y;
            Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    const int nk = K / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4 + 64]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4 + 64]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 4 + (i & 3) + 64 * (i >> 2);
        if (m >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int n = n0 + tx * 4 + 64 * jh;
            if (n >= N) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float t = acc[i][jh * 4 + j];
                if (bias && n + j < N) t += bias[n + j];
                v[j] = apply_act(t, act);
            }
            float* dst = C + (int64_t)m * ldc + n;
            if (vec_ok && n + 3 < N) {
                *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < N) dst[j] = v[j];
            }
        }
    }
}

// N <= 4 outputs (the 1- and 2-column linears that end the output heads): one warp per row, a 128x128 tile
// would waste 98 % of its work.  Summation: per-lane partials in K order, then a shuffle tree.
__global__ void linear_small_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W,
                                    const float* __restrict__ bias, int M, int N, int K, int act,
                                    float* __restrict__ C, int64_t ldc) {
    const int row = blockIdx.x * (blockDim.x >> 5) + adk::warp_id();
    if (row >= M) return;
    const int lane = adk::lane_id();
    const float* a = A + (int64_t)row * lda;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane * 4; k < K; k += 128) {
        const float4 av = *reinterpret_cast<const float4*>(a + k);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            if (n < N) {
                const float4 wv = *reinterpret_cast<const float4*>(W + (int64_t)n * K + k);
                acc[n] = fmaf(av.x, wv.x, fmaf(av.y, wv.y, fmaf(av.z, wv.z, fmaf(av.w, wv.w, acc[n]))));
            }
        }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        if (n < N) {
            const float v = adk::warp_sum(acc[n]) + (bias ? bias[n] : 0.f);
            if (lane == 0) C[(int64_t)row * ldc + n] = apply_act(v, act);
        }
    }
}

}  // namespace

extern "C" int adk_linear(const float* A, int64_t lda, const float* W, const float* bias, int M, int N, int K,
                          int act, float* C, int64_t ldc, void* stream) {
    if (!A || !W || !C || M <= 0 || N <= 0 || K <= 0 || (K % BK) != 0 || (lda & 3) != 0 ||
        (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
        return ADK_EINVAL;
    if (N <= 4 && (K & 3) == 0) {
        linear_small_kernel<<<(M + 7) / 8, 256, 0, adk::as_stream(stream)>>>(A, lda, W, bias, M, N, K, act, C, ldc);
        ADK_LAUNCH_CHECK();
        return 0;
    }
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    linear_kernel<<<grid, LT, 0, adk::as_stream(stream)>>>(A, lda, W, bias, M, N, K, act, C, ldc);
    ADK_LAUNCH_CHECK();
    return 0;
}

int adk_linear_set_attrs() { return 0; }
