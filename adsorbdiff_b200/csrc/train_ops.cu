// Operand preparation for the tensor-core GEMMs of the TRAINING step.
//
// The sampling path measures its fp16x2 prescales once (PaiNN.calibrate) and passes them by value.  Gradients have no
// stable range, so here every GEMM operand gets its prescale on the device, per call, with no host round trip:
//   adk_amax_scale          rec = {s, 1/s}, s = the power of two that puts max|x| into [target/2, target]
//   adk_split_f16_dev       fp32 [M][K]  -> fp16x2 planes [2][plane_rows][K]       scaled by rec[0], pad rows zeroed
//   adk_split_f16_t_dev     fp32 [M][C]  -> planes of the TRANSPOSE [2][plane_rows][Kp] (Kp >= M), pads zeroed
// and adk_linear_tc_dev (csrc/linear_tc.cu) undoes the two scales in its epilogue from the same records.  Together
// they give  Y = X W^T,  dX = dY W,  dW = dY^T X  (torch.nn.Linear forward / backward, reference:
// models/painn/painn_denoising.py:508-512, 580-587, 667-676 under torch autograd) on tcgen05 at fp32 parity.
// Powers of two only: scaling and unscaling are exact.
#include "common.cuh"

namespace {

// scratch[0] = running max of |x| as float bits (non-negative floats order like unsigned ints), scratch[1] = blocks done
__global__ void amax_scale_kernel(const float* __restrict__ src, int64_t n, float target, float* __restrict__ rec,
                                  unsigned int* __restrict__ scratch) {
    float m = 0.f;
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) m = fmaxf(m, fabsf(src[(n4 << 2) + threadIdx.x]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(ADK_FULL_MASK, m, o));
    __shared__ float s_m[8];
    __shared__ bool s_last;
    if (adk::lane_id() == 0) s_m[adk::warp_id()] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s_m[w]);
        if (!(m <= 3.0e38f)) m = 3.0e38f;   // inf / nan: the split will raise the overflow bit
        atomicMax(&scratch[0], __float_as_uint(m));
        __threadfence();
        s_last = atomicAdd(&scratch[1], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        const float amax = __uint_as_float(atomicMax(&scratch[0], 0u));
        float s = 1.0f;
        if (amax > 0.f) {
            int e;
            frexpf(target / amax, &e);      // target / amax = f * 2^e, f in [0.5, 1)  ->  2^(e-1) <= target / amax
            e = max(-100, min(100, e - 1));
            s = ldexpf(1.0f, e);
        }
        rec[0] = s;
        rec[1] = 1.0f / s;
        scratch[0] = 0u;                      // ready for the next call on this scratch
        scratch[1] = 0u;
    }
}

__device__ __forceinline__ void split4(float4 v, float scale, uint2& hi2, uint2& lo2, bool& overflow) {
    __align__(8) __half hi[4];
    __align__(8) __half lo[4];
    adk::split_f16x2(v.x, scale, hi[0], lo[0], overflow);
    adk::split_f16x2(v.y, scale, hi[1], lo[1], overflow);
    adk::split_f16x2(v.z, scale, hi[2], lo[2], overflow);
    adk::split_f16x2(v.w, scale, hi[3], lo[3], overflow);
    hi2 = *reinterpret_cast<const uint2*>(hi);
    lo2 = *reinterpret_cast<const uint2*>(lo);
}

// one thread per 4 consecutive k of one (possibly padding) row
__global__ void split_dev_kernel(const float* __restrict__ src, int64_t ld, int M, int K, const float* __restrict__ rec,
                                 __half* __restrict__ dst, int64_t plane_rows, uint32_t* status) {
    const int K4 = K >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= plane_rows * K4) return;
    const int64_t m = idx / K4;
    const int k4 = (int)(idx - m * K4);
    uint2 hi = make_uint2(0u, 0u), lo = make_uint2(0u, 0u);
    bool overflow = false;
    if (m < M) split4(*reinterpret_cast<const float4*>(src + m * ld + 4 * k4), rec[0], hi, lo, overflow);
    *reinterpret_cast<uint2*>(dst + m * K + 4 * k4) = hi;
    *reinterpret_cast<uint2*>(dst + plane_rows * K + m * K + 4 * k4) = lo;
    if (overflow && status) atomicOr(status, ADK_STATUS_F16_OVERFLOW);
}

// dst[c][m] = split(src[m][c]) through a 32 x 33 tile; grid (Kp / 32, plane_rows / 32), 256 threads
__global__ void split_t_dev_kernel(const float* __restrict__ src, int64_t ld, int M, int C, const float* __restrict__ rec,
                                   __half* __restrict__ dst, int64_t plane_rows, int64_t Kp, uint32_t* status) {
    __shared__ float tile[32][33];
    const int m0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 rows per pass
    for (int r = ty; r < 32; r += 8) {
        const int m = m0 + r, c = c0 + tx;
        tile[r][tx] = (m < M && c < C) ? src[(int64_t)m * ld + c] : 0.f;
    }
    __syncthreads();
    const float scale = rec[0];
    bool overflow = false;
    for (int r = ty; r < 32; r += 8) {
        const int64_t c = c0 + r, m = m0 + tx;
        if (c < plane_rows && m < Kp) {
            __half h, l;
            adk::split_f16x2(tile[tx][r], scale, h, l, overflow);
            dst[c * Kp + m] = h;
            dst[plane_rows * Kp + c * Kp + m] = l;
        }
    }
    if (overflow && status) atomicOr(status, ADK_STATUS_F16_OVERFLOW);
}

}  // namespace

extern "C" int adk_amax_scale(const float* src, int64_t n, float target, float* rec, uint32_t* scratch, void* stream) {
    if (!src || !rec || !scratch || n <= 0 || !(target > 0.f) || (reinterpret_cast<uintptr_t>(src) & 15)) return ADK_EINVAL;
    int64_t blocks = (n / 4 + 255) / 256;
    blocks = blocks < 1 ? 1 : (blocks > 592 ? 592 : blocks);
    amax_scale_kernel<<<(unsigned)blocks, 256, 0, adk::as_stream(stream)>>>(src, n, target, rec, scratch);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_split_f16_dev(const float* src, int64_t ld, int M, int K, const float* rec, void* dst,
                                 int64_t plane_rows, uint32_t* status, void* stream) {
    if (!src || !dst || !rec || M <= 0 || K <= 0 || (K & 3) || (ld & 3) || plane_rows < M) return ADK_EINVAL;
    const int64_t n = plane_rows * (K >> 2);
    split_dev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, adk::as_stream(stream)>>>(
        src, ld, M, K, rec, reinterpret_cast<__half*>(dst), plane_rows, status);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_split_f16_t_dev(const float* src, int64_t ld, int M, int C, const float* rec, void* dst,
                                   int64_t plane_rows, int64_t Kp, uint32_t* status, void* stream) {
    if (!src || !dst || !rec || M <= 0 || C <= 0 || plane_rows < C || Kp < M || ld < C) return ADK_EINVAL;
    dim3 grid((unsigned)((Kp + 31) / 32), (unsigned)((plane_rows + 31) / 32));
    split_t_dev_kernel<<<grid, 256, 0, adk::as_stream(stream)>>>(src, ld, M, C, rec, reinterpret_cast<__half*>(dst),
                                                                 plane_rows, Kp, status);
    ADK_LAUNCH_CHECK();
    return 0;
}

// ---- backward of the update block's element-wise halves (forward: adk_update_prep / adk_update_gate, csrc/node_ops.cu;
// reference: PaiNNUpdate.forward painn_denoising.py:601-623 under torch autograd) ------------------------------------
namespace {

// dot = sum_c v1 v2 / sqrt(F), cat = [x | sqrt(sum_c v2^2 + 1e-8)]  =>
//   g_x = g_cat[:, :F];  g_v1[c] = g_dot v2[c] / sqrt(F);  g_v2[c] = g_dot v1[c] / sqrt(F) + g_cat[:, F:] v2[c] / norm
__global__ void update_prep_bwd_kernel(const float* __restrict__ vp, const float* __restrict__ g_dot,
                                       const float* __restrict__ g_cat, int N, int F, float inv_sqrt_h,
                                       float* __restrict__ g_x, float* __restrict__ g_vp) {
    const int F4 = F >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * F4) return;
    const int n = (int)(idx / F4), f = ((int)(idx - (int64_t)n * F4)) << 2;
    const float* r = vp + (int64_t)n * 3 * 2 * F + f;
    float4 v1[3], v2[3];
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        v1[c] = *reinterpret_cast<const float4*>(r + c * 2 * F);
        v2[c] = *reinterpret_cast<const float4*>(r + c * 2 * F + F);
        q.x += v2[c].x * v2[c].x; q.y += v2[c].y * v2[c].y; q.z += v2[c].z * v2[c].z; q.w += v2[c].w * v2[c].w;
    }
    const float4 gd = *reinterpret_cast<const float4*>(g_dot + (int64_t)n * F + f);
    const float4 gc0 = *reinterpret_cast<const float4*>(g_cat + (int64_t)n * 2 * F + f);
    const float4 gc1 = *reinterpret_cast<const float4*>(g_cat + (int64_t)n * 2 * F + F + f);
    *reinterpret_cast<float4*>(g_x + (int64_t)n * F + f) = gc0;
    const float4 gs = make_float4(gd.x * inv_sqrt_h, gd.y * inv_sqrt_h, gd.z * inv_sqrt_h, gd.w * inv_sqrt_h);
    const float4 gn = make_float4(gc1.x / sqrtf(q.x + 1e-8f), gc1.y / sqrtf(q.y + 1e-8f), gc1.z / sqrtf(q.z + 1e-8f),
                                  gc1.w / sqrtf(q.w + 1e-8f));
    float* o = g_vp + (int64_t)n * 3 * 2 * F + f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        *reinterpret_cast<float4*>(o + c * 2 * F) = make_float4(gs.x * v2[c].x, gs.y * v2[c].y, gs.z * v2[c].z, gs.w * v2[c].w);
        *reinterpret_cast<float4*>(o + c * 2 * F + F) =
            make_float4(gs.x * v1[c].x + gn.x * v2[c].x, gs.y * v1[c].y + gn.y * v2[c].y, gs.z * v1[c].z + gn.z * v2[c].z,
                        gs.w * v1[c].w + gn.w * v2[c].w);
    }
}

// x' = (x + (a + b dot) / sqrt(2)) s, vec' = vec + c v1  (s == 0: no multiply)  =>
//   g_x = g_x' s;  g_a = g_x' s / sqrt(2);  g_b = g_a dot;  g_dot = g_a b;  g_c = sum_xyz g_vec' v1;  g_vec = g_vec';
//   g_v1[xyz] = c g_vec'[xyz], g_v2 = 0  (written as this op's own g_vp; autograd adds the prep half's)
__global__ void update_gate_bwd_kernel(const float* __restrict__ h, const float* __restrict__ dot,
                                       const float* __restrict__ vp, const float* __restrict__ scale,
                                       const float* __restrict__ g_xo, const float* __restrict__ g_vo, int N, int F,
                                       float* __restrict__ g_x, float* __restrict__ g_h, float* __restrict__ g_dot,
                                       float* __restrict__ g_vp) {
    const int F4 = F >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * F4) return;
    const int n = (int)(idx / F4), f = ((int)(idx - (int64_t)n * F4)) << 2;
    const float sc = *scale;
    const float s = sc != 0.f ? sc : 1.0f;
    const float is2 = 0.70710678118654752440f;
    const float* hr = h + (int64_t)n * 3 * F + f;
    const float4 bq = *reinterpret_cast<const float4*>(hr + F), cg = *reinterpret_cast<const float4*>(hr + 2 * F);
    const int64_t xo = (int64_t)n * F + f;
    const float4 dt = *reinterpret_cast<const float4*>(dot + xo);
    const float4 gx = *reinterpret_cast<const float4*>(g_xo + xo);
    const float4 gxs = make_float4(gx.x * s, gx.y * s, gx.z * s, gx.w * s);
    const float4 ga = make_float4(gxs.x * is2, gxs.y * is2, gxs.z * is2, gxs.w * is2);
    *reinterpret_cast<float4*>(g_x + xo) = gxs;
    *reinterpret_cast<float4*>(g_dot + xo) = make_float4(ga.x * bq.x, ga.y * bq.y, ga.z * bq.z, ga.w * bq.w);
    float4 gc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* r = vp + (int64_t)n * 3 * 2 * F + f;
    float* o = g_vp + (int64_t)n * 3 * 2 * F + f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float4 v1 = *reinterpret_cast<const float4*>(r + c * 2 * F);
        const float4 gv = *reinterpret_cast<const float4*>(g_vo + ((int64_t)n * 3 + c) * F + f);
        gc.x += gv.x * v1.x; gc.y += gv.y * v1.y; gc.z += gv.z * v1.z; gc.w += gv.w * v1.w;
        *reinterpret_cast<float4*>(o + c * 2 * F) = make_float4(cg.x * gv.x, cg.y * gv.y, cg.z * gv.z, cg.w * gv.w);
        *reinterpret_cast<float4*>(o + c * 2 * F + F) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float* go = g_h + (int64_t)n * 3 * F + f;
    *reinterpret_cast<float4*>(go) = ga;
    *reinterpret_cast<float4*>(go + F) = make_float4(ga.x * dt.x, ga.y * dt.y, ga.z * dt.z, ga.w * dt.w);
    *reinterpret_cast<float4*>(go + 2 * F) = gc;
}

}  // namespace

extern "C" int adk_update_prep_bwd(const float* vp, const float* g_dot, const float* g_cat, int N, int F, float* g_x,
                                   float* g_vp, void* stream) {
    if (!vp || !g_dot || !g_cat || !g_x || !g_vp || N <= 0 || F <= 0 || (F & 3)) return ADK_EINVAL;
    const int64_t n = (int64_t)N * (F >> 2);
    update_prep_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, adk::as_stream(stream)>>>(
        vp, g_dot, g_cat, N, F, 1.0f / sqrtf((float)F), g_x, g_vp);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_update_gate_bwd(const float* h, const float* dot, const float* vp, const float* scale,
                                   const float* g_x_out, const float* g_vec_out, int N, int F, float* g_x, float* g_h,
                                   float* g_dot, float* g_vp, void* stream) {
    if (!h || !dot || !vp || !scale || !g_x_out || !g_vec_out || !g_x || !g_h || !g_dot || !g_vp || N <= 0 || F <= 0 ||
        (F & 3))
        return ADK_EINVAL;
    const int64_t n = (int64_t)N * (F >> 2);
    update_gate_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, adk::as_stream(stream)>>>(
        h, dot, vp, scale, g_x_out, g_vec_out, N, F, g_x, g_h, g_dot, g_vp);
    ADK_LAUNCH_CHECK();
    return 0;
}

// ---- one call per nn.Linear pass (host-side fusion: the launches below are what `TcLinearFn` used to issue one ctypes
// call at a time; at ~40 Linear layers per step that was 7 ms of Python per training step) ------------------------
namespace {
inline int64_t pad_to(int64_t n, int64_t m) { return (n + m - 1) / m * m; }
inline uint8_t* carve(uint8_t*& p, int64_t bytes) {
    uint8_t* r = p;
    p += (bytes + 255) / 256 * 256;
    return r;
}
}  // namespace

extern "C" int64_t adk_linear_train_ws_bytes(int M, int K, int N) {
    if (M <= 0 || K <= 0 || N <= 0) return ADK_EINVAL;
    const int64_t mp = pad_to(M, 128), np = pad_to(N, 128), kred = pad_to(M, 64);
    auto planes = [](int64_t rows, int64_t cols) { return (2 * rows * cols * 2 + 255) / 256 * 256; };
    const int64_t fwd = planes(mp, K) + planes(N, K);
    const int64_t bwd = planes(mp, N) + planes(K, N) + planes(np, kred) + planes(K, kred);
    return (fwd > bwd ? fwd : bwd) + 256;
}

// y[M][N] = x[M][K] . w[N][K]^T + bias;  recs[4] <- {s_x, 1/s_x, s_w, 1/s_w} (kept by the caller for the backward)
extern "C" int adk_linear_train_fwd(const float* x, const float* w, const float* bias, int M, int K, int N, float target,
                                    float* recs, void* ws, uint32_t* scratch, uint32_t* status, float* y, void* stream) {
    if (!x || !w || !recs || !ws || !scratch || !y) return ADK_EINVAL;
    const int64_t mp = pad_to(M, 128);
    uint8_t* p = reinterpret_cast<uint8_t*>(ws);
    void* xa = carve(p, 2 * mp * K * 2);
    void* wa = carve(p, 2 * (int64_t)N * K * 2);
    int rc;
    if ((rc = adk_amax_scale(x, (int64_t)M * K, target, recs, scratch, stream)) != 0) return rc;
    if ((rc = adk_amax_scale(w, (int64_t)N * K, target, recs + 2, scratch, stream)) != 0) return rc;
    if ((rc = adk_split_f16_dev(x, K, M, K, recs, xa, mp, status, stream)) != 0) return rc;
    if ((rc = adk_split_f16_dev(w, K, N, K, recs + 2, wa, N, status, stream)) != 0) return rc;
    return adk_linear_tc_dev(xa, mp, M, wa, N, K, bias, recs, recs + 2, y, N, status, stream);
}

// dx[M][K] = g[M][N] . w[N][K]  (if dx)   and   dw[N][K] = g[M][N]^T . x[M][K]  (if dw);  recs from the forward
extern "C" int adk_linear_train_bwd(const float* g, const float* x, const float* w, int M, int K, int N, float target,
                                    const float* recs, float* rec_g, void* ws, uint32_t* scratch, uint32_t* status,
                                    float* dx, float* dw, void* stream) {
    if (!g || !x || !w || !recs || !rec_g || !ws || !scratch) return ADK_EINVAL;
    const int64_t mp = pad_to(M, 128), np = pad_to(N, 128), kred = pad_to(M, 64);
    uint8_t* p = reinterpret_cast<uint8_t*>(ws);
    int rc;
    if ((rc = adk_amax_scale(g, (int64_t)M * N, target, rec_g, scratch, stream)) != 0) return rc;
    if (dx) {
        void* ga = carve(p, 2 * mp * N * 2);
        void* wt = carve(p, 2 * (int64_t)K * N * 2);
        if ((rc = adk_split_f16_dev(g, N, M, N, rec_g, ga, mp, status, stream)) != 0) return rc;
        if ((rc = adk_split_f16_t_dev(w, K, N, K, recs + 2, wt, K, N, status, stream)) != 0) return rc;
        if ((rc = adk_linear_tc_dev(ga, mp, M, wt, K, N, nullptr, rec_g, recs + 2, dx, K, status, stream)) != 0) return rc;
    }
    if (dw) {
        void* gt = carve(p, 2 * np * kred * 2);
        void* xt = carve(p, 2 * (int64_t)K * kred * 2);
        if ((rc = adk_split_f16_t_dev(g, N, M, N, rec_g, gt, np, kred, status, stream)) != 0) return rc;
        if ((rc = adk_split_f16_t_dev(x, K, M, K, recs, xt, K, kred, status, stream)) != 0) return rc;
        if ((rc = adk_linear_tc_dev(gt, np, N, xt, K, (int)kred, nullptr, rec_g, recs, dw, K, status, stream)) != 0) return rc;
    }
    return 0;
}
