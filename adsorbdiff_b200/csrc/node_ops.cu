// Small per-atom kernels around the dense contractions of PaiNN: embedding lookup, LayerNorm,
// the element-wise halves of PaiNNUpdate and of the gated-equivariant output blocks.
// Reference (under /root/reference/adsorbdiff/): models/painn/painn_denoising.py:425-426, 531,
// 601-623, 688-697; models/gemnet_oc/layers/embedding_block.py:35-43;
// modules/scaling/scale_factor.py:157-172.
#include "common.cuh"

namespace {

__global__ void embed_kernel(const int64_t* __restrict__ z, const float* __restrict__ emb, int num_elements,
                             int N, int F, float* __restrict__ x, float* __restrict__ vec) {
    // one float4 per thread over x[N][F]; the same thread clears the 3 vec rows
    const int F4 = F >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * F4) return;
    const int nidx = (int)(idx / F4), f4 = (int)(idx - (int64_t)nidx * F4);
    int e = (int)z[nidx] - 1;  // AtomEmbedding: embeddings(Z - 1)
    e = min(max(e, 0), num_elements - 1);
    reinterpret_cast<float4*>(x)[idx] = reinterpret_cast<const float4*>(emb)[(int64_t)e * F4 + f4];
    if (vec) {
        float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        float4* v = reinterpret_cast<float4*>(vec) + (int64_t)nidx * 3 * F4 + f4;
        v[0] = zero; v[F4] = zero; v[2 * F4] = zero;
    }
}

// one warp per row; two-pass (mean, then biased variance) like torch's CPU LayerNorm
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int M, int F, float eps, float* __restrict__ y,
                                 __half* __restrict__ sp, int64_t plane, float sp_scale, uint32_t* status) {
    const int row = blockIdx.x * (blockDim.x >> 5) + adk::warp_id();
    if (row >= M) return;
    const int lane = adk::lane_id();
    const float* xr = x + (int64_t)row * F;
    float s = 0.f;
    for (int f = lane; f < F; f += 32) s += xr[f];
    const float mean = adk::warp_sum(s) / (float)F;
    float v = 0.f;
    for (int f = lane; f < F; f += 32) { float d = xr[f] - mean; v += d * d; }
    const float rstd = rsqrtf(adk::warp_sum(v) / (float)F + eps);
    bool overflow = false;
    for (int f = lane; f < F; f += 32) {
        const float v = (xr[f] - mean) * rstd * gamma[f] + beta[f];
        if (y) y[(int64_t)row * F + f] = v;
        if (sp) {  // fp16x2 planes for the tensor-core GEMM that consumes the normalised features
            __half hi, lo;
            adk::split_f16x2(v, sp_scale, hi, lo, overflow);
            sp[(int64_t)row * F + f] = hi;
            sp[plane + (int64_t)row * F + f] = lo;
        }
    }
    if (overflow && status) atomicOr(status, ADK_STATUS_F16_OVERFLOW);
}

__global__ void update_prep_kernel(const float* __restrict__ x, const float* __restrict__ vp, int N, int F,
                                   float inv_sqrt_h, float* __restrict__ dot, float* __restrict__ cat,
                                   __half* __restrict__ sp, int64_t plane, float sp_scale, uint32_t* status) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * F) return;
    const int nidx = (int)(idx / F), f = (int)(idx - (int64_t)nidx * F);
    const float* r = vp + (int64_t)nidx * 3 * 2 * F;
    float d = 0.f, q = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v1 = r[c * 2 * F + f], v2 = r[c * 2 * F + F + f];
        d += v1 * v2;
        q += v2 * v2;
    }
    dot[idx] = d * inv_sqrt_h;
    const float c0 = x[idx], c1 = sqrtf(q + 1e-8f);
    if (cat) {
        cat[(int64_t)nidx * 2 * F + f] = c0;
        cat[(int64_t)nidx * 2 * F + F + f] = c1;
    }
    if (sp) {
        bool overflow = false;
        __half hi, lo;
        adk::split_f16x2(c0, sp_scale, hi, lo, overflow);
        sp[(int64_t)nidx * 2 * F + f] = hi;
        sp[plane + (int64_t)nidx * 2 * F + f] = lo;
        adk::split_f16x2(c1, sp_scale, hi, lo, overflow);
        sp[(int64_t)nidx * 2 * F + F + f] = hi;
        sp[plane + (int64_t)nidx * 2 * F + F + f] = lo;
        if (overflow && status) atomicOr(status, ADK_STATUS_F16_OVERFLOW);
    }
}

__global__ void update_gate_kernel(const float* __restrict__ h, const float* __restrict__ dot,
                                   const float* __restrict__ vp, const float* __restrict__ scale, int N, int F,
                                   float* __restrict__ x, float* __restrict__ vec) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * F) return;
    const int nidx = (int)(idx / F), f = (int)(idx - (int64_t)nidx * F);
    const float* hr = h + (int64_t)nidx * 3 * F;
    const float a = hr[f], bq = hr[F + f], c = hr[2 * F + f];
    float dx = (a + bq * dot[idx]) * 0.70710678118654752440f;
    float xn = x[idx] + dx;
    const float sc = *scale;
    if (sc != 0.f) xn *= sc;  // ScaleFactor.forward multiplies only when fitted
    x[idx] = xn;
    const float* r = vp + (int64_t)nidx * 3 * 2 * F;
    float* v = vec + (int64_t)nidx * 3 * F;
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) v[cc * F + f] += c * r[cc * 2 * F + f];
}

__global__ void head_prep_kernel(const float* __restrict__ x, const float* __restrict__ v1p, int N, int C,
                                 float* __restrict__ cat, __half* __restrict__ sp, int64_t plane, float sp_scale,
                                 uint32_t* status) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * C) return;
    const int nidx = (int)(idx / C), f = (int)(idx - (int64_t)nidx * C);
    const float* r = v1p + (int64_t)nidx * 3 * C;
    float q = r[f] * r[f] + r[C + f] * r[C + f] + r[2 * C + f] * r[2 * C + f];
    const float c0 = x[idx], c1 = sqrtf(q);  // torch.norm(dim=-2), no epsilon
    if (cat) {
        cat[(int64_t)nidx * 2 * C + f] = c0;
        cat[(int64_t)nidx * 2 * C + C + f] = c1;
    }
    if (sp) {
        bool overflow = false;
        __half hi, lo;
        adk::split_f16x2(c0, sp_scale, hi, lo, overflow);
        sp[(int64_t)nidx * 2 * C + f] = hi;
        sp[plane + (int64_t)nidx * 2 * C + f] = lo;
        adk::split_f16x2(c1, sp_scale, hi, lo, overflow);
        sp[(int64_t)nidx * 2 * C + C + f] = hi;
        sp[plane + (int64_t)nidx * 2 * C + C + f] = lo;
        if (overflow && status) atomicOr(status, ADK_STATUS_F16_OVERFLOW);
    }
}

__global__ void head_gate_kernel(const float* __restrict__ u, const float* __restrict__ v2p, int N, int Co,
                                 float* __restrict__ x_out, float* __restrict__ v_out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * Co) return;
    const int nidx = (int)(idx / Co), f = (int)(idx - (int64_t)nidx * Co);
    const float s = u[(int64_t)nidx * 2 * Co + f], g = u[(int64_t)nidx * 2 * Co + Co + f];
    if (x_out) x_out[idx] = adk::ssilu(s);
    const float* r = v2p + (int64_t)nidx * 3 * Co;
    float* v = v_out + (int64_t)nidx * 3 * Co;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c * Co + f] = g * r[c * Co + f];
}

inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

extern "C" int adk_embed(const int64_t* z, const float* emb, int num_elements, int N, int F, float* x,
                         float* vec, void* stream) {
    if (!z || !emb || !x || N <= 0 || F <= 0 || (F & 3)) return ADK_EINVAL;
    embed_kernel<<<blocks_for((int64_t)N * (F >> 2), 256), 256, 0, adk::as_stream(stream)>>>(z, emb, num_elements,
                                                                                          N, F, x, vec);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_layernorm(const float* x, const float* gamma, const float* beta, int M, int F, float eps,
                             float* y, void* y_split, int64_t split_rows, float split_scale, uint32_t* status,
                             void* stream) {
    if (!x || !gamma || !beta || (!y && !y_split) || M <= 0 || F <= 0 || (y_split && split_rows < M)) return ADK_EINVAL;
    layernorm_kernel<<<blocks_for(M, 8), 256, 0, adk::as_stream(stream)>>>(
        x, gamma, beta, M, F, eps, y, reinterpret_cast<__half*>(y_split), split_rows * (int64_t)F, split_scale, status);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_update_prep(const float* x, const float* vp, int N, int F, float* dot, float* cat,
                               void* cat_split, int64_t split_rows, float split_scale, uint32_t* status,
                               void* stream) {
    if (!x || !vp || !dot || (!cat && !cat_split) || N <= 0 || F <= 0 || (cat_split && split_rows < N)) return ADK_EINVAL;
    update_prep_kernel<<<blocks_for((int64_t)N * F, 256), 256, 0, adk::as_stream(stream)>>>(
        x, vp, N, F, 1.0f / sqrtf((float)F), dot, cat, reinterpret_cast<__half*>(cat_split),
        split_rows * 2 * (int64_t)F, split_scale, status);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_update_gate(const float* h, const float* dot, const float* vp, const float* scale, int N,
                               int F, float* x, float* vec, void* stream) {
    if (!h || !dot || !vp || !scale || !x || !vec || N <= 0 || F <= 0) return ADK_EINVAL;
    update_gate_kernel<<<blocks_for((int64_t)N * F, 256), 256, 0, adk::as_stream(stream)>>>(h, dot, vp, scale, N, F,
                                                                                         x, vec);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_head_prep(const float* x, const float* v1p, int N, int C, float* cat, void* cat_split,
                             int64_t split_rows, float split_scale, uint32_t* status, void* stream) {
    if (!x || !v1p || (!cat && !cat_split) || N <= 0 || C <= 0 || (cat_split && split_rows < N)) return ADK_EINVAL;
    head_prep_kernel<<<blocks_for((int64_t)N * C, 256), 256, 0, adk::as_stream(stream)>>>(
        x, v1p, N, C, cat, reinterpret_cast<__half*>(cat_split), split_rows * 2 * (int64_t)C, split_scale, status);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_head_gate(const float* u, const float* v2p, int N, int Co, float* x_out, float* v_out,
                             void* stream) {
    if (!u || !v2p || !v_out || N <= 0 || Co <= 0) return ADK_EINVAL;
    head_gate_kernel<<<blocks_for((int64_t)N * Co, 256), 256, 0, adk::as_stream(stream)>>>(u, v2p, N, Co, x_out,
                                                                                        v_out);
    ADK_LAUNCH_CHECK();
    return 0;
}
