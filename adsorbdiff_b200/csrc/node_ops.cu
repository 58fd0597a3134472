// Small per-atom kernels around the dense contractions of PaiNN: embedding lookup, LayerNorm,
// the element-wise halves of PaiNNUpdate and of the gated-equivariant output blocks.
// Reference (under /root/reference/adsorbdiff/): models/painn/painn_denoising.py:425-426, 531,
// 601-623, 688-697; models/gemnet_oc/layers/embedding_block.py:35-43;
// modules/scaling/scale_factor.py:157-172.
#include "common.cuh"

namespace {

__global__ void embed_kernel(const int64_t* __restrict__ z, const float* __restrict__ emb, int num_elements,
                             int N, int F, float* __restrict__ x, float* __restrict__ vec, uint32_t* status) {
    // one float4 per thread over x[N][F]; the same thread clears the 3 vec rows
    const int F4 = F >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * F4) return;
    const int nidx = (int)(idx / F4), f4 = (int)(idx - (int64_t)nidx * F4);
    int e = (int)z[nidx] - 1;  // AtomEmbedding: embeddings(Z - 1)
    if (e < 0 || e >= num_elements) {
        // nn.Embedding fails hard on such an index; here: status bit (raised by the host), row clamped meanwhile
        if (status && f4 == 0) atomicOr(status, ADK_STATUS_BAD_ELEMENT);
        e = min(max(e, 0), num_elements - 1);
    }
    reinterpret_cast<float4*>(x)[idx] = reinterpret_cast<const float4*>(emb)[(int64_t)e * F4 + f4];
    if (vec) {
        float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        float4* v = reinterpret_cast<float4*>(vec) + (int64_t)nidx * 3 * F4 + f4;
        v[0] = zero; v[F4] = zero; v[2 * F4] = zero;
    }
}

// four features per thread: 16-byte loads / stores, 8-byte stores of the packed fp16 planes
__device__ __forceinline__ void store_split4(__half* sp, int64_t plane, int64_t off, float4 v, float scale, bool& overflow) {
    __half h[4], l[4];
    adk::split_f16x2(v.x, scale, h[0], l[0], overflow);
    adk::split_f16x2(v.y, scale, h[1], l[1], overflow);
    adk::split_f16x2(v.z, scale, h[2], l[2], overflow);
    adk::split_f16x2(v.w, scale, h[3], l[3], overflow);
    const __half2 h01 = __halves2half2(h[0], h[1]), h23 = __halves2half2(h[2], h[3]);
    const __half2 l01 = __halves2half2(l[0], l[1]), l23 = __halves2half2(l[2], l[3]);
    uint2 ph, pl;
    ph.x = *reinterpret_cast<const uint32_t*>(&h01); ph.y = *reinterpret_cast<const uint32_t*>(&h23);
    pl.x = *reinterpret_cast<const uint32_t*>(&l01); pl.y = *reinterpret_cast<const uint32_t*>(&l23);
    *reinterpret_cast<uint2*>(sp + off) = ph;
    *reinterpret_cast<uint2*>(sp + plane + off) = pl;
}

// one warp per row; two-pass (mean, then biased variance) like torch's CPU LayerNorm; four features per lane and trip
__global__ void layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int M, int F, float eps, float* __restrict__ y,
                                 __half* __restrict__ sp, int64_t plane, float sp_scale, uint32_t* status) {
    const int row = blockIdx.x * (blockDim.x >> 5) + adk::warp_id();
    if (row >= M) return;
    const int lane = adk::lane_id();
    const float* xr = x + (int64_t)row * F;
    float s = 0.f;
    for (int f = 4 * lane; f < F; f += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + f);
        s += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = adk::warp_sum(s) / (float)F;
    float q = 0.f;
    for (int f = 4 * lane; f < F; f += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + f);
        const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
        q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    const float rstd = rsqrtf(adk::warp_sum(q) / (float)F + eps);
    bool overflow = false;
    for (int f = 4 * lane; f < F; f += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + f);
        const float4 g = *reinterpret_cast<const float4*>(gamma + f), bt = *reinterpret_cast<const float4*>(beta + f);
        const float4 o = make_float4((v.x - mean) * rstd * g.x + bt.x, (v.y - mean) * rstd * g.y + bt.y,
                                     (v.z - mean) * rstd * g.z + bt.z, (v.w - mean) * rstd * g.w + bt.w);
        const int64_t off = (int64_t)row * F + f;
        if (y) *reinterpret_cast<float4*>(y + off) = o;
        if (sp) store_split4(sp, plane, off, o, sp_scale, overflow);  // planes for the tensor-core GEMM that consumes them
    }
    if (overflow && status) atomicOr(status, ADK_STATUS_F16_OVERFLOW);
}

__global__ void update_prep_kernel(const float* __restrict__ x, const float* __restrict__ vp, int N, int F,
                                   float inv_sqrt_h, float* __restrict__ dot, float* __restrict__ cat,
                                   __half* __restrict__ sp, int64_t plane, float sp_scale, uint32_t* status) {
    const int F4 = F >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * F4) return;
    const int nidx = (int)(idx / F4), f = ((int)(idx - (int64_t)nidx * F4)) << 2;
    const float* r = vp + (int64_t)nidx * 3 * 2 * F;
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f), q = d;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float4 v1 = *reinterpret_cast<const float4*>(r + c * 2 * F + f);
        const float4 v2 = *reinterpret_cast<const float4*>(r + c * 2 * F + F + f);
        d.x += v1.x * v2.x; d.y += v1.y * v2.y; d.z += v1.z * v2.z; d.w += v1.w * v2.w;
        q.x += v2.x * v2.x; q.y += v2.y * v2.y; q.z += v2.z * v2.z; q.w += v2.w * v2.w;
    }
    const int64_t xo = (int64_t)nidx * F + f;
    *reinterpret_cast<float4*>(dot + xo) = make_float4(d.x * inv_sqrt_h, d.y * inv_sqrt_h, d.z * inv_sqrt_h, d.w * inv_sqrt_h);
    const float4 c0 = *reinterpret_cast<const float4*>(x + xo);
    const float4 c1 = make_float4(sqrtf(q.x + 1e-8f), sqrtf(q.y + 1e-8f), sqrtf(q.z + 1e-8f), sqrtf(q.w + 1e-8f));
    const int64_t co = (int64_t)nidx * 2 * F + f;
    if (cat) {
        *reinterpret_cast<float4*>(cat + co) = c0;
        *reinterpret_cast<float4*>(cat + co + F) = c1;
    }
    if (sp) {
        bool overflow = false;
        store_split4(sp, plane, co, c0, sp_scale, overflow);
        store_split4(sp, plane, co + F, c1, sp_scale, overflow);
        if (overflow && status) atomicOr(status, ADK_STATUS_F16_OVERFLOW);
    }
}

__global__ void update_gate_kernel(const float* __restrict__ h, const float* __restrict__ dot,
                                   const float* __restrict__ vp, const float* __restrict__ scale, int N, int F,
                                   float* __restrict__ x, float* __restrict__ vec, __half* __restrict__ vsp,
                                   int64_t vplane, float vsp_scale, uint32_t* status) {
    const int F4 = F >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * F4) return;
    const int nidx = (int)(idx / F4), f = ((int)(idx - (int64_t)nidx * F4)) << 2;
    const float* hr = h + (int64_t)nidx * 3 * F + f;
    const float4 a = *reinterpret_cast<const float4*>(hr), bq = *reinterpret_cast<const float4*>(hr + F),
                 c = *reinterpret_cast<const float4*>(hr + 2 * F);
    const int64_t xo = (int64_t)nidx * F + f;
    const float4 dt = *reinterpret_cast<const float4*>(dot + xo);
    float4 xn = *reinterpret_cast<const float4*>(x + xo);
    const float is2 = 0.70710678118654752440f;
    xn.x += (a.x + bq.x * dt.x) * is2; xn.y += (a.y + bq.y * dt.y) * is2;
    xn.z += (a.z + bq.z * dt.z) * is2; xn.w += (a.w + bq.w * dt.w) * is2;
    const float sc = *scale;
    if (sc != 0.f) { xn.x *= sc; xn.y *= sc; xn.z *= sc; xn.w *= sc; }  // ScaleFactor.forward multiplies only when fitted
    *reinterpret_cast<float4*>(x + xo) = xn;
    const float* r = vp + (int64_t)nidx * 3 * 2 * F + f;
    float* v = vec + (int64_t)nidx * 3 * F + f;
    bool overflow = false;
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
        const float4 v1 = *reinterpret_cast<const float4*>(r + cc * 2 * F);
        float4 vv = *reinterpret_cast<float4*>(v + cc * F);
        vv.x += c.x * v1.x; vv.y += c.y * v1.y; vv.z += c.z * v1.z; vv.w += c.w * v1.w;
        *reinterpret_cast<float4*>(v + cc * F) = vv;
        if (vsp) store_split4(vsp, vplane, ((int64_t)nidx * 3 + cc) * F + f, vv, vsp_scale, overflow);
    }
    if (overflow && status) atomicOr(status, ADK_STATUS_F16_OVERFLOW);
}

__global__ void head_prep_kernel(const float* __restrict__ x, const float* __restrict__ v1p, int N, int C,
                                 float* __restrict__ cat, __half* __restrict__ sp, int64_t plane, float sp_scale,
                                 uint32_t* status) {
    const int C4 = C >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * C4) return;
    const int nidx = (int)(idx / C4), f = ((int)(idx - (int64_t)nidx * C4)) << 2;
    const float* r = v1p + (int64_t)nidx * 3 * C + f;
    const float4 a = *reinterpret_cast<const float4*>(r), b = *reinterpret_cast<const float4*>(r + C),
                 c = *reinterpret_cast<const float4*>(r + 2 * C);
    const float4 c0 = *reinterpret_cast<const float4*>(x + (int64_t)nidx * C + f);
    // torch.norm(dim=-2), no epsilon
    const float4 c1 = make_float4(sqrtf(a.x * a.x + b.x * b.x + c.x * c.x), sqrtf(a.y * a.y + b.y * b.y + c.y * c.y),
                                  sqrtf(a.z * a.z + b.z * b.z + c.z * c.z), sqrtf(a.w * a.w + b.w * b.w + c.w * c.w));
    const int64_t co = (int64_t)nidx * 2 * C + f;
    if (cat) {
        *reinterpret_cast<float4*>(cat + co) = c0;
        *reinterpret_cast<float4*>(cat + co + C) = c1;
    }
    if (sp) {
        bool overflow = false;
        store_split4(sp, plane, co, c0, sp_scale, overflow);
        store_split4(sp, plane, co + C, c1, sp_scale, overflow);
        if (overflow && status) atomicOr(status, ADK_STATUS_F16_OVERFLOW);
    }
}

// VEC = 4 when Co % 4 == 0 (the F/2-wide first block), 1 for the single output column of the last block
template <int VEC>
__global__ void head_gate_kernel(const float* __restrict__ u, const float* __restrict__ v2p, int N, int Co,
                                 float* __restrict__ x_out, float* __restrict__ v_out, __half* __restrict__ vsp,
                                 int64_t vplane, float vsp_scale, uint32_t* status) {
    const int CV = Co / VEC;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)N * CV) return;
    const int nidx = (int)(idx / CV), f = ((int)(idx - (int64_t)nidx * CV)) * VEC;
    const float* ur = u + (int64_t)nidx * 2 * Co + f;
    const float* r = v2p + (int64_t)nidx * 3 * Co + f;
    float* v = v_out + (int64_t)nidx * 3 * Co + f;
    if (VEC == 4) {
        bool overflow = false;
        const float4 s4 = *reinterpret_cast<const float4*>(ur), g = *reinterpret_cast<const float4*>(ur + Co);
        if (x_out)
            *reinterpret_cast<float4*>(x_out + (int64_t)nidx * Co + f) =
                make_float4(adk::ssilu(s4.x), adk::ssilu(s4.y), adk::ssilu(s4.z), adk::ssilu(s4.w));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float4 rv = *reinterpret_cast<const float4*>(r + c * Co);
            const float4 o = make_float4(g.x * rv.x, g.y * rv.y, g.z * rv.z, g.w * rv.w);
            *reinterpret_cast<float4*>(v + c * Co) = o;
            if (vsp) store_split4(vsp, vplane, ((int64_t)nidx * 3 + c) * Co + f, o, vsp_scale, overflow);
        }
        if (overflow && status) atomicOr(status, ADK_STATUS_F16_OVERFLOW);
    } else {
        const float s1 = ur[0], g = ur[Co];
        if (x_out) x_out[(int64_t)nidx * Co + f] = adk::ssilu(s1);
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c * Co] = g * r[c * Co];
    }
}

template <bool SCATTER, typename T>   // T = float4 (W % 4 == 0) or float
__global__ void move_rows_kernel(const T* __restrict__ src, const int32_t* __restrict__ rows, int M, int WT,
                                 T* __restrict__ dst) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)M * WT) return;
    const int i = (int)(idx / WT), c = (int)(idx - (int64_t)i * WT);
    const int64_t r = rows[i];
    if (SCATTER) dst[r * WT + c] = src[(int64_t)i * WT + c];
    else dst[(int64_t)i * WT + c] = src[r * WT + c];
}

template <bool SCATTER>
int move_rows(const float* src, const int32_t* rows, int M, int W, float* dst, void* stream) {
    if (!src || !rows || !dst || M <= 0 || W <= 0) return ADK_EINVAL;
    cudaStream_t st = adk::as_stream(stream);
    if ((W & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        const int64_t n = (int64_t)M * (W >> 2);
        move_rows_kernel<SCATTER, float4><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const float4*>(src), rows, M, W >> 2, reinterpret_cast<float4*>(dst));
    } else {
        const int64_t n = (int64_t)M * W;
        move_rows_kernel<SCATTER, float><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, rows, M, W, dst);
    }
    ADK_LAUNCH_CHECK();
    return 0;
}

__global__ void fill_i32_kernel(int32_t* __restrict__ out, int N, int32_t v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) out[i] = v;
}
// one warp per selected row: mark the row and the sources of its in-edges (benign races: every writer stores 1)
__global__ void mark_sources_kernel(const int32_t* __restrict__ row_start, const int32_t* __restrict__ row_deg,
                                    const int32_t* __restrict__ e_src, const int32_t* __restrict__ sel, int n_sel,
                                    int32_t* __restrict__ out) {
    const int w = blockIdx.x * (blockDim.x >> 5) + adk::warp_id();
    if (w >= n_sel) return;
    const int t = sel[w];
    if (adk::lane_id() == 0) out[t] = 1;
    const int start = row_start[t], deg = row_deg[t];
    for (int e = adk::lane_id(); e < deg; e += 32) out[e_src[start + e]] = 1;
}

inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

extern "C" int adk_embed(const int64_t* z, const float* emb, int num_elements, int N, int F, float* x,
                         float* vec, uint32_t* status, void* stream) {
    if (!z || !emb || !x || N <= 0 || F <= 0 || (F & 3)) return ADK_EINVAL;
    embed_kernel<<<blocks_for((int64_t)N * (F >> 2), 256), 256, 0, adk::as_stream(stream)>>>(z, emb, num_elements,
                                                                                          N, F, x, vec, status);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_layernorm(const float* x, const float* gamma, const float* beta, int M, int F, float eps,
                             float* y, void* y_split, int64_t split_rows, float split_scale, uint32_t* status,
                             void* stream) {
    if (!x || !gamma || !beta || (!y && !y_split) || M <= 0 || F <= 0 || (F & 3) || (y_split && split_rows < M)) return ADK_EINVAL;
    layernorm_kernel<<<blocks_for(M, 8), 256, 0, adk::as_stream(stream)>>>(
        x, gamma, beta, M, F, eps, y, reinterpret_cast<__half*>(y_split), split_rows * (int64_t)F, split_scale, status);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_update_prep(const float* x, const float* vp, int N, int F, float* dot, float* cat,
                               void* cat_split, int64_t split_rows, float split_scale, uint32_t* status,
                               void* stream) {
    if (!x || !vp || !dot || (!cat && !cat_split) || N <= 0 || F <= 0 || (F & 3) || (cat_split && split_rows < N)) return ADK_EINVAL;
    update_prep_kernel<<<blocks_for((int64_t)N * (F >> 2), 256), 256, 0, adk::as_stream(stream)>>>(
        x, vp, N, F, 1.0f / sqrtf((float)F), dot, cat, reinterpret_cast<__half*>(cat_split),
        split_rows * 2 * (int64_t)F, split_scale, status);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_update_gate(const float* h, const float* dot, const float* vp, const float* scale, int N,
                               int F, float* x, float* vec, void* vec_split, int64_t split_rows, float split_scale,
                               uint32_t* status, void* stream) {
    if (!h || !dot || !vp || !scale || !x || !vec || N <= 0 || F <= 0 || (F & 3) ||
        (vec_split && split_rows < 3 * (int64_t)N))
        return ADK_EINVAL;
    update_gate_kernel<<<blocks_for((int64_t)N * (F >> 2), 256), 256, 0, adk::as_stream(stream)>>>(
        h, dot, vp, scale, N, F, x, vec, reinterpret_cast<__half*>(vec_split), split_rows * (int64_t)F, split_scale, status);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_head_prep(const float* x, const float* v1p, int N, int C, float* cat, void* cat_split,
                             int64_t split_rows, float split_scale, uint32_t* status, void* stream) {
    if (!x || !v1p || (!cat && !cat_split) || N <= 0 || C <= 0 || (C & 3) || (cat_split && split_rows < N)) return ADK_EINVAL;
    head_prep_kernel<<<blocks_for((int64_t)N * (C >> 2), 256), 256, 0, adk::as_stream(stream)>>>(
        x, v1p, N, C, cat, reinterpret_cast<__half*>(cat_split), split_rows * 2 * (int64_t)C, split_scale, status);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_head_gate(const float* u, const float* v2p, int N, int Co, float* x_out, float* v_out,
                             void* v_split, int64_t split_rows, float split_scale, uint32_t* status, void* stream) {
    if (!u || !v2p || !v_out || N <= 0 || Co <= 0 || (v_split && ((Co & 3) || split_rows < 3 * (int64_t)N))) return ADK_EINVAL;
    __half* vsp = reinterpret_cast<__half*>(v_split);
    const int64_t plane = split_rows * (int64_t)Co;
    if ((Co & 3) == 0)
        head_gate_kernel<4><<<blocks_for((int64_t)N * (Co >> 2), 256), 256, 0, adk::as_stream(stream)>>>(
            u, v2p, N, Co, x_out, v_out, vsp, plane, split_scale, status);
    else
        head_gate_kernel<1><<<blocks_for((int64_t)N * Co, 256), 256, 0, adk::as_stream(stream)>>>(
            u, v2p, N, Co, x_out, v_out, nullptr, 0, 0.f, status);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_gather_rows(const float* src, const int32_t* rows, int M, int W, float* dst, void* stream) {
    return move_rows<false>(src, rows, M, W, dst, stream);
}

extern "C" int adk_scatter_rows(const float* src, const int32_t* rows, int M, int W, float* dst, void* stream) {
    return move_rows<true>(src, rows, M, W, dst, stream);
}

extern "C" int adk_mark_sources(const int32_t* row_start, const int32_t* row_deg, const int32_t* e_src,
                                const int32_t* sel, int n_sel, int N, int32_t* out, void* stream) {
    if (!row_start || !row_deg || !e_src || !sel || !out || n_sel <= 0 || N <= 0) return ADK_EINVAL;
    cudaStream_t st = adk::as_stream(stream);
    fill_i32_kernel<<<(N + 255) / 256, 256, 0, st>>>(out, N, 2);
    ADK_LAUNCH_CHECK();
    mark_sources_kernel<<<(n_sel + 7) / 8, 256, 0, st>>>(row_start, row_deg, e_src, sel, n_sel, out);
    ADK_LAUNCH_CHECK();
    return 0;
}
