// Backward of the fused message op (training step).
//
// Forward (csrc/message_t5.cu / message_mma.cu / message.cu; reference: models/painn/painn_denoising.py:534-567,
// models/gemnet_oc/layers/radial_basis.py:235-244), per in-edge e = (j -> i) with r_g[e] = W_g . rbf[e] + b_g:
//     dx[i]      = sum_e xh1[j] * r1[e]
//     dvec[i][c] = sum_e ( vec[j][c] * xh2[j] * r2[e] * k2  +  xh3[j] * r3[e] * rhat_c[e] * k3 ),  k2 = 1/sqrt(3 F), k3 = 1/sqrt(F)
// Given g_x = dL/d dx [N][F] and g_v = dL/d dvec [N][3][F] this produces
//     d_xh [N][3F], d_vec [N][3][F] (the part that flows through the messages), d_W [3F][R], d_b [3F].
// Like the forward: no per-edge tensor in HBM (the reference's autograd keeps rbf [E][R], rbf_proj(rbf) [E][3F] and the
// messages [E][3][F] per layer), no atomics, deterministic.
//
//  * Gradients w.r.t. the SOURCE features need the transpose of the aggregation (a sum over the out-edges of j).  The
//    edge list is symmetric by construction (symmetrize_edges, painn_denoising.py:262-327: every kept edge is mirrored
//    with the same distance and the negated unit vector), so the out-edges of j are the mirrors of its in-edges and one
//    pass over row j of the SAME in-edge CSR does it: the BWD instantiation of the forward kernel (csrc/message.cu),
//    which gathers (g_x, g_v) of the atom at the other end where the forward gathers (xh, vec).
//  * d_W[r][k] = sum_e d_r[e][r] * rbf_k(e) is a [3F x E] . [E x R] contraction over all edges with a banded right
//    operand (16 live taps per edge), as many FLOPs as rbf_proj itself.  A thread owns one feature (its three
//    projection rows): every gather is a coalesced 128-byte line per warp and the per-edge factors d_r need no
//    exchange between lanes.  The difficulty is the accumulator: 3 x 128 sums per thread do not fit registers, and
//    the 16-tap window moves with the edge's distance.  Solution: `message_bwd_plan_kernel` (once per graph, reused by
//    the six layers) sorts each row chunk's edges by q = klo >> 4 into a flat list (stable, deterministic) that
//    carries everything per edge (source, target row, scaled distance, envelope, unit vector).  The weight kernel
//    walks that list: during pass q only the 32 taps [16 q, 16 q + 32) can be live, they sit in 96 registers, the
//    lower half is STORED (each partial-sum slot is written exactly once, no read-modify-write, no shared-memory
//    accumulators) when the pass ends and the upper half becomes the lower half of the next pass.  The 32 tap values
//    of an edge are evaluated by the 32 lanes of the warp (one exp each) and broadcast through 128 bytes of shared
//    memory.  Chunks' partial sums are added in a fixed order by `reduce_chunks_kernel`.
// Exact fp32 SIMT arithmetic: this is the training path (tens of systems per GPU per step).
#include "common.cuh"

namespace {

constexpr int BW_TAPS = 16;
constexpr int BW_SLOTS = 8;          // R = 128 = 8 slots of 16 taps
constexpr int BW_HDR = 16;           // ints per chunk header: [0..8] list offsets of the 8 passes (+ end), [9] i0, [10] i1
constexpr int BW_ENTRY = 12;         // ints per list entry: {src, row, klo, -} {s, env, -, -} {rx, ry, rz, -}
constexpr int BWP_THREADS = 128;
constexpr int BWP_MAX_ROWS = 2048;   // rows per chunk the plan kernel can rank (dynamic shared memory: 9 ints per row)
constexpr int BWW_THREADS = 128;     // = features per CTA of the weight kernel

struct BwParams {
    const float* xh;        // [N][3F]
    const float* vec;       // [N][3][F] or null (layer 0)
    const float* offset;    // [R] Gaussian centres (scaled distance)
    const float* g_x;       // [N][F]
    const float* g_v;       // [N][3][F]
    const int32_t* plan;
    int N, F, R, chunks;
    float coeff, k2, k3;
};

// One CTA per chunk of target rows.  klo = first of the 16 live centres, exactly as the forward kernels pick it.
__global__ void __launch_bounds__(BWP_THREADS) message_bwd_plan_kernel(
    const int32_t* __restrict__ row_start, const int32_t* __restrict__ row_deg, const int32_t* __restrict__ e_src,
    const float4* __restrict__ e_geo, int N, int R, int rows_per_chunk, float inv_cutoff, int env_p, float env_a,
    float env_b, float env_c, int32_t* __restrict__ plan, int chunks) {
    extern __shared__ int s_cnt[];   // [rows][9]: per row the number of edges per pass, then their list offsets
    __shared__ int s_red[BWP_THREADS];
    __shared__ int s_tot[BW_SLOTS + 1];
    const int c = blockIdx.x, tid = threadIdx.x;
    const int i0 = min(N, c * rows_per_chunk), i1 = min(N, i0 + rows_per_chunk), nrows = i1 - i0;
    auto klo_of = [&](float d) {
        const float sc = d * inv_cutoff;
        int klo = (int)floorf(sc * (float)(R - 1)) - 7;
        return max(0, min(klo, R - BW_TAPS));
    };
    // edges before this chunk (the list is packed in row order)
    int before = 0;
    for (int i = tid; i < i0; i += BWP_THREADS) before += row_deg[i];
    s_red[tid] = before;
    for (int r = tid; r < nrows; r += BWP_THREADS) {
        int cnt[BW_SLOTS];
#pragma unroll
        for (int q = 0; q < BW_SLOTS; ++q) cnt[q] = 0;
        const int start = row_start[i0 + r], deg = row_deg[i0 + r];
        for (int e = 0; e < deg; ++e) {
            const int q = klo_of(e_geo[start + e].x) >> 4;
#pragma unroll
            for (int u = 0; u < BW_SLOTS; ++u) cnt[u] += (u == q) ? 1 : 0;
        }
#pragma unroll
        for (int q = 0; q < BW_SLOTS; ++q) s_cnt[r * 9 + q] = cnt[q];
    }
    __syncthreads();
    if (tid == 0) {
        int b = 0;
        for (int t = 0; t < BWP_THREADS; ++t) b += s_red[t];
        s_tot[BW_SLOTS] = b;   // base of this chunk's list region
    }
    if (tid < BW_SLOTS) {      // pass tid: exclusive scan of its per-row counts
        int run = 0;
        for (int r = 0; r < nrows; ++r) {
            const int v = s_cnt[r * 9 + tid];
            s_cnt[r * 9 + tid] = run;
            run += v;
        }
        s_tot[tid] = run;
    }
    __syncthreads();
    int pass_base[BW_SLOTS + 1];
    pass_base[0] = s_tot[BW_SLOTS];
#pragma unroll
    for (int q = 0; q < BW_SLOTS; ++q) pass_base[q + 1] = pass_base[q] + s_tot[q];
    if (tid <= BW_SLOTS) plan[c * BW_HDR + tid] = pass_base[tid];
    if (tid == 9) plan[c * BW_HDR + 9] = i0;
    if (tid == 10) plan[c * BW_HDR + 10] = i1;
    int4* list = reinterpret_cast<int4*>(plan + (size_t)chunks * BW_HDR);
    for (int r = tid; r < nrows; r += BWP_THREADS) {
        int pos[BW_SLOTS];
#pragma unroll
        for (int q = 0; q < BW_SLOTS; ++q) pos[q] = pass_base[q] + s_cnt[r * 9 + q];
        const int start = row_start[i0 + r], deg = row_deg[i0 + r];
        for (int e = 0; e < deg; ++e) {
            const float4 geo = e_geo[start + e];
            const float sc = geo.x * inv_cutoff;
            float sp = sc;
            for (int u = 1; u < env_p; ++u) sp *= sc;
            float env = 1.0f + env_a * sp;
            sp *= sc; env += env_b * sp;
            sp *= sc; env += env_c * sp;
            env = sc < 1.0f ? env : 0.0f;
            const int klo = klo_of(geo.x), q = klo >> 4;
            int at = 0;
#pragma unroll
            for (int u = 0; u < BW_SLOTS; ++u)
                if (u == q) at = pos[u]++;
            int4* ent = list + (size_t)at * 3;
            ent[0] = make_int4(e_src[start + e], i0 + r, klo, 0);
            ent[1] = make_int4(__float_as_int(sc), __float_as_int(env), 0, 0);
            ent[2] = make_int4(__float_as_int(geo.y), __float_as_int(geo.z), __float_as_int(geo.w), 0);
        }
    }
}

struct BwItem {
    int4 a, b, c;          // the list entry
};
struct BwFeat {
    float x[3], v[3], gx, gv[3];
};

// grid (F / 128, chunks), 128 threads (thread = feature), 3 CTAs per SM
__global__ void __launch_bounds__(BWW_THREADS, 3) message_bwd_weights_kernel(BwParams P, float* __restrict__ part_w,
                                                                            float* __restrict__ part_b) {
    __shared__ __align__(16) float s_tap[BWW_THREADS / 32][2][32];
    const int F = P.F, R = P.R;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x * BWW_THREADS + threadIdx.x;
    const int c = blockIdx.y;
    const bool has_vec = P.vec != nullptr;
    const int32_t* hdr = P.plan + (size_t)c * BW_HDR;
    const int4* list = reinterpret_cast<const int4*>(P.plan + (size_t)P.chunks * BW_HDR);
    float W[3][32];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int r = 0; r < 32; ++r) W[g][r] = 0.f;
    float db[3] = {0.f, 0.f, 0.f};
    float* pw = part_w + (size_t)c * 3 * F * R;

    auto load_item = [&](int idx) {
        BwItem it;
        const int4* e = list + (size_t)idx * 3;
        it.a = e[0]; it.b = e[1]; it.c = e[2];
        return it;
    };
    auto load_feat = [&](const BwItem& it) {
        BwFeat ft;
        const float* xs = P.xh + (size_t)it.a.x * 3 * F + f;
        ft.x[0] = xs[0]; ft.x[1] = xs[F]; ft.x[2] = xs[2 * F];
        ft.v[0] = ft.v[1] = ft.v[2] = 0.f;
        if (has_vec) {
            const float* vs = P.vec + (size_t)it.a.x * 3 * F + f;
            ft.v[0] = vs[0]; ft.v[1] = vs[F]; ft.v[2] = vs[2 * F];
        }
        ft.gx = P.g_x[(size_t)it.a.y * F + f];
        const float* gs = P.g_v + (size_t)it.a.y * 3 * F + f;
        ft.gv[0] = gs[0]; ft.gv[1] = gs[F]; ft.gv[2] = gs[2 * F];
        return ft;
    };

    const int end_all = hdr[BW_SLOTS];
    int idx = hdr[0];
    int par = 0;
    // software pipeline: entries two items ahead, gathers one item ahead
    BwItem it0 = load_item(min(idx, max(end_all - 1, 0))), it1 = load_item(min(idx + 1, max(end_all - 1, 0)));
    BwFeat f0 = load_feat(it0);
    for (int q = 0; q < BW_SLOTS; ++q) {
        const int end = hdr[q + 1];
        for (; idx < end; ++idx) {
            const BwItem it2 = load_item(min(idx + 2, end_all - 1));
            const BwFeat f1 = load_feat(it1);
            // this lane's tap of the pass window: centre 16 q + lane, live inside [klo, klo + 16)
            const int o = it0.a.z - 16 * q;
            const float sc = __int_as_float(it0.b.x), env = __int_as_float(it0.b.y);
            const float diff = sc - P.offset[min(16 * q + lane, R - 1)];
            const float tap = (lane >= o && lane < o + BW_TAPS) ? env * expf(P.coeff * diff * diff) : 0.f;
            s_tap[warp][par][lane] = tap;
            // the three per-edge factors of this feature
            const float rx = __int_as_float(it0.c.x), ry = __int_as_float(it0.c.y), rz = __int_as_float(it0.c.z);
            float dr[3];
            dr[0] = f0.x[0] * f0.gx;
            dr[1] = f0.x[1] * (f0.v[0] * f0.gv[0] + f0.v[1] * f0.gv[1] + f0.v[2] * f0.gv[2]) * P.k2;
            dr[2] = f0.x[2] * (rx * f0.gv[0] + ry * f0.gv[1] + rz * f0.gv[2]) * P.k3;
#pragma unroll
            for (int g = 0; g < 3; ++g) db[g] += dr[g];
            __syncwarp();
            const float4* tp = reinterpret_cast<const float4*>(s_tap[warp][par]);
            // the 16 live taps [o, o + 16) lie inside five consecutive 4-tap chunks starting at chunk o >> 2: a
            // warp-uniform switch selects that chunk range with STATIC register indices (60 FMAs instead of 96)
            switch (o >> 2) {
#define ADK_BWW_CASE(C)                                                              \
    case C:                                                                          \
        _Pragma("unroll") for (int r4 = C; r4 < C + 5; ++r4) {                       \
            const float4 t4 = tp[r4];                                                \
            _Pragma("unroll") for (int g = 0; g < 3; ++g) {                          \
                W[g][4 * r4 + 0] = fmaf(dr[g], t4.x, W[g][4 * r4 + 0]);              \
                W[g][4 * r4 + 1] = fmaf(dr[g], t4.y, W[g][4 * r4 + 1]);              \
                W[g][4 * r4 + 2] = fmaf(dr[g], t4.z, W[g][4 * r4 + 2]);              \
                W[g][4 * r4 + 3] = fmaf(dr[g], t4.w, W[g][4 * r4 + 3]);              \
            }                                                                        \
        }                                                                            \
        break;
                ADK_BWW_CASE(0) ADK_BWW_CASE(1) ADK_BWW_CASE(2)
                default: {   // o >> 2 == 3: chunks 3..7 (the window cannot start beyond tap 15)
#pragma unroll
                    for (int r4 = 3; r4 < 8; ++r4) {
                        const float4 t4 = tp[r4];
#pragma unroll
                        for (int g = 0; g < 3; ++g) {
                            W[g][4 * r4 + 0] = fmaf(dr[g], t4.x, W[g][4 * r4 + 0]);
                            W[g][4 * r4 + 1] = fmaf(dr[g], t4.y, W[g][4 * r4 + 1]);
                            W[g][4 * r4 + 2] = fmaf(dr[g], t4.z, W[g][4 * r4 + 2]);
                            W[g][4 * r4 + 3] = fmaf(dr[g], t4.w, W[g][4 * r4 + 3]);
                        }
                    }
                }
#undef ADK_BWW_CASE
            }
            par ^= 1;   // the other buffer is free again after the __syncwarp of the NEXT item
            it0 = it1; it1 = it2; f0 = f1;
        }
        // the pass is over: its lower 16 taps are final for this chunk
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            float4* dst = reinterpret_cast<float4*>(pw + (size_t)(g * F + f) * R + 16 * q);
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4)
                dst[r4] = make_float4(W[g][4 * r4], W[g][4 * r4 + 1], W[g][4 * r4 + 2], W[g][4 * r4 + 3]);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                W[g][r] = W[g][16 + r];
                W[g][16 + r] = 0.f;
            }
        }
    }
#pragma unroll
    for (int g = 0; g < 3; ++g) part_b[(size_t)c * 3 * F + g * F + f] = db[g];
}

// out[i] = sum over chunks (fixed order) of part[c][i]
__global__ void reduce_chunks_kernel(const float* __restrict__ part, int chunks, int64_t n, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += part[(size_t)c * n + i];
    out[i] = s;
}

}  // namespace

static int bwd_chunks(int N) {
    int chunks = N < 111 ? N : 111;   // 4 x 111 CTAs of the weight kernel = one wave of 148 SMs x 3
    const int need = (N + BWP_MAX_ROWS - 1) / BWP_MAX_ROWS;
    return chunks < need ? need : chunks;
}

extern "C" int64_t adk_message_bwd_scratch_floats(int N, int F, int R, int* chunks_out) {
    if (N <= 0 || F <= 0 || R <= 0) return ADK_EINVAL;
    const int chunks = bwd_chunks(N);
    if (chunks_out) *chunks_out = chunks;
    return (int64_t)chunks * 3 * F * (R + 1);
}

extern "C" int64_t adk_message_bwd_plan_ints(int N, int64_t e_cap) {
    if (N <= 0 || e_cap < 0) return ADK_EINVAL;
    return (int64_t)bwd_chunks(N) * BW_HDR + (int64_t)BW_ENTRY * e_cap;
}

extern "C" int adk_message_bwd_plan(const int32_t* row_start, const int32_t* row_deg, const int32_t* e_src,
                                    const float* e_geo, int N, int R, float cutoff, int envelope_exponent,
                                    int32_t* plan, void* stream) {
    if (!row_start || !row_deg || !e_src || !e_geo || !plan || N <= 0) return ADK_EINVAL;
    if (R != 16 * BW_SLOTS || envelope_exponent < 1) return ADK_EINVAL;
    const int chunks = bwd_chunks(N);
    const int rows_per_chunk = (N + chunks - 1) / chunks;
    const double p = (double)envelope_exponent;
    message_bwd_plan_kernel<<<chunks, BWP_THREADS, sizeof(int) * 9 * rows_per_chunk, adk::as_stream(stream)>>>(
        row_start, row_deg, e_src, reinterpret_cast<const float4*>(e_geo), N, R, rows_per_chunk,
        (float)(1.0 / (double)cutoff), envelope_exponent, (float)(-(p + 1) * (p + 2) / 2), (float)(p * (p + 2)),
        (float)(-p * (p + 1) / 2), plan, chunks);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_message_bwd(const int32_t* row_start, const int32_t* row_deg, const int32_t* e_src, const float* e_geo,
                               const int32_t* plan, const float* xh, const float* vec_in, const float* w_rbf,
                               const float* b_rbf, const float* rbf_offset, int N, int F, int R, float cutoff,
                               int envelope_exponent, const float* g_dx, const float* g_dvec, float* d_xh, float* d_vec,
                               float* d_w, float* d_b, float* scratch, void* stream) {
    if (!row_start || !row_deg || !e_src || !e_geo || !plan || !xh || !w_rbf || !b_rbf || !rbf_offset || !g_dx ||
        !g_dvec || !d_xh || !d_vec || !d_w || !d_b || !scratch || N <= 0)
        return ADK_EINVAL;
    if (F % BWW_THREADS != 0 || R != 16 * BW_SLOTS || envelope_exponent < 1) return ADK_EINVAL;
    int rc = adk_message_bwd_nodes(row_start, row_deg, e_src, e_geo, xh, vec_in, w_rbf, b_rbf, rbf_offset, N, F, R, cutoff,
                                   envelope_exponent, g_dx, g_dvec, d_xh, d_vec, stream);
    if (rc) return rc;
    BwParams P;
    P.xh = xh; P.vec = vec_in; P.offset = rbf_offset; P.g_x = g_dx; P.g_v = g_dvec; P.plan = plan;
    P.N = N; P.F = F; P.R = R; P.chunks = bwd_chunks(N);
    const double spacing = 1.0 / (double)(R - 1);
    P.coeff = (float)(-0.5 / (spacing * spacing));
    P.k2 = (float)(1.0 / sqrt(3.0 * (double)F));
    P.k3 = (float)(1.0 / sqrt((double)F));
    cudaStream_t st = adk::as_stream(stream);
    const int chunks = P.chunks;
    float* part_w = scratch;
    float* part_b = scratch + (size_t)chunks * 3 * F * R;
    message_bwd_weights_kernel<<<dim3(F / BWW_THREADS, chunks), BWW_THREADS, 0, st>>>(P, part_w, part_b);
    ADK_LAUNCH_CHECK();
    const int64_t nw = (int64_t)3 * F * R, nb = (int64_t)3 * F;
    reduce_chunks_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(part_w, chunks, nw, d_w);
    ADK_LAUNCH_CHECK();
    reduce_chunks_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(part_b, chunks, nb, d_b);
    ADK_LAUNCH_CHECK();
    return 0;
}

int adk_message_bwd_set_attrs() {
    return (int)cudaFuncSetAttribute(message_bwd_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(sizeof(int) * 9 * BWP_MAX_ROWS));
}
