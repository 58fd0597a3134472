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
//    operand (16 live taps per edge).  `message_bwd_weights_kernel`: a half-warp owns 4 features; its 16 lanes are the
//    16 residues k mod 16, so for every edge each lane has exactly one live tap (one exp per lane, no redundancy) and
//    keeps the 8 taps of its residue x 3 projections x 4 features in registers.  Lanes 0-11 each compute one of the 12
//    per-edge factors d_r and the half-warp exchanges them by shuffle.  A CTA walks a chunk of target rows; the chunks'
//    partial sums are added in a fixed order by `reduce_chunks_kernel`.
// Exact fp32 SIMT arithmetic: this is the training path (tens of systems per GPU per step).
#include "common.cuh"

namespace {

constexpr int BW_TAPS = 16;

struct BwParams {
    const int32_t* row_start;
    const int32_t* row_deg;
    const int32_t* e_src;
    const float4* e_geo;
    const float* xh;        // [N][3F]
    const float* vec;       // [N][3][F] or null (layer 0)
    const float* w;         // [3F][R]
    const float* b;         // [3F]
    const float* offset;    // [R] Gaussian centres (scaled distance)
    const float* g_x;       // [N][F]
    const float* g_v;       // [N][3][F]
    int N, F, R;
    float inv_cutoff, coeff, env_a, env_b, env_c;
    int env_p;
    float k2, k3;
};

constexpr int BWW_THREADS = 256;
constexpr int BWW_F = 64;        // features per CTA: 8 warps x 2 half-warps x 4 features
constexpr int BWW_SLOTS = 8;     // taps per residue lane (R = 128 = 16 residues x 8)

struct EdgeData {
    float a, u0, u1, u2;
};
constexpr int BWW_EB = 4;        // edges whose gathers are issued together (two such sets are in flight)
constexpr int BWW_STAGE = 128;   // edge records staged per pass (rows longer than this take several passes)

// grid (F / 64, chunks), 256 threads, 2 CTAs per SM.
// dynamic smem: the accumulators [8 slots x 12][256 threads] (a thread's own column: conflict-free, dynamic slot index)
__global__ void __launch_bounds__(BWW_THREADS, 2) message_bwd_weights_kernel(BwParams P, int rows_per_chunk,
                                                                            float* __restrict__ part_w,
                                                                            float* __restrict__ part_b) {
    extern __shared__ __align__(16) float s_acc[];
    __shared__ int4 s_rec[BWW_STAGE];                 // {src, klo, s, env}
    __shared__ float4 s_rh[BWW_STAGE];                // unit vector
    __shared__ float s_mu[16 * BWW_SLOTS];
    __shared__ __align__(16) float s_dr[BWW_THREADS / 32][BWW_EB][2][16];
    const int F = P.F, R = P.R, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int m = lane & 15, half = lane >> 4;
    const int fq = blockIdx.x * BWW_F + (warp * 2 + half) * 4;      // first of this half-warp's 4 features
    const int v = m < 12 ? m : 0;                                   // the per-edge factor this lane computes
    const int vg = v >> 2, vf = fq + (v & 3);
    const bool has_vec = P.vec != nullptr;
    const float kg = vg == 1 ? P.k2 : P.k3;
    for (int k = tid; k < BWW_SLOTS * 12 * BWW_THREADS; k += BWW_THREADS) s_acc[k] = 0.f;
    for (int k = tid; k < R; k += BWW_THREADS) s_mu[k] = P.offset[k];
    // Within a row the edges are sorted by distance, so q0 = klo >> 4 only grows: lane m's live tap sits in slot q0
    // (m >= klo & 15) or q0 + 1 (m < klo & 15).  Two register sets follow the row; they are folded into the shared
    // accumulators when q0 moves (a warp-uniform event, a handful of times per row).
    float lo[12], hi[12];
#pragma unroll
    for (int u = 0; u < 12; ++u) lo[u] = hi[u] = 0.f;
    int q0cur = 0;
    float db = 0.f;
    auto fold = [&](const float (&r)[12], int q) {
        if (q < BWW_SLOTS) {
#pragma unroll
            for (int u = 0; u < 12; ++u) s_acc[(q * 12 + u) * BWW_THREADS + tid] += r[u];
        }
    };

    const int i0 = blockIdx.y * rows_per_chunk, i1 = min(P.N, i0 + rows_per_chunk);
    for (int i = i0; i < i1; ++i) {
        const int start = P.row_start[i], deg = P.row_deg[i];
        if (deg == 0) continue;
        const float gx = P.g_x[(size_t)i * F + vf];
        const float g0 = P.g_v[((size_t)i * 3 + 0) * F + vf], g1 = P.g_v[((size_t)i * 3 + 1) * F + vf],
                    g2 = P.g_v[((size_t)i * 3 + 2) * F + vf];
        for (int p0 = 0; p0 < deg; p0 += BWW_STAGE) {
            const int cnt = min(BWW_STAGE, deg - p0);
            __syncthreads();   // everyone is done with the previous pass's records
            if (tid < cnt) {   // per-edge record, computed once for the whole CTA
                const float4 geo = P.e_geo[start + p0 + tid];
                const float sc = geo.x * P.inv_cutoff;
                float sp = sc;
                for (int q = 1; q < P.env_p; ++q) sp *= sc;
                float env = 1.0f + P.env_a * sp;
                sp *= sc; env += P.env_b * sp;
                sp *= sc; env += P.env_c * sp;
                env = sc < 1.0f ? env : 0.0f;
                int klo = (int)floorf(sc * (float)(R - 1)) - 7;
                klo = max(0, min(klo, R - BW_TAPS));
                s_rec[tid] = make_int4(P.e_src[start + p0 + tid], klo, __float_as_int(sc), __float_as_int(env));
                s_rh[tid] = make_float4(geo.y, geo.z, geo.w, 0.f);
            }
            __syncthreads();
            auto load4 = [&](EdgeData (&d)[BWW_EB], int b) {
#pragma unroll
                for (int t = 0; t < BWW_EB; ++t) {
                    const int e = b * BWW_EB + t;
                    d[t].a = d[t].u0 = d[t].u1 = d[t].u2 = 0.f;
                    if (e < cnt) {
                        const int src = s_rec[e].x;
                        d[t].a = P.xh[(size_t)src * 3 * F + vg * F + vf];
                        if (vg == 1) {
                            if (has_vec) {
                                const float* vj = P.vec + (size_t)src * 3 * F + vf;
                                d[t].u0 = vj[0]; d[t].u1 = vj[F]; d[t].u2 = vj[2 * F];
                            }
                        } else if (vg == 2) {
                            const float4 rh = s_rh[e];
                            d[t].u0 = rh.x; d[t].u1 = rh.y; d[t].u2 = rh.z;
                        }
                    }
                }
            };
            auto compute4 = [&](const EdgeData (&d)[BWW_EB], int b) {
                // phase 1 (branch-free, the four edges' chains interleave): this lane's factor and live tap per edge
                float tap[BWW_EB];
                int klo[BWW_EB];
                __syncwarp();
#pragma unroll
                for (int t = 0; t < BWW_EB; ++t) {
                    const int e = min(b * BWW_EB + t, cnt - 1);
                    const float sdot = vg == 0 ? gx : (d[t].u0 * g0 + d[t].u1 * g1 + d[t].u2 * g2) * kg;
                    const float dr = d[t].a * sdot;      // zero for edges past the end (d is zeroed)
                    db += m < 12 ? dr : 0.f;
                    s_dr[warp][t][half][m] = dr;
                    const int4 r = s_rec[e];
                    klo[t] = r.y;
                    const int k = r.y + ((m - r.y) & 15);  // the tap with k = m (mod 16) inside [klo, klo + 16)
                    const float diff = __int_as_float(r.z) - s_mu[k];
                    tap[t] = __int_as_float(r.w) * expf(P.coeff * diff * diff);
                }
                __syncwarp();
                // phase 2: 12 factors x this lane's tap, into the register set of its slot
#pragma unroll
                for (int t = 0; t < BWW_EB; ++t) {
                    if (b * BWW_EB + t >= cnt) break;
                    const int q0 = klo[t] >> 4;
                    if (q0 != q0cur) {
                        fold(lo, q0cur);
                        if (q0 == q0cur + 1) {
#pragma unroll
                            for (int u = 0; u < 12; ++u) lo[u] = hi[u];
                        } else {
                            fold(hi, q0cur + 1);
#pragma unroll
                            for (int u = 0; u < 12; ++u) lo[u] = 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < 12; ++u) hi[u] = 0.f;
                        q0cur = q0;
                    }
                    const float4* dv = reinterpret_cast<const float4*>(&s_dr[warp][t][half][0]);
                    const float4 v0 = dv[0], v1 = dv[1], v2 = dv[2];
                    const float val[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
                    const bool up = m < (klo[t] & 15);
                    const float tl = up ? 0.f : tap[t], th = up ? tap[t] : 0.f;
#pragma unroll
                    for (int u = 0; u < 12; ++u) {
                        lo[u] = fmaf(val[u], tl, lo[u]);
                        hi[u] = fmaf(val[u], th, hi[u]);
                    }
                }
            };
            const int nb = (cnt + BWW_EB - 1) / BWW_EB;
            EdgeData A[BWW_EB], B[BWW_EB];
            load4(A, 0);
            for (int b = 0; b < nb; b += 2) {
                load4(B, b + 1);
                compute4(A, b);
                load4(A, b + 2);
                if (b + 1 < nb) compute4(B, b + 1);
            }
        }
    }
    fold(lo, q0cur);
    fold(hi, q0cur + 1);
    float* pw = part_w + (size_t)blockIdx.y * 3 * F * R;
    for (int q = 0; q < BWW_SLOTS; ++q)
#pragma unroll
        for (int t = 0; t < 12; ++t)
            pw[(size_t)((t >> 2) * F + fq + (t & 3)) * R + q * 16 + m] = s_acc[(q * 12 + t) * BWW_THREADS + tid];
    if (m < 12) part_b[(size_t)blockIdx.y * 3 * F + vg * F + vf] = db;
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

extern "C" int64_t adk_message_bwd_scratch_floats(int N, int F, int R, int* chunks_out) {
    if (N <= 0 || F <= 0 || R <= 0) return ADK_EINVAL;
    int chunks = N < 37 ? N : 37;   // row chunks of the weight-gradient pass (8 x 37 CTAs = one wave of 148 SMs x 2)
    if (chunks_out) *chunks_out = chunks;
    return (int64_t)chunks * 3 * F * (R + 1);
}

extern "C" int adk_message_bwd(const int32_t* row_start, const int32_t* row_deg, const int32_t* e_src, const float* e_geo,
                               const float* xh, const float* vec_in, const float* w_rbf, const float* b_rbf,
                               const float* rbf_offset, int N, int F, int R, float cutoff, int envelope_exponent,
                               const float* g_dx, const float* g_dvec, float* d_xh, float* d_vec, float* d_w,
                               float* d_b, float* scratch, void* stream) {
    if (!row_start || !row_deg || !e_src || !e_geo || !xh || !w_rbf || !b_rbf || !rbf_offset || !g_dx || !g_dvec ||
        !d_xh || !d_vec || !d_w || !d_b || !scratch || N <= 0)
        return ADK_EINVAL;
    if (F % BWW_F != 0 || R != 16 * BWW_SLOTS || envelope_exponent < 1) return ADK_EINVAL;
    int rc = adk_message_bwd_nodes(row_start, row_deg, e_src, e_geo, xh, vec_in, w_rbf, b_rbf, rbf_offset, N, F, R, cutoff,
                                   envelope_exponent, g_dx, g_dvec, d_xh, d_vec, stream);
    if (rc) return rc;
    BwParams P;
    P.row_start = row_start; P.row_deg = row_deg; P.e_src = e_src; P.e_geo = reinterpret_cast<const float4*>(e_geo);
    P.xh = xh; P.vec = vec_in; P.w = w_rbf; P.b = b_rbf; P.offset = rbf_offset; P.g_x = g_dx; P.g_v = g_dvec;
    P.N = N; P.F = F; P.R = R;
    P.inv_cutoff = (float)(1.0 / (double)cutoff);
    const double spacing = 1.0 / (double)(R - 1);
    P.coeff = (float)(-0.5 / (spacing * spacing));
    const double p = (double)envelope_exponent;
    P.env_p = envelope_exponent;
    P.env_a = (float)(-(p + 1) * (p + 2) / 2);
    P.env_b = (float)(p * (p + 2));
    P.env_c = (float)(-p * (p + 1) / 2);
    P.k2 = (float)(1.0 / sqrt(3.0 * (double)F));
    P.k3 = (float)(1.0 / sqrt((double)F));
    cudaStream_t st = adk::as_stream(stream);
    int chunks = 0;
    adk_message_bwd_scratch_floats(N, F, R, &chunks);
    const int rows_per_chunk = (N + chunks - 1) / chunks;
    float* part_w = scratch;
    float* part_b = scratch + (size_t)chunks * 3 * F * R;
    const size_t smem = sizeof(float) * BWW_SLOTS * 12 * BWW_THREADS;
    message_bwd_weights_kernel<<<dim3(F / BWW_F, chunks), BWW_THREADS, smem, st>>>(P, rows_per_chunk, part_w, part_b);
    ADK_LAUNCH_CHECK();
    const int64_t nw = (int64_t)3 * F * R, nb = (int64_t)3 * F;
    reduce_chunks_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(part_w, chunks, nw, d_w);
    ADK_LAUNCH_CHECK();
    reduce_chunks_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(part_b, chunks, nb, d_b);
    ADK_LAUNCH_CHECK();
    return 0;
}

int adk_message_bwd_set_attrs() {
    return (int)cudaFuncSetAttribute(message_bwd_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(sizeof(float) * BWW_SLOTS * 12 * BWW_THREADS));
}
