// K3 (warp-MMA variant): fused RBF featurisation + rbf_proj + PaiNN message + segmented reduction
// with everything an edge touches resident in shared memory.
//
// CTA = (one adsorbate+slab system) x (slice of 32 features of each of the three groups).  Staged once
// per CTA: the fp16x2 (hi, lo) planes of the w_rbf slice [128 centres][3][32], and the fp32 xh / vec
// slices of ALL atoms of the system [n][3][32] -- so the per-edge gathers of xh[src] / vec[src]
// (12 KB per edge from L2 in the row-tiled kernel, its dominant stall) become shared-memory reads and
// HBM/L2 traffic drops to the algorithmic minimum (every feature read once per layer).
//
// A warp walks one target row of the CSR, 16 distance-sorted in-edges at a time.  For such a chunk the
// Gaussian windows of the edges overlap almost entirely, so rbfh[16 edges][96 features] is one small
// GEMM over the union window (K = 16..48 centres): the warp evaluates the RBF values directly in the
// m16n8k16 A-fragment layout, splits them into fp16 hi/lo, loads the weight fragments with
// ldmatrix.trans and issues mma.sync (Ah.Wl, Al.Wh, Ah.Wh; fp32 accumulate) -- 3 x K/16 x 12 MMAs
// instead of 16 taps x 96 features x 16 edges scalar FMAs.  The C fragments (edge = row, feature =
// column) are multiplied by the staged xh / vec of the edge's source atom and accumulated per lane;
// one shuffle reduction per target row finishes the segmented sum.  No atomics, deterministic.
//
// The tcgen05 kernel (csrc/message_t5.cu) is the default; this one serves the shapes it does not take (other
// num_rbf, hidden % 64 != 0, systems whose staged sources exceed its shared memory).
// Reference arithmetic: models/gemnet_oc/layers/radial_basis.py:235-244, models/painn/painn_denoising.py:534-567, 443-445.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int MM_THREADS = 384;          // 12 warps: 166 registers per thread, no spills (512 threads: 128 + spills, 10 % slower)
constexpr int MM_WARPS = MM_THREADS / 32;
constexpr int MM_SF = 32;                       // features per group in a CTA slice
constexpr int MM_WROW = 208;                    // bytes per centre row of a weight plane: 3 x 64 + 16 pad
constexpr float MM_RBF_SCALE = 1024.0f;
// staged source features, per atom: [group 3][n-tile pair 2][column pair qt 4][n-tile in pair 2][2] floats + 16 pad.
// A lane's eight features of a group are two float4, and the four lanes of a quad read 16 CONTIGUOUS words;
// with a row stride of 112 = 16 (mod 32) words the 8 atoms a warp reads at once alternate between the two
// halves of the banks (a 96-float stride put them all on the same banks: measured 49 % conflict wavefronts).
constexpr int MM_SRC_STRIDE = 112;

struct MmParams {
    const int32_t* atom_off;
    const int32_t* row_sel;     // optional [N]: 1 = compute the row, 2 = pass vec through (x untouched), 0 = leave untouched
    const int32_t* row_start;
    const int32_t* row_deg;
    const int32_t* e_src;
    const float4* e_geo;
    const float* xh;
    const float* vec_in;
    const __half* wt_split;   // [2][R][3F]: transposed fp16x2 planes of w_rbf
    const float* b_rbf;
    const float* rbf_offset;
    int F, R, n_max;
    int B, sys_per_cta;   // a CTA walks sys_per_cta consecutive systems (weights staged once)
    float inv_cutoff, coeff, env_a, env_b, env_c;
    int env_p;
    float acc_scale, coeff_sqrt;
    float* x_io;
    float* vec_out;
    __half* vsplit;        // optional fp16x2 planes of vec_out, [2][vsplit_rows][F] (row = atom * 3 + xyz)
    int64_t vsplit_plane;
    float vsplit_scale;
    uint32_t* status;
};

__host__ __device__ inline size_t mm_smem_bytes(int R, int n_max) {
    return 2 * (size_t)R * MM_WROW + sizeof(float) * ((size_t)R + 96) + sizeof(float) * (size_t)n_max * 2 * MM_SRC_STRIDE +
           (((size_t)n_max * 2 + 15) & ~(size_t)15) + 64;   // + row order (int16) + row counter / slack
}

// the column offset is an immediate of the instruction: one address register per (k-step, plane) instead of six
template <int OFF>
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4+%5];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr), "n"(OFF));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// first product of a chunk: C = 0 comes from the zero register, the accumulators need no clearing
__device__ __forceinline__ void mma16816_z(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%10, %10, %10, %10};"
        : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.0f));
}
// two scaled fp32 values -> packed fp16 hi pair and lo pair
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __half2 hh = __floats2half2_rn(v0, v1);       // one cvt.rn.f16x2.f32
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(v0 - back.x, v1 - back.y);
    hi = *reinterpret_cast<const uint32_t*>(&hh);
    lo = *reinterpret_cast<const uint32_t*>(&ll);
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(MM_THREADS, 1) message_mma_kernel(MmParams P) {
    extern __shared__ __align__(16) unsigned char mm_smem[];
    const int F = P.F, R = P.R;
    unsigned char* s_wh = mm_smem;                                   // hi plane: [R] rows of MM_WROW bytes
    unsigned char* s_wl = s_wh + (size_t)R * MM_WROW;                // lo plane
    float* s_mu = reinterpret_cast<float*>(s_wl + (size_t)R * MM_WROW);
    float* s_bias = s_mu + R;                                        // [3][qt 4][nt 4][2], same order as a source row
    float* s_xh = s_bias + 96;                                       // [n][MM_SRC_STRIDE]
    float* s_vec = s_xh + (size_t)P.n_max * MM_SRC_STRIDE;           // [n][MM_SRC_STRIDE]
    int16_t* s_order = reinterpret_cast<int16_t*>(s_vec + (size_t)P.n_max * MM_SRC_STRIDE);   // rows, longest first
    int* s_next = reinterpret_cast<int*>(s_order + ((P.n_max + 7) & ~7));                   // next unclaimed position
    const int f0 = blockIdx.y * MM_SF;
    const int lane = adk::lane_id(), warp = adk::warp_id();
    const bool has_vec = P.vec_in != nullptr;

    // ---- stage the weight slice: 16-byte chunks (8 features) of wt_split[plane][k][g*F + f0 ..] ----
    // cp.async (LDGSTS): all copies of the CTA are in flight at once, no register staging
    for (int i = threadIdx.x; i < 2 * R * 12; i += MM_THREADS) {
        const int plane = i / (R * 12), rem = i - plane * (R * 12);
        const int k = rem / 12, ch = rem - k * 12;          // ch = g * 4 + chunk
        const int g = ch >> 2, c4 = ch & 3;
        const __half* src = P.wt_split + ((size_t)plane * R + k) * 3 * F + g * F + f0 + c4 * 8;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared((plane ? s_wl : s_wh) + (size_t)k * MM_WROW + ch * 16);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    for (int k = threadIdx.x; k < R; k += MM_THREADS) s_mu[k] = P.rbf_offset[k];
    if (threadIdx.x < 96) {
        const int g = threadIdx.x >> 5, f = threadIdx.x & 31;
        s_bias[g * 32 + (f >> 4) * 16 + ((f >> 1) & 3) * 4 + ((f >> 3) & 1) * 2 + (f & 1)] = P.b_rbf[g * F + f0 + f];
    }
    const int b_first = blockIdx.x * P.sys_per_cta, b_end = min(P.B, b_first + P.sys_per_cta);
    for (int b = b_first; b < b_end; ++b) {
    const int a0 = P.atom_off[b], n = P.atom_off[b + 1] - a0;
    if (n > P.n_max) continue;   // larger than the staging area this launch was sized for: another kernel owns it (row_sel == 0 there)
    // ---- stage the system's source features: one 128-byte segment per (atom, group) per warp ----
    {
        // feature f = nt * 8 + 2 * qt + h of the slice lands at [g][qt][nt][h]
        const int pos = (lane >> 4) * 16 + ((lane >> 1) & 3) * 4 + ((lane >> 3) & 1) * 2 + (lane & 1);
        for (int seg = warp; seg < n * 3; seg += MM_WARPS) {
            const int j = seg / 3, g = seg - j * 3;
            const size_t go = (size_t)(a0 + j) * 3 * F + g * F + f0 + lane;
            const uint32_t so = (uint32_t)(j * MM_SRC_STRIDE + g * 32 + pos) * 4u;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(s_xh) + so),
                         "l"(P.xh + go)
                         : "memory");
            if (has_vec)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(s_vec) + so),
                             "l"(P.vec_in + go)
                             : "memory");
        }
    }
    // Rows are claimed dynamically, longest first (LPT): with 12 warps and ~82 rows of 3-4 chunks each a fixed
    // round-robin leaves the slowest warp ~5 % behind the mean.  Rank by (degree desc, index asc): a permutation.
    for (int r = threadIdx.x; r < n; r += MM_THREADS) {
        if (P.row_sel && P.row_sel[a0 + r] != 1) continue;
        const int dr = P.row_deg[a0 + r];
        int rank = 0;
        for (int u = 0; u < n; ++u) {
            if (P.row_sel && P.row_sel[a0 + u] != 1) continue;
            const int du = P.row_deg[a0 + u];
            rank += (du > dr || (du == dr && u < r)) ? 1 : 0;
        }
        s_order[rank] = (int16_t)r;
    }
    if (threadIdx.x < 32) {   // warp 0: number of selected rows, claim counter
        int c = 0;
        for (int r = threadIdx.x; r < n; r += 32) c += (!P.row_sel || P.row_sel[a0 + r] == 1) ? 1 : 0;
        c = __reduce_add_sync(ADK_FULL_MASK, c);
        if (threadIdx.x == 0) { s_next[0] = 0; s_next[1] = c; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int qr = lane >> 2, qt = lane & 3;   // fragment coordinates: row group, column pair
    const uint32_t wh_base = (uint32_t)__cvta_generic_to_shared(s_wh);
    const uint32_t wl_base = (uint32_t)__cvta_generic_to_shared(s_wl);
    // ldmatrix.x4.trans lane address: matrices (k 0-7 | k 8-15) x (n-tile nt | nt+1)
    const int lm_krow = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int lm_ntile = lane >> 4;
    const uint32_t lm_base_hi = wh_base + (uint32_t)lm_krow * MM_WROW + (uint32_t)lm_ntile * 16;
    const uint32_t lm_plane = wl_base - wh_base;
    const float sqrt_3 = 1.7320508075688772f;
    const float inv_sqrt_h = 0.57735026918962576451f / sqrtf((float)F);   // includes the 1/sqrt(3) of x_ij2

    // pass-through rows (row_sel == 2): this slice of vec_in is copied to vec_out (+ its operand planes), so that
    // everything downstream of an unselected row still sees values of the network's own scale
    if (P.row_sel && blockIdx.z == 0) {
        for (int r = warp; r < n; r += MM_WARPS) {
            if (P.row_sel[a0 + r] != 2) continue;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const size_t off = ((size_t)(a0 + r) * 3 + c) * F + f0 + lane;
                const float v = has_vec ? P.vec_in[off] : 0.f;
                P.vec_out[off] = v;
                if (P.vsplit) {
                    __half h, l;
                    bool overflow = false;
                    adk::split_f16x2(v, P.vsplit_scale, h, l, overflow);
                    P.vsplit[off] = h;
                    P.vsplit[P.vsplit_plane + off] = l;
                    if (overflow && P.status) atomicOr(P.status, ADK_STATUS_F16_OVERFLOW);
                }
            }
        }
    }
    // gridDim.z > 1 (a handful of systems only): the target rows of a system are dealt to several CTAs
    const int n_rows = s_next[1];
    while (true) {
        int claim = 0;
        if (lane == 0) claim = atomicAdd(s_next, 1);
        claim = __shfl_sync(ADK_FULL_MASK, claim, 0) * (int)gridDim.z + (int)blockIdx.z;
        if (claim >= n_rows) break;
        const int tl = s_order[claim];
        const int t = a0 + tl;
        const int start = P.row_start[t], deg = P.row_deg[t];
        float dxa[4][2], dva[3][4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            dxa[nt][0] = dxa[nt][1] = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) dva[c][nt][0] = dva[c][nt][1] = 0.f;
        }
        // this lane's two outputs of the row-end write (see the reduce-scatter below): fetch the residual x early
        const int blk = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        float2 x_res[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const int o = 2 * blk + w;
            if (o < 4) x_res[w] = *reinterpret_cast<const float2*>(P.x_io + (size_t)t * F + f0 + o * 8 + 2 * qt);
        }
        // software pipeline: the CSR records of the next chunk are fetched while this one is processed
        int nx_src = 0;
        float4 nx_geo = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < min(16, deg)) {
            nx_src = P.e_src[start + lane];   // global index; made local at use so the load stays in flight
            nx_geo = P.e_geo[start + lane];
        }
        for (int e0 = 0; e0 < deg; e0 += 16) {
            const int cnt = min(16, deg - e0);
            // lanes 0..15 own one edge record each
            const int my_src = nx_src - a0;
            const float4 my_geo = nx_geo;
            int my_klo = 0;
            float my_s = 0.f, my_env = 0.f;
            if (lane < min(16, deg - e0 - 16)) {
                nx_src = P.e_src[start + e0 + 16 + lane];
                nx_geo = P.e_geo[start + e0 + 16 + lane];
            }
            if (lane < cnt) {
                my_s = my_geo.x * P.inv_cutoff;
                float sp;
                if (P.env_p == 5) {           // the shipped configuration: s^5 in three multiplies, no loop
                    const float s2 = my_s * my_s;
                    sp = s2 * s2 * my_s;
                } else {
                    sp = my_s;
                    for (int q = 1; q < P.env_p; ++q) sp *= my_s;
                }
                float env = 1.0f + P.env_a * sp;
                sp *= my_s; env += P.env_b * sp;
                sp *= my_s; env += P.env_c * sp;
                my_env = (my_s < 1.0f) ? env * MM_RBF_SCALE : 0.0f;
                my_klo = (int)floorf(my_s * (float)(R - 1)) - 7;
                my_klo = max(0, min(my_klo, R - 16));
            }
            // union window of the chunk (rows are sorted by distance: first edge has the lowest window)
            const int klo_first = __shfl_sync(ADK_FULL_MASK, my_klo, 0);
            const int klo_last = __shfl_sync(ADK_FULL_MASK, my_klo, cnt - 1);
            const int nks = (max(klo_last, klo_first) + 16 - klo_first + 15) >> 4;
            const int kbase = min(klo_first, R - 16 * nks);
            // this lane's two fragment rows (edges qr and qr + 8 of the chunk)
            const float s_a = __shfl_sync(ADK_FULL_MASK, my_s, qr), s_b = __shfl_sync(ADK_FULL_MASK, my_s, qr + 8);
            const float env_a = __shfl_sync(ADK_FULL_MASK, my_env, qr), env_b = __shfl_sync(ADK_FULL_MASK, my_env, qr + 8);

            float acc[12][4];
            {   // k-step 0 (every chunk has one): its first product initialises the accumulators
                const int ks = 0;
                const int k0 = kbase + ks * 16 + 2 * qt;
                // A fragment: rbf values (already scaled by MM_RBF_SCALE through env), split into hi / lo
                float g8[8];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float mu0 = s_mu[k0 + 8 * h], mu1 = s_mu[k0 + 8 * h + 1];
                    float t;
                    t = (s_a - mu0) * P.coeff_sqrt; g8[4 * h + 0] = env_a * ex2_approx(-t * t);
                    t = (s_a - mu1) * P.coeff_sqrt; g8[4 * h + 1] = env_a * ex2_approx(-t * t);
                    t = (s_b - mu0) * P.coeff_sqrt; g8[4 * h + 2] = env_b * ex2_approx(-t * t);
                    t = (s_b - mu1) * P.coeff_sqrt; g8[4 * h + 3] = env_b * ex2_approx(-t * t);
                }
                uint32_t ah[4], al[4];
                split_pair(g8[0], g8[1], ah[0], al[0]);   // (row qr,     k 2t..2t+1)
                split_pair(g8[2], g8[3], ah[1], al[1]);   // (row qr + 8, k 2t..2t+1)
                split_pair(g8[4], g8[5], ah[2], al[2]);   // (row qr,     k 2t+8..)
                split_pair(g8[6], g8[7], ah[3], al[3]);   // (row qr + 8, k 2t+8..)
                const uint32_t a_hi = lm_base_hi + (uint32_t)(kbase + ks * 16) * MM_WROW, a_lo = a_hi + lm_plane;
#pragma unroll
                for (int p4 = 0; p4 < 3; ++p4) {           // group g = p4: its four n-tiles, two ldmatrix.x4 per plane
                    uint32_t bh[8], bl[8];
                    if (p4 == 0) {
                        ldmatrix_x4_trans<0>(a_hi, bh[0], bh[1], bh[2], bh[3]);
                        ldmatrix_x4_trans<32>(a_hi, bh[4], bh[5], bh[6], bh[7]);
                        ldmatrix_x4_trans<0>(a_lo, bl[0], bl[1], bl[2], bl[3]);
                        ldmatrix_x4_trans<32>(a_lo, bl[4], bl[5], bl[6], bl[7]);
                    } else if (p4 == 1) {
                        ldmatrix_x4_trans<64>(a_hi, bh[0], bh[1], bh[2], bh[3]);
                        ldmatrix_x4_trans<96>(a_hi, bh[4], bh[5], bh[6], bh[7]);
                        ldmatrix_x4_trans<64>(a_lo, bl[0], bl[1], bl[2], bl[3]);
                        ldmatrix_x4_trans<96>(a_lo, bl[4], bl[5], bl[6], bl[7]);
                    } else {
                        ldmatrix_x4_trans<128>(a_hi, bh[0], bh[1], bh[2], bh[3]);
                        ldmatrix_x4_trans<160>(a_hi, bh[4], bh[5], bh[6], bh[7]);
                        ldmatrix_x4_trans<128>(a_lo, bl[0], bl[1], bl[2], bl[3]);
                        ldmatrix_x4_trans<160>(a_lo, bl[4], bl[5], bl[6], bl[7]);
                    }
                    // pass-major order: four independent accumulators between dependent MMAs
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma16816_z(acc[4 * p4 + nt], ah, bl[2 * nt], bl[2 * nt + 1]);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma16816(acc[4 * p4 + nt], al, bh[2 * nt], bh[2 * nt + 1]);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma16816(acc[4 * p4 + nt], ah, bh[2 * nt], bh[2 * nt + 1]);
                }
            }
            for (int ks = 1; ks < nks; ++ks) {
                const int k0 = kbase + ks * 16 + 2 * qt;
                // A fragment: rbf values (already scaled by MM_RBF_SCALE through env), split into hi / lo
                float g8[8];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float mu0 = s_mu[k0 + 8 * h], mu1 = s_mu[k0 + 8 * h + 1];
                    float t;
                    t = (s_a - mu0) * P.coeff_sqrt; g8[4 * h + 0] = env_a * ex2_approx(-t * t);
                    t = (s_a - mu1) * P.coeff_sqrt; g8[4 * h + 1] = env_a * ex2_approx(-t * t);
                    t = (s_b - mu0) * P.coeff_sqrt; g8[4 * h + 2] = env_b * ex2_approx(-t * t);
                    t = (s_b - mu1) * P.coeff_sqrt; g8[4 * h + 3] = env_b * ex2_approx(-t * t);
                }
                uint32_t ah[4], al[4];
                split_pair(g8[0], g8[1], ah[0], al[0]);   // (row qr,     k 2t..2t+1)
                split_pair(g8[2], g8[3], ah[1], al[1]);   // (row qr + 8, k 2t..2t+1)
                split_pair(g8[4], g8[5], ah[2], al[2]);   // (row qr,     k 2t+8..)
                split_pair(g8[6], g8[7], ah[3], al[3]);   // (row qr + 8, k 2t+8..)
                const uint32_t a_hi = lm_base_hi + (uint32_t)(kbase + ks * 16) * MM_WROW, a_lo = a_hi + lm_plane;
#pragma unroll
                for (int p4 = 0; p4 < 3; ++p4) {           // group g = p4: its four n-tiles, two ldmatrix.x4 per plane
                    uint32_t bh[8], bl[8];
                    if (p4 == 0) {
                        ldmatrix_x4_trans<0>(a_hi, bh[0], bh[1], bh[2], bh[3]);
                        ldmatrix_x4_trans<32>(a_hi, bh[4], bh[5], bh[6], bh[7]);
                        ldmatrix_x4_trans<0>(a_lo, bl[0], bl[1], bl[2], bl[3]);
                        ldmatrix_x4_trans<32>(a_lo, bl[4], bl[5], bl[6], bl[7]);
                    } else if (p4 == 1) {
                        ldmatrix_x4_trans<64>(a_hi, bh[0], bh[1], bh[2], bh[3]);
                        ldmatrix_x4_trans<96>(a_hi, bh[4], bh[5], bh[6], bh[7]);
                        ldmatrix_x4_trans<64>(a_lo, bl[0], bl[1], bl[2], bl[3]);
                        ldmatrix_x4_trans<96>(a_lo, bl[4], bl[5], bl[6], bl[7]);
                    } else {
                        ldmatrix_x4_trans<128>(a_hi, bh[0], bh[1], bh[2], bh[3]);
                        ldmatrix_x4_trans<160>(a_hi, bh[4], bh[5], bh[6], bh[7]);
                        ldmatrix_x4_trans<128>(a_lo, bl[0], bl[1], bl[2], bl[3]);
                        ldmatrix_x4_trans<160>(a_lo, bl[4], bl[5], bl[6], bl[7]);
                    }
                    // pass-major order: four independent accumulators between dependent MMAs
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma16816(acc[4 * p4 + nt], ah, bl[2 * nt], bl[2 * nt + 1]);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma16816(acc[4 * p4 + nt], al, bh[2 * nt], bh[2 * nt + 1]);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma16816(acc[4 * p4 + nt], ah, bh[2 * nt], bh[2 * nt + 1]);
                }
            }

            // rbfh = acc * scale + bias for both fragment rows at once: the bias is read from shared memory once per
            // chunk (it was 14 % of the kernel's shared-memory wavefronts when each row block re-read it)
#pragma unroll
            for (int g = 0; g < 3; ++g)
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    const float4 bq = *reinterpret_cast<const float4*>(s_bias + g * 32 + np * 16 + qt * 4);
                    float* a0p = acc[4 * g + 2 * np];
                    float* a1p = acc[4 * g + 2 * np + 1];
                    a0p[0] = fmaf(a0p[0], P.acc_scale, bq.x); a0p[2] = fmaf(a0p[2], P.acc_scale, bq.x);
                    a0p[1] = fmaf(a0p[1], P.acc_scale, bq.y); a0p[3] = fmaf(a0p[3], P.acc_scale, bq.y);
                    a1p[0] = fmaf(a1p[0], P.acc_scale, bq.z); a1p[2] = fmaf(a1p[2], P.acc_scale, bq.z);
                    a1p[1] = fmaf(a1p[1], P.acc_scale, bq.w); a1p[3] = fmaf(a1p[3], P.acc_scale, bq.w);
                }
            // messages of this lane's two edges (rows qr, qr + 8); acc[g * 4 + nt] = {row qr: c0 c1, row qr+8: c2 c3}
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int er = qr + 8 * hrow;
                const int src = __shfl_sync(ADK_FULL_MASK, my_src, er);
                // r_hat carries sqrt(3); the row-end scale carries 1/sqrt(3) for the whole dvec sum
                const float rx = __shfl_sync(ADK_FULL_MASK, my_geo.y, er) * sqrt_3;
                const float ry = __shfl_sync(ADK_FULL_MASK, my_geo.z, er) * sqrt_3;
                const float rz = __shfl_sync(ADK_FULL_MASK, my_geo.w, er) * sqrt_3;
                if (er < cnt) {
                    const float* xs = s_xh + (size_t)src * MM_SRC_STRIDE + qt * 4;
                    const float* vs = s_vec + (size_t)src * MM_SRC_STRIDE + qt * 4;
                    float h1[8], h2[8], h3[8];
                    *reinterpret_cast<float4*>(h1) = *reinterpret_cast<const float4*>(xs);
                    *reinterpret_cast<float4*>(h1 + 4) = *reinterpret_cast<const float4*>(xs + 16);
                    *reinterpret_cast<float4*>(h2) = *reinterpret_cast<const float4*>(xs + 32);
                    *reinterpret_cast<float4*>(h2 + 4) = *reinterpret_cast<const float4*>(xs + 48);
                    *reinterpret_cast<float4*>(h3) = *reinterpret_cast<const float4*>(xs + 64);
                    *reinterpret_cast<float4*>(h3 + 4) = *reinterpret_cast<const float4*>(xs + 80);
                    float m2[8];
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float r1 = acc[nt][2 * hrow + h], r2 = acc[4 + nt][2 * hrow + h], r3 = acc[8 + nt][2 * hrow + h];
                            dxa[nt][h] = fmaf(h1[2 * nt + h], r1, dxa[nt][h]);
                            const float m3 = h3[2 * nt + h] * r3;
                            dva[0][nt][h] = fmaf(m3, rx, dva[0][nt][h]);
                            dva[1][nt][h] = fmaf(m3, ry, dva[1][nt][h]);
                            dva[2][nt][h] = fmaf(m3, rz, dva[2][nt][h]);
                            m2[2 * nt + h] = h2[2 * nt + h] * r2;
                        }
                    if (has_vec) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            float vj[8];
                            *reinterpret_cast<float4*>(vj) = *reinterpret_cast<const float4*>(vs + c * 32);
                            *reinterpret_cast<float4*>(vj + 4) = *reinterpret_cast<const float4*>(vs + c * 32 + 16);
#pragma unroll
                            for (int i = 0; i < 8; ++i) dva[c][i >> 1][i & 1] = fmaf(vj[i], m2[i], dva[c][i >> 1][i & 1]);
                        }
                    }
                }
            }
        }
        // segmented sum over the row's edges = sum over the 8 lanes that share this column pair (qt).  Done as a
        // reduce-scatter: 32 partial values per lane (16 float2 outputs: dx[4 n-tiles], dvec[3][4]) are halved three
        // times (xor 16, 8, 4), 28 shuffles instead of 96, and every lane ends up owning TWO finished outputs,
        // so the write-out (fp32 + fp16x2 planes) is spread over all 32 lanes instead of 4.
        float v32[32];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            v32[2 * nt] = dxa[nt][0];
            v32[2 * nt + 1] = dxa[nt][1];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                v32[8 + (c * 4 + nt) * 2] = dva[c][nt][0];
                v32[8 + (c * 4 + nt) * 2 + 1] = dva[c][nt][1];
            }
        }
        float v16[16], v8[8], v4[4];
        {
            const bool up = (lane & 16) != 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float mine = up ? v32[16 + i] : v32[i], other = up ? v32[i] : v32[16 + i];
                v16[i] = mine + __shfl_xor_sync(ADK_FULL_MASK, other, 16);
            }
        }
        {
            const bool up = (lane & 8) != 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float mine = up ? v16[8 + i] : v16[i], other = up ? v16[i] : v16[8 + i];
                v8[i] = mine + __shfl_xor_sync(ADK_FULL_MASK, other, 8);
            }
        }
        {
            const bool up = (lane & 4) != 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float mine = up ? v8[4 + i] : v8[i], other = up ? v8[i] : v8[4 + i];
                v4[i] = mine + __shfl_xor_sync(ADK_FULL_MASK, other, 4);
            }
        }
        // this lane owns outputs o0 = 2 * blk and o0 + 1 (o < 4: dx of n-tile o; else dvec[(o-4)/4] of n-tile (o-4)%4)
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const int o = 2 * blk + w;
            const float rx0 = v4[2 * w], rx1 = v4[2 * w + 1];
            if (o < 4) {
                const int fo = f0 + o * 8 + 2 * qt;
                *reinterpret_cast<float2*>(P.x_io + (size_t)t * F + fo) =
                    make_float2((x_res[w].x + rx0) * 0.70710678118654752440f, (x_res[w].y + rx1) * 0.70710678118654752440f);
            } else {
                const int c = (o - 4) >> 2, nt = (o - 4) & 3;
                const int fo = f0 + nt * 8 + 2 * qt;
                float2 base = make_float2(0.f, 0.f);
                if (has_vec) base = *reinterpret_cast<const float2*>(s_vec + (size_t)tl * MM_SRC_STRIDE + c * 32 + (nt >> 1) * 16 + qt * 4 + (nt & 1) * 2);
                const float2 vo = make_float2(base.x + rx0 * inv_sqrt_h, base.y + rx1 * inv_sqrt_h);
                *reinterpret_cast<float2*>(P.vec_out + (size_t)t * 3 * F + c * F + fo) = vo;
                if (P.vsplit) {  // operand planes of the vec_proj GEMM that follows
                    __half h0, l0, h1, l1;
                    bool overflow = false;
                    adk::split_f16x2(vo.x, P.vsplit_scale, h0, l0, overflow);
                    adk::split_f16x2(vo.y, P.vsplit_scale, h1, l1, overflow);
                    const size_t off = ((size_t)t * 3 + c) * F + fo;
                    *reinterpret_cast<__half2*>(P.vsplit + off) = __halves2half2(h0, h1);
                    *reinterpret_cast<__half2*>(P.vsplit + P.vsplit_plane + off) = __halves2half2(l0, l1);
                    if (overflow && P.status) atomicOr(P.status, ADK_STATUS_F16_OVERFLOW);
                }
            }
        }
    }
    __syncthreads();   // every warp is done with this system's staged sources before the next system overwrites them
    }
}

// w[rows][cols] fp32 -> fp16x2 planes of the TRANSPOSE: dst[2][cols][rows]
__global__ void split_transpose_kernel(const float* __restrict__ w, int rows, int cols, float scale,
                                       __half* __restrict__ dst, uint32_t* status) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int i = ty; i < 32; i += 8) tile[i][tx] = (r0 + i < rows && c0 + tx < cols) ? w[(size_t)(r0 + i) * cols + c0 + tx] : 0.f;
    __syncthreads();
    bool overflow = false;
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (c < cols && r < rows) {
            const float sv = tile[tx][i] * scale;
            overflow |= !(fabsf(sv) <= 65504.0f);
            const __half hi = __float2half_rn(sv);
            dst[(size_t)c * rows + r] = hi;
            dst[(size_t)cols * rows + (size_t)c * rows + r] = __float2half_rn(sv - __half2float(hi));
        }
    }
    if (overflow && status) atomicOr(status, ADK_STATUS_F16_OVERFLOW);
}

}  // namespace

extern "C" int adk_split_f16_transpose(const float* w, int rows, int cols, float scale, void* dst, uint32_t* status,
                                       void* stream) {
    if (!w || !dst || rows <= 0 || cols <= 0) return ADK_EINVAL;
    split_transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), 256, 0, adk::as_stream(stream)>>>(
        w, rows, cols, scale, reinterpret_cast<__half*>(dst), status);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t adk_message_mma_smem_bytes(int R, int n_max) {
    if (R < 16 || R > 128 || (R & 15) || n_max <= 0) return ADK_ERANGE;
    const size_t b = mm_smem_bytes(R, n_max);
    return b > 227 * 1024 ? (int64_t)ADK_ERANGE : (int64_t)b;
}

extern "C" int adk_message_mma(const int32_t* atom_off, int B, int n_max, const int32_t* row_sel, const int32_t* row_start,
                               const int32_t* row_deg, const int32_t* e_src, const float* e_geo, const float* xh,
                               const float* vec_in, const void* wt_split, float w_scale, const float* b_rbf,
                               const float* rbf_offset, int F, int R, float cutoff, int envelope_exponent,
                               float comp, float* x_io, float* vec_out, void* vec_split, int64_t split_rows,
                               float split_scale, uint32_t* status, void* stream) {
    if (!atom_off || !row_start || !row_deg || !e_src || !e_geo || !xh || !wt_split || !b_rbf || !rbf_offset ||
        !x_io || !vec_out || B <= 0 || n_max <= 0)
        return ADK_EINVAL;
    if (F % MM_SF != 0 || (F & 7) || R < 16 || R > 128 || (R & 15) || envelope_exponent < 1 || vec_in == vec_out)
        return ADK_EINVAL;
    const size_t smem = mm_smem_bytes(R, n_max);
    if (smem > 227 * 1024) return ADK_ERANGE;
    MmParams P;
    P.atom_off = atom_off; P.row_sel = row_sel; P.row_start = row_start; P.row_deg = row_deg; P.e_src = e_src;
    P.e_geo = reinterpret_cast<const float4*>(e_geo);
    P.xh = xh; P.vec_in = vec_in; P.wt_split = reinterpret_cast<const __half*>(wt_split);
    P.b_rbf = b_rbf; P.rbf_offset = rbf_offset;
    P.F = F; P.R = R; P.n_max = n_max;
    P.inv_cutoff = (float)(1.0 / (double)cutoff);
    const double spacing = 1.0 / (double)(R - 1);
    P.coeff = (float)(-0.5 / (spacing * spacing));
    const double p = (double)envelope_exponent;
    P.env_p = envelope_exponent;
    P.env_a = (float)(-(p + 1) * (p + 2) / 2);
    P.env_b = (float)(p * (p + 2));
    P.env_c = (float)(-p * (p + 1) / 2);
    P.acc_scale = (1.0f + comp) / (MM_RBF_SCALE * w_scale);
    P.coeff_sqrt = (float)(sqrt(0.5 * 1.4426950408889634) / spacing);  // exp(coeff d^2) = 2^-(coeff_sqrt d)^2
    P.x_io = x_io; P.vec_out = vec_out;
    P.vsplit = reinterpret_cast<__half*>(vec_split); P.vsplit_plane = split_rows * (int64_t)F;
    P.vsplit_scale = split_scale; P.status = status;
    // When only a few rows per system are selected (the sampler's last layer: the adsorbate atoms), staging the weight
    // slice dominates a CTA: let one CTA walk several systems and stage the weights once.
    P.B = B;
    P.sys_per_cta = 1;
    if (row_sel && B >= 148) P.sys_per_cta = 4;
    // one CTA per (system, feature slice) fills the GPU from ~10 systems on; below that split the rows as well,
    // as long as all CTAs are resident at once (one per SM)
    int row_splits = 1;
    while (row_splits < 8 && (long long)B * (F / MM_SF) * row_splits * 2 <= 148) row_splits *= 2;
    message_mma_kernel<<<dim3((B + P.sys_per_cta - 1) / P.sys_per_cta, F / MM_SF, row_splits), MM_THREADS, smem, adk::as_stream(stream)>>>(P);
    ADK_LAUNCH_CHECK();
    return 0;
}

int adk_message_mma_set_attrs() {
    return (int)cudaFuncSetAttribute(message_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
