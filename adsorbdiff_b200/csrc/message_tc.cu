// K3 (tensor-core variant): fused RBF featurisation + rbf_proj on tcgen05 + PaiNN message +
// segmented reduction, one CTA per adsorbate+slab system.
//
// The reference's dominant contraction is rbf_proj, [E,128] x [128,1536] per layer (68 % of its
// FLOPs, models/painn/painn_denoising.py:534).  Here it never leaves the SM: for a tile of 128
// consecutive in-edges (CSR order) the generator warps evaluate the Gaussian basis x envelope
// (models/gemnet_oc/layers/radial_basis.py:235-244) and write it as the fp16x2 (hi, lo) B operand
// [128 edges x 128 centres] straight into shared memory in the UMMA 128-byte-swizzled K-major
// layout; the pre-split weights w_rbf[g][cb] (128 features x 128 centres, hi/lo) stream in by TMA;
// tcgen05.mma accumulates D[feature][edge] = sum_k W[feature][k] rbf[edge][k] in TMEM (feature =
// TMEM lane, edge = column).  An epilogue thread therefore owns ONE feature and sees the rbfh of all
// 128 edges of the tile in its registers: it gathers xh[src] / vec[src] (coalesced across the warp's
// 32 features), forms the messages (painn_denoising.py:548-555) and reduces them along the CSR row
// in registers -- no atomics, summation order = CSR order, deterministic.
//
// Per tile 12 "units" (4 feature blocks of 128 x 3 groups m1|m2|m3; 8 when vec == 0 in layer 0), each
// 24 MMAs (K = 128 = 8 k-steps x {hi*lo, lo*hi, hi*hi}; corrections first, see linear_tc.cu).
// Warp roles: 0 = TMA (weights ring, 2 x 64 KB), 1 = MMA issuer + TMEM owner, 2-5 = generators,
// 6-13 = epilogue (two sets of 4 warps; set s takes feature blocks cb with cb % 2 == s, and keeps that
// block's row accumulators in registers across tiles).  TMEM: ring of 4 slots x 128 columns.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

using namespace adk::tc;

constexpr int MT_THREADS = 448;
constexpr int MT_TILE = 128;                    // edges per tile (MMA N)
constexpr int MT_CB = 128;                      // features per block (MMA M)
constexpr uint32_t MT_SUB = 128 * 64 * 2;       // one [128 x 64] fp16 swizzled sub-tile = 16 KB
constexpr uint32_t MT_B_BYTES = 4 * MT_SUB;     // B: (hi|lo) x (k-block 0|1)
constexpr uint32_t MT_A_BYTES = 4 * MT_SUB;     // A stage: same shape for one unit
constexpr int MT_A_STAGES = 2;
constexpr int MT_SLOTS = 4;
constexpr uint32_t MT_META_BYTES = 128 * 4 + 128 * 4 + 128 * 16 + 32;  // srcoff, tgt, rhat(float4), masks
constexpr uint32_t MT_SMEM_BYTES = MT_B_BYTES + MT_A_STAGES * MT_A_BYTES + 2 * MT_META_BYTES + 256 + 1024;
constexpr float MT_RBF_SCALE = 1024.0f;

struct MtParams {
    const int32_t* atom_off;
    const int32_t* sys_counts;  // [B][2]: (raw, kept) -> E_b = 2 * kept
    const int32_t* row_deg;
    const int32_t* e_src;
    const int32_t* e_tgt;
    const float4* e_geo;
    const float* xh;
    const float* vec_in;
    const float* b_rbf;
    const float* rbf_offset;
    int F, R, k_nbrs;
    float inv_cutoff, coeff, env_a, env_b, env_c;
    int env_p;
    float acc_scale;  // 1 / (rbf scale * weight scale), times the truncation compensation
    float* x_io;
    float* vec_out;
};

struct Meta {
    int* srcoff;      // [128] element offset of the source atom's xh/vec row block (src * 3F)
    int* tgt;         // [128] global target atom, -1 on padding columns
    float4* rhat;     // [128]
    uint32_t* masks;  // [0..3] row-end mask per 32-column chunk, [4..7] valid mask
};

__device__ __forceinline__ Meta meta_at(uint8_t* base) {
    Meta m;
    m.srcoff = reinterpret_cast<int*>(base);
    m.tgt = reinterpret_cast<int*>(base + 512);
    m.rhat = reinterpret_cast<float4*>(base + 1024);
    m.masks = reinterpret_cast<uint32_t*>(base + 1024 + 2048);
    return m;
}

template <int G>
struct Acc {
    float v[G == 0 ? 1 : 3];
};

// One epilogue unit: 128 TMEM columns (edges) of feature `f` for group G (0: dx, 1: vec_j * m2, 2: m3 * rhat).
// Columns are handled 16 at a time: all gathers of a 16-column batch are issued before any of them is
// consumed, so their L2 latencies overlap (a row-end flush between two columns would otherwise serialise them).
template <int G>
__device__ __forceinline__ void epilogue_unit(const MtParams& P, const Meta& M, uint32_t t_addr, int f, float bias,
                                              float* acc, bool has_vec) {
    const int F = P.F;
    const float inv_sqrt_h = 1.0f / sqrtf((float)F);
    const float inv_sqrt_3 = 0.57735026918962576451f;
#pragma unroll 1
    for (int c = 0; c < MT_TILE / 32; ++c) {
        const uint32_t end_mask = M.masks[c], valid_mask = M.masks[4 + c];
        if (valid_mask == 0u) continue;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(t_addr + c * 32, v);
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
            float h[16];
            float vg[G == 1 ? 48 : 1];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int jj = hb * 16 + j;
                const bool ok = (valid_mask >> jj) & 1u;
                const size_t so = (size_t)M.srcoff[c * 32 + jj];
                h[j] = ok ? __ldg(P.xh + so + G * F + f) : 0.0f;
                if (G == 1) {
                    vg[3 * j + 0] = ok ? __ldg(P.vec_in + so + f) : 0.0f;
                    vg[3 * j + 1] = ok ? __ldg(P.vec_in + so + F + f) : 0.0f;
                    vg[3 * j + 2] = ok ? __ldg(P.vec_in + so + 2 * F + f) : 0.0f;
                }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int jj = hb * 16 + j;
                const int col = c * 32 + jj;
                const float rbfh = fmaf(__uint_as_float(v[jj]), P.acc_scale, bias);
                const float m = h[j] * rbfh;
                if (G == 0) {
                    acc[0] += m;
                } else if (G == 1) {
                    const float m2 = m * inv_sqrt_3 * inv_sqrt_h;
                    acc[0] = fmaf(vg[3 * j + 0], m2, acc[0]);
                    acc[1] = fmaf(vg[3 * j + 1], m2, acc[1]);
                    acc[2] = fmaf(vg[3 * j + 2], m2, acc[2]);
                } else {
                    const float4 rh = M.rhat[col];
                    const float m3 = m * inv_sqrt_h;
                    acc[0] = fmaf(m3, rh.x, acc[0]);
                    acc[1] = fmaf(m3, rh.y, acc[1]);
                    acc[2] = fmaf(m3, rh.z, acc[2]);
                }
                if ((end_mask >> jj) & 1u) {  // last in-edge of its target row: flush (warp-uniform branch)
                    const size_t t = (size_t)M.tgt[col];
                    if (G == 0) {
                        float* xo = P.x_io + t * F + f;
                        *xo = (*xo + acc[0]) * 0.70710678118654752440f;
                        acc[0] = 0.f;
                    } else if (G == 1) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            P.vec_out[t * 3 * F + q * F + f] = P.vec_in[t * 3 * F + q * F + f] + acc[q];
                            acc[q] = 0.f;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            float* vo = P.vec_out + t * 3 * F + q * F + f;
                            *vo = has_vec ? (*vo + acc[q]) : acc[q];
                            acc[q] = 0.f;
                        }
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(MT_THREADS, 1)
message_tc_kernel(const __grid_constant__ CUtensorMap tmW, MtParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));  // generic pointer to the aligned base
    const uint32_t b_smem = base;
    const uint32_t a_smem = base + MT_B_BYTES;
    uint8_t* meta_g = gbase + MT_B_BYTES + MT_A_STAGES * MT_A_BYTES;
    const uint32_t bar_base = base + MT_B_BYTES + MT_A_STAGES * MT_A_BYTES + 2 * MT_META_BYTES;
    auto a_full = [&](int s) { return bar_base + 8u * s; };
    auto a_empty = [&](int s) { return bar_base + 16u + 8u * s; };
    auto d_full = [&](int s) { return bar_base + 32u + 8u * s; };
    auto d_empty = [&](int s) { return bar_base + 64u + 8u * s; };
    const uint32_t b_full = bar_base + 96u, b_empty = bar_base + 104u;
    auto meta_full = [&](int s) { return bar_base + 112u + 8u * s; };
    auto meta_empty = [&](int s) { return bar_base + 128u + 8u * s; };
    const uint32_t tmem_slot = bar_base + 144u;

    const int warp = adk::warp_id(), lane = adk::lane_id();
    const int b = blockIdx.x;
    const int a0 = P.atom_off[b], n_atoms = P.atom_off[b + 1] - a0;
    const int E = 2 * P.sys_counts[2 * b + 1];
    const int edge_base = 2 * P.k_nbrs * a0;
    const int num_tiles = (E + MT_TILE - 1) / MT_TILE;
    const bool has_vec = P.vec_in != nullptr;
    const int F = P.F;
    const int upt = has_vec ? 12 : 8;  // units per tile

    if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < MT_A_STAGES; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
            for (int s = 0; s < MT_SLOTS; ++s) { mbar_init(d_full(s), 1); mbar_init(d_empty(s), 4); }
            mbar_init(b_full, 128);
            mbar_init(b_empty, 1);
            for (int s = 0; s < 2; ++s) { mbar_init(meta_full(s), 128); mbar_init(meta_empty(s), 8); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // unit u of a tile -> (feature block, group); consecutive units alternate between the epilogue sets
    auto unit_cb = [&](int u) { return u & 3; };
    auto unit_g = [&](int u) { return has_vec ? (u >> 2) : ((u >> 2) == 0 ? 0 : 2); };

    if (warp == 0) {
        // ===================== TMA producer: weights of every unit =====================
        if (lane == 0) {
            int n = 0;
            for (int t = 0; t < num_tiles; ++t) {
                for (int u = 0; u < upt; ++u, ++n) {
                    const int s = n & 1;
                    mbar_wait(a_empty(s), ((n >> 1) & 1) ^ 1);
                    const uint32_t dst = a_smem + s * MT_A_BYTES;
                    const int row_hi = unit_g(u) * F + unit_cb(u) * MT_CB, row_lo = 3 * F + row_hi;
                    mbar_expect_tx(a_full(s), MT_A_BYTES);
                    tma_load_2d(dst, &tmW, a_full(s), 0, row_hi);
                    tma_load_2d(dst + MT_SUB, &tmW, a_full(s), 64, row_hi);
                    tma_load_2d(dst + 2 * MT_SUB, &tmW, a_full(s), 0, row_lo);
                    tma_load_2d(dst + 3 * MT_SUB, &tmW, a_full(s), 64, row_lo);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(MT_TILE >> 3) << 17) | ((uint32_t)(MT_CB >> 4) << 24);
            int n = 0;
            for (int t = 0; t < num_tiles; ++t) {
                mbar_wait(b_full, t & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int u = 0; u < upt; ++u, ++n) {
                    const int s = n & 1, slot = n & 3;
                    mbar_wait(d_empty(slot), ((n >> 2) & 1) ^ 1);
                    mbar_wait(a_full(s), (n >> 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d_tmem = tmem_base + (uint32_t)slot * MT_TILE;
                    const uint32_t sa = a_smem + s * MT_A_BYTES;
                    // corrections first (tiny accumulator), then the eight hi*hi steps
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t ah = umma_desc(sa + kb * MT_SUB), al = umma_desc(sa + (2 + kb) * MT_SUB);
                        const uint64_t bh = umma_desc(b_smem + kb * MT_SUB), bl = umma_desc(b_smem + (2 + kb) * MT_SUB);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t koff = (uint64_t)(ks * 2);
                            umma_f16(d_tmem, ah + koff, bl + koff, idesc, (kb | ks) != 0);
                            umma_f16(d_tmem, al + koff, bh + koff, idesc, 1u);
                        }
                    }
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t ah = umma_desc(sa + kb * MT_SUB), bh = umma_desc(b_smem + kb * MT_SUB);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) umma_f16(d_tmem, ah + (uint64_t)(ks * 2), bh + (uint64_t)(ks * 2), idesc, 1u);
                    }
                    umma_commit(a_empty(s));
                    umma_commit(d_full(slot));
                    if (u == upt - 1) umma_commit(b_empty);  // B tile may be overwritten once these MMAs retire
                }
            }
        }
    } else if (warp < 6) {
        // ===================== generators: RBF operand tile + per-edge metadata =====================
        const int r = (warp - 2) * 32 + lane;  // this thread's edge row inside the tile
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        const int R = P.R;
        for (int t = 0; t < num_tiles; ++t) {
            const int e = t * MT_TILE + r;
            const bool valid = e < E;
            int src = 0, tgt = -1, next_tgt = -2;
            float4 geo = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                src = P.e_src[edge_base + e];
                tgt = P.e_tgt[edge_base + e];
                geo = P.e_geo[edge_base + e];
                next_tgt = (e + 1 < E) ? P.e_tgt[edge_base + e + 1] : -2;
            }
            // the 24 centres of the three aligned 8-blocks that cover the 16-tap window
            const float s = geo.x * P.inv_cutoff;
            float sp = s;
            for (int q = 1; q < P.env_p; ++q) sp *= s;
            float env = 1.0f + P.env_a * sp;
            sp *= s; env += P.env_b * sp;
            sp *= s; env += P.env_c * sp;
            env = (valid && s < 1.0f) ? env : 0.0f;
            int klo = (int)floorf(s * (float)(R - 1)) - 7;
            klo = max(0, min(klo, R - 16));
            const int k8 = min(klo & ~7, R - 24);  // first centre written (multiple of 8)
            uint32_t hi[12], lo[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                float g2[2];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const float diff = s - P.rbf_offset[k8 + 2 * i + h2];
                    g2[h2] = env * expf(P.coeff * (diff * diff)) * MT_RBF_SCALE;
                }
                const __half h0 = __float2half_rn(g2[0]), h1 = __float2half_rn(g2[1]);
                const __half l0 = __float2half_rn(g2[0] - __half2float(h0)), l1 = __float2half_rn(g2[1] - __half2float(h1));
                const __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
                hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
                lo[i] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            // wait until the MMAs of the previous tile have finished reading B, then rewrite this row
            mbar_wait(b_empty, (t & 1) ^ 1);
            // zero the row's 4 x 128 bytes; chunk order rotated by the row so a warp's stores spread over banks
#pragma unroll
            for (int sub = 0; sub < 4; ++sub) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t chunk = (uint32_t)(j ^ (r & 7));
                    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(b_smem + sub * MT_SUB + row_off + chunk * 16u),
                                 "r"(0u)
                                 : "memory");
                }
            }
            // three 16-byte chunks of 8 centres each, hi and lo planes, at their swizzled position
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int k = k8 + 8 * q;          // first centre of this chunk
                const uint32_t kb = (uint32_t)(k >> 6);
                const uint32_t chunk = (uint32_t)((k & 63) >> 3) ^ (uint32_t)(r & 7);
                const uint32_t addr = b_smem + kb * MT_SUB + row_off + chunk * 16u;
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hi[4 * q]), "r"(hi[4 * q + 1]),
                             "r"(hi[4 * q + 2]), "r"(hi[4 * q + 3])
                             : "memory");
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr + 2 * MT_SUB), "r"(lo[4 * q]),
                             "r"(lo[4 * q + 1]), "r"(lo[4 * q + 2]), "r"(lo[4 * q + 3])
                             : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async-proxy (MMA) reads
            mbar_arrive(b_full);

            // metadata for the epilogue warps (double buffered)
            const int mb = t & 1;
            mbar_wait(meta_empty(mb), ((t >> 1) & 1) ^ 1);
            Meta M = meta_at(meta_g + mb * MT_META_BYTES);
            M.srcoff[r] = src * 3 * F;
            M.tgt[r] = tgt;
            M.rhat[r] = make_float4(geo.y, geo.z, geo.w, 0.f);
            const unsigned endm = __ballot_sync(ADK_FULL_MASK, valid && (next_tgt != tgt));
            const unsigned valm = __ballot_sync(ADK_FULL_MASK, valid);
            if (lane == 0) {
                M.masks[warp - 2] = endm;
                M.masks[4 + warp - 2] = valm;
            }
            mbar_arrive(meta_full(mb));
        }
    } else {
        // ===================== epilogue: warps 6..13, two sets of four =====================
        const int q = warp & 3;            // TMEM lane quarter
        const int set = (warp - 6) >> 2;   // handles feature blocks cb with (cb & 1) == set
        const int fl = q * 32 + lane;      // feature inside the block
        // rows without in-edges never see a flush: give them the residual-only result up front
        for (int i = 0; i < n_atoms; ++i) {
            if (P.row_deg[a0 + i] == 0) {
                for (int cb = set; cb < 4; cb += 2) {
                    const int f = cb * MT_CB + fl;
                    const size_t t = (size_t)(a0 + i);
                    P.x_io[t * F + f] *= 0.70710678118654752440f;
                    for (int c = 0; c < 3; ++c)
                        P.vec_out[t * 3 * F + c * F + f] = has_vec ? P.vec_in[t * 3 * F + c * F + f] : 0.f;
                }
            }
        }
        float acc0[2] = {0.f, 0.f};
        float acc1[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
        float acc2[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
        float bias[2][3];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int g = 0; g < 3; ++g) bias[i][g] = P.b_rbf[g * F + (set + 2 * i) * MT_CB + fl];
        int n = 0;
        for (int t = 0; t < num_tiles; ++t) {
            const int mb = t & 1;
            mbar_wait(meta_full(mb), (t >> 1) & 1);
            const Meta M = meta_at(meta_g + mb * MT_META_BYTES);
            for (int u = 0; u < upt; ++u, ++n) {
                const int cb = unit_cb(u);
                if ((cb & 1) != set) continue;
                const int g = unit_g(u), slot = n & 3, i = cb >> 1;
                const int f = cb * MT_CB + fl;
                mbar_wait(d_full(slot), (n >> 2) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * MT_TILE;
                if (g == 0) epilogue_unit<0>(P, M, t_addr, f, bias[i][0], &acc0[i], has_vec);
                else if (g == 1) epilogue_unit<1>(P, M, t_addr, f, bias[i][1], acc1[i], has_vec);
                else epilogue_unit<2>(P, M, t_addr, f, bias[i][2], acc2[i], has_vec);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(d_empty(slot));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(meta_empty(mb));
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

extern "C" int adk_message_tc(const int32_t* atom_off, int B, const int32_t* sys_counts, const int32_t* row_deg,
                              const int32_t* e_src, const int32_t* e_tgt, const float* e_geo, const float* xh,
                              const float* vec_in, const void* w_rbf_split, float w_scale, const float* b_rbf,
                              const float* rbf_offset, int F, int R, int max_nbrs, float cutoff,
                              int envelope_exponent, float comp, float* x_io, float* vec_out, void* stream) {
    if (!atom_off || !sys_counts || !row_deg || !e_src || !e_tgt || !e_geo || !xh || !w_rbf_split || !b_rbf ||
        !rbf_offset || !x_io || !vec_out || B <= 0)
        return ADK_EINVAL;
    if (F != 4 * MT_CB || R != 128 || envelope_exponent < 1 || vec_in == vec_out) return ADK_EINVAL;
    alignas(64) CUtensorMap tmW;
    int rc = make_map_f16(&tmW, w_rbf_split, 2 * 3 * (uint64_t)F, (uint64_t)R, 64, MT_CB);
    if (rc != 0) return rc;
    MtParams P;
    P.atom_off = atom_off; P.sys_counts = sys_counts; P.row_deg = row_deg;
    P.e_src = e_src; P.e_tgt = e_tgt; P.e_geo = reinterpret_cast<const float4*>(e_geo);
    P.xh = xh; P.vec_in = vec_in; P.b_rbf = b_rbf; P.rbf_offset = rbf_offset;
    P.F = F; P.R = R; P.k_nbrs = max_nbrs;
    P.inv_cutoff = (float)(1.0 / (double)cutoff);
    const double spacing = 1.0 / (double)(R - 1);
    P.coeff = (float)(-0.5 / (spacing * spacing));
    const double p = (double)envelope_exponent;
    P.env_p = envelope_exponent;
    P.env_a = (float)(-(p + 1) * (p + 2) / 2);
    P.env_b = (float)(p * (p + 2));
    P.env_c = (float)(-p * (p + 1) / 2);
    P.acc_scale = (1.0f + comp) / (MT_RBF_SCALE * w_scale);
    P.x_io = x_io; P.vec_out = vec_out;
    message_tc_kernel<<<B, MT_THREADS, MT_SMEM_BYTES, adk::as_stream(stream)>>>(tmW, P);
    ADK_LAUNCH_CHECK();
    return 0;
}

int adk_message_tc_set_attrs() {
    return (int)cudaFuncSetAttribute(message_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MT_SMEM_BYTES);
}
