// K3 (tcgen05 variant, the default for full batches): fused RBF featurisation + rbf_proj on the 5th-generation
// tensor cores + PaiNN message + CSR segmented reduction, with everything an edge touches resident on the SM.
//
// Reference arithmetic: models/gemnet_oc/layers/radial_basis.py:235-244 (Gaussian basis x polynomial envelope),
// models/painn/painn_denoising.py:534-567 (rbf_proj, message, aggregation), 443-445 (residual).
//
// CTA = (one adsorbate+slab system, slice of 64 features).  rbf_proj is 68 % of the reference's FLOPs:
//   rbfh[e][g][f] = sum_k W[g*F + f][k] * rbf[e][k] + b      (g = the three message groups m1 | m2 | m3)
// It is evaluated as  D[feature row][edge column] = W_tile[128 rows][K] . rbf_tile[128 columns][K]^T  by
// tcgen05.mma (kind::f16, fp16x2-split operands: Wh.Bl + Wl.Bh + Wh.Bh, fp32 accumulation in TMEM), so a TMEM
// LANE is a feature and a COLUMN an in-edge: an epilogue thread owns one feature, reads the rbfh of successive
// edges from its lane, multiplies with the source atom's features -- staged in shared memory for the whole
// system, read as 32 consecutive words per warp, i.e. conflict free -- and keeps the segmented sums of the
// target rows in registers.  No atomics, no shuffles, deterministic; per-edge tensors never exist in HBM.
//
// Two phases per CTA (one weight tile -- the MMA's A operand, resident in tensor memory -- and one set of sources in
// shared memory at a time):
//   A: W rows = [m1 slice | m3 slice]   lanes 0-63 accumulate dx, lanes 64-127 the m3 * r_hat part of dvec
//   B: W rows = [m2 slice | m2 slice]   lanes 0-63 take row slots 0-3 of a tile, lanes 64-127 slots 4-7; the
//                                       source is p2 = xh2 * vec (3 components), formed while staging
// A tile has 128 columns = 8 target rows x 16 in-edges: the rows of a "row group" (rows ranked by degree) and,
// of each, the t-th chunk of 16 edges in distance order.  Edges of equal rank have similar distances, so the
// Gaussian windows (16 centres each) of a tile overlap and only the 16..64 centres of their union enter the
// MMA (K = 128 dense otherwise); a tile whose union exceeds 64 centres is issued as two k-halves into the same
// accumulator.  Warp roles: warp 0 = MMA issue + TMEM owner; two generator warpgroups on alternate tiles (one
// column per thread: CSR record -> basis values -> fp16x2 -> UMMA 128-byte-swizzled operand row, plus the per-column
// metadata); four epilogue warpgroups in two teams on alternate tiles, each warpgroup taking half of a tile's eight
// rows and keeping their sums in registers; the epilogue warpgroups also write the weight rows to TMEM at the start
// of a phase (tcgen05.st, thread = lane = weight row).
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

using namespace adk::tc;

constexpr int T5_EPI_WG = 4;                                  // epilogue warpgroups
constexpr int T5_TEAMS = 2;                                   // teams of T5_EPI_WG / T5_TEAMS warpgroups; team = tile parity
constexpr int T5_TEAM_WG = T5_EPI_WG / T5_TEAMS;              // the warpgroups of a team split a tile's 8 rows
// warpgroup 0: warp 0 = TMA + MMA issue (3 idle warps complete the group), warpgroup 1: generators, then the epilogue
// warpgroups; register budgets are re-dealt per role with setmaxnreg: 768 x 80 = 128 x 40 + 128 x 72 + 512 x 88 (+ slack)
constexpr int T5_GEN_GROUPS = 2;                              // generator warpgroups: group g makes the tiles ti % 2 == g
constexpr int T5_THREADS = 128 * (1 + T5_GEN_GROUPS + T5_EPI_WG);
constexpr int T5_GEN_WARP0 = 4, T5_EPI_WARP0 = 4 + 4 * T5_GEN_GROUPS;
constexpr int T5_SF = 64;                                     // features per CTA slice
constexpr int T5_TILE = 128;                                  // columns (edge slots) per tile = MMA N
constexpr int T5_ROWS = 8;                                    // target rows per tile
constexpr int T5_SLOT = 16;                                   // columns per row
// T5_WTMEM (default): the weight tile is the TMEM-resident A operand of the MMA (columns 384..511: hi plane | lo
// plane, a lane = a weight row, two fp16 per column), written once per phase with tcgen05.st by one epilogue
// warpgroup.  The tensor core then fetches only the rbf tile from shared memory (half the operand wavefronts of the
// shared-memory variant), no weight TMA / swizzle is needed, and the 64 KB the weights occupied hold two more operand
// buffers.  Price: three accumulator slots instead of four.  -DT5_WTMEM=0 builds the round-2 shared-memory variant.
#ifndef T5_WTMEM
#define T5_WTMEM 1
#endif
constexpr int T5_DSLOTS = T5_WTMEM ? 3 : 4;                   // TMEM accumulator ring of 128-column slots
// operand-tile buffers: a kernel template parameter.  T5_WTMEM: 4 (two per generator group) while the system's
// sources leave room for them (up to 111 atoms at F = 512), 2 for larger systems (up to 200 atoms); else 2.
constexpr int T5_BBUFS_MAX = T5_WTMEM ? 4 : 2;
constexpr uint32_t T5_W_TMEM_COL = 384;                       // T5_WTMEM: first column of the weight planes (64 + 64)
constexpr int T5_MAX_TILES = 512;
constexpr int T5_MAX_ATOMS = T5_WTMEM ? 200 : 128;
constexpr uint32_t T5_W_KBLOCK = 128 * 128;                   // [128 rows][64 centres] fp16, 128-byte rows
constexpr uint32_t T5_W_PLANE = 2 * T5_W_KBLOCK;
constexpr uint32_t T5_W_BYTES = 2 * T5_W_PLANE;               // hi + lo: 64 KB
constexpr uint32_t T5_B_PLANE = 128 * 128;                    // [128 columns][64 centres] fp16
constexpr uint32_t T5_B_BYTES = 2 * T5_B_PLANE;               // hi + lo: 32 KB
// r_hat of a block of eight columns = 24 floats [x0..x7 | y0..y7 | z0..z7]: the generator lanes' stores are conflict
// free (stride 24 words across blocks) and an edge block reads its unit vectors with six 16-byte loads
constexpr uint32_t T5_META_SRC = 0, T5_META_RHAT = 512, T5_META_CNT = 2048, T5_META_ROW = 2080;
constexpr uint32_t T5_META_BYTES = 2176;
constexpr uint32_t T5_SRC_PITCH_A = 2 * T5_SF * 4;            // [xh1 | xh3] per atom
constexpr uint32_t T5_SRC_PITCH_B = 3 * T5_SF * 4;            // p2 = xh2 * vec[x|y|z] per atom
constexpr float T5_RBF_SCALE = 1024.0f;

// optional pipeline trace (debug builds: -DT5_TRACE): clock64 stamps of CTA (0,0,0), phase 0, [tile][16 events]
#ifdef T5_TRACE
#ifndef T5_TRACE_X
#define T5_TRACE_X 0   // the traced CTA: slice T5_TRACE_X of system T5_TRACE_Y (first-wave CTAs see cold caches)
#define T5_TRACE_Y 0
#endif
__device__ long long g_t5_trace[T5_MAX_TILES * 16];
#define T5_STAMP(ti, ev) do { if (blockIdx.x == T5_TRACE_X && blockIdx.y == T5_TRACE_Y && blockIdx.z == 0 && phase == 0 && lane == 0) g_t5_trace[(ti) * 16 + (ev)] = clock64(); } while (0)
#define T5_CTA_STAMP(ev) do { if (blockIdx.x == T5_TRACE_X && blockIdx.y == T5_TRACE_Y && blockIdx.z == 0 && threadIdx.x == 0) g_t5_trace[(T5_MAX_TILES - 1) * 16 + (ev)] = clock64(); } while (0)
#else
#define T5_STAMP(ti, ev) do { } while (0)
#define T5_CTA_STAMP(ev) do { } while (0)
#endif

struct T5Params {
    const int32_t* atom_off;
    const int32_t* row_sel;     // optional [N]: 1 = compute the row, 2 = pass vec through (x untouched), 0 = leave untouched
    const int32_t* row_start;
    const int32_t* row_deg;
    const int32_t* e_src;
    const float4* e_geo;
    const float* xh;
    const float* vec_in;
    const float* b_rbf;
    const float* rbf_offset;
    const __half* w_split;      // [2 planes][3F][R] fp16 (hi, lo) of s_w * rbf_proj.weight
    int F, R, n_max;
    float inv_cutoff, coeff_sqrt, env_a, env_b, env_c;
    int env_p;
    float acc_scale;
    float* x_io;
    float* vec_out;
    __half* vsplit;
    int64_t vsplit_plane;
    float vsplit_scale;
    uint32_t* status;
};

__host__ __device__ inline size_t t5_fixed_bytes(int nbufs) {
    return (T5_WTMEM ? 0 : T5_W_BYTES) + nbufs * T5_B_BYTES + T5_DSLOTS * T5_META_BYTES + 512 /*barriers, item meta, windows*/ +
           T5_MAX_TILES * 4 + 2 * T5_MAX_ATOMS /*row order*/ + 128 /*mu*/ * 4 + 2 * 4 * T5_MAX_ATOMS /*row start, degree*/;
}
__host__ __device__ inline size_t t5_smem_bytes(int n_max, int nbufs) {
    return 1024 /*alignment slack*/ + t5_fixed_bytes(nbufs) + (size_t)(n_max + 1) * T5_SRC_PITCH_B;
}
// the deepest operand ring whose shared memory still fits a system of n_max atoms (0: none does)
inline int t5_pick_bufs(int n_max) {
    for (int nb = T5_BBUFS_MAX; nb >= 2; nb -= 2)
        if (t5_smem_bytes(n_max, nb) <= 227 * 1024) return nb;
    return 0;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ int4 lds_i4(uint32_t addr) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// shared-memory matrix descriptor (see tc_common.cuh umma_desc) split into its variable low word and constant high word
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo) {
    constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
    return ((uint64_t)HI << 32) | lo;
}

// tile table entry: row group | chunk << 8 | epilogue warpgroup << 16 | last-chunk-of-its-group << 24
__device__ __forceinline__ int tile_group(uint32_t t) { return (int)(t & 0xffu); }
__device__ __forceinline__ int tile_chunk(uint32_t t) { return (int)((t >> 8) & 0xffu); }
__device__ __forceinline__ int tile_wg(uint32_t t) { return (int)((t >> 16) & 0xffu); }
__device__ __forceinline__ bool tile_last(uint32_t t) { return (t >> 24) != 0u; }

// ---- epilogue bodies -------------------------------------------------------------------------------------
// NE (4 or 8) consecutive in-edges of one target row, two per packed fp32x2 instruction.  All shared-memory reads of
// the block (source offsets, r_hat, source features: plain loads, so the compiler may issue them back to back) come
// before the arithmetic; an earlier version that loaded edge by edge spent ~65 cycles of exposed latency per four edges.
// `v` holds the accumulator columns of this thread's feature.  Padding columns point at a zero source row, so no
// per-edge predicates are needed.  p[c] = (sum over even edges, sum over odd edges) of output component c.
#ifndef T5_EXP
#define T5_EXP 0   // ablation builds (scripts/build_exp_libs.sh): 1 = no TMEM loads, 2 = no source loads, 3 = empty epilogue,
                   // 4 = no MMAs, 5 = empty epilogue + no generator arithmetic / operand stores, 6 = 5 + no MMAs
#endif
template <int MODE, int NE>   // MODE 0: dx (m1), 1: m3 * r_hat, 2: p2 * m2
__device__ __forceinline__ void edge_block(const uint32_t* v, const uint8_t* meta, int col, const uint8_t* src_lane,
                                           float2 scale2, float2 bias2, float2* p) {
    int so[NE];
#pragma unroll
    for (int i = 0; i < NE / 4; ++i) {
        const int4 t = *reinterpret_cast<const int4*>(meta + T5_META_SRC + (size_t)(col + 4 * i) * 4);
        so[4 * i] = t.x; so[4 * i + 1] = t.y; so[4 * i + 2] = t.z; so[4 * i + 3] = t.w;
    }
    constexpr int W = MODE == 2 ? 3 : 1;
    float x[NE][W];
#pragma unroll
    for (int e = 0; e < NE; ++e)
#pragma unroll
        for (int c = 0; c < W; ++c)
            x[e][c] = (T5_EXP == 2) ? __int_as_float(so[e] + c) : *reinterpret_cast<const float*>(src_lane + so[e] + c * T5_SF * 4);
    float4 rx[NE / 4], ry[NE / 4], rz[NE / 4];
    if (MODE == 1) {
        const float4* ra = reinterpret_cast<const float4*>(meta + T5_META_RHAT + (size_t)(col >> 3) * 96 + (size_t)(col & 7) * 4);
#pragma unroll
        for (int i = 0; i < NE / 4; ++i) {
            rx[i] = ra[i];
            ry[i] = ra[2 + i];
            rz[i] = ra[4 + i];
        }
    }
#pragma unroll
    for (int h = 0; h < NE / 2; ++h) {
        const float2 r2 = adk::fma2(make_float2(__uint_as_float(v[2 * h]), __uint_as_float(v[2 * h + 1])), scale2, bias2);
        if (MODE == 0) {
            p[0] = adk::fma2(make_float2(x[2 * h][0], x[2 * h + 1][0]), r2, p[0]);
        } else if (MODE == 1) {
            const float2 m2 = adk::fma2(make_float2(x[2 * h][0], x[2 * h + 1][0]), r2, make_float2(0.f, 0.f));
            const float4 qx = rx[h >> 1], qy = ry[h >> 1], qz = rz[h >> 1];
            p[0] = adk::fma2(m2, (h & 1) ? make_float2(qx.z, qx.w) : make_float2(qx.x, qx.y), p[0]);
            p[1] = adk::fma2(m2, (h & 1) ? make_float2(qy.z, qy.w) : make_float2(qy.x, qy.y), p[1]);
            p[2] = adk::fma2(m2, (h & 1) ? make_float2(qz.z, qz.w) : make_float2(qz.x, qz.y), p[2]);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) p[c] = adk::fma2(make_float2(x[2 * h][c], x[2 * h + 1][c]), r2, p[c]);
        }
    }
}

// One row slot (16 columns = up to 16 in-edges of one target row) in blocks of NE edges (warp-uniform trip count).
template <int MODE>
__device__ __forceinline__ void slot_body(const uint32_t* v, int cnt, const uint8_t* meta, int col0, const uint8_t* src_lane,
                                          float2 scale2, float2 bias2, float2* p) {
    constexpr int NE = 8;   // eight edges per block: all shared-memory reads of the block are issued before its arithmetic
#pragma unroll
    for (int b = 0; b < 16 / NE; ++b)
        if (cnt > b * NE) edge_block<MODE, NE>(v + b * NE, meta, col0 + b * NE, src_lane, scale2, bias2, p);
}

// NS consecutive row slots [s0, s0 + NS) of one tile for this thread (the warpgroups of an epilogue team split a
// tile's eight rows between them, so the row sums acc[NS][W] are indexed statically).  One TMEM load in flight per
// warp: the other warps of the scheduler cover its latency (a second register buffer spilled).
template <int MODE, int NS>
__device__ __forceinline__ void tile_body(uint32_t t_addr, const uint8_t* meta, const uint8_t* src_lane, float2 scale2,
                                          float2 bias2, int s0, float* acc) {
    constexpr int W = MODE == 0 ? 1 : 3;
    const int* cnts = reinterpret_cast<const int*>(meta + T5_META_CNT);
    int cnt[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) cnt[i] = cnts[s0 + i];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        if (cnt[i] > 0) {   // warp-uniform
            uint32_t v[16];
            if (T5_EXP == 1) {
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = t_addr + k;
            } else {
                tmem_ld16_async(t_addr + (uint32_t)(s0 + i) * T5_SLOT, v);
                tmem_ld_wait();
            }
            tmem_pin16(v);
            float2 p[W];
#pragma unroll
            for (int c = 0; c < W; ++c) p[c] = make_float2(0.f, 0.f);
            slot_body<MODE>(v, cnt[i], meta, (s0 + i) * T5_SLOT, src_lane, scale2, bias2, p);
#pragma unroll
            for (int c = 0; c < W; ++c) acc[i * W + c] += p[c].x + p[c].y;
        }
    }
}

// Sources of one phase into shared memory: A = [atom][xh1 (64) | xh3 (64)], B = [atom][p2x | p2y | p2z] with
// p2c = xh2 * vec_c; row n = zeros (what padding columns read).  Staging is latency bound (a few 16-byte items per
// thread, each a round trip to L2 / HBM), so every thread issues the loads of BATCH items before it stores any:
// `tid` of `nthreads` callers, items tid + k * nthreads from `first` on.
__device__ __forceinline__ float4 stage_load_a(const T5Params& P, int i, int total, int a0, int n, int f0) {
    constexpr int Q = T5_SF / 4;   // float4 per 64-feature row
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < total) {
        const int j = i / (2 * Q), r = i - j * 2 * Q;
        const int g = r >= Q ? 2 : 0, q4 = r & (Q - 1);
        if (j < n) v = *reinterpret_cast<const float4*>(P.xh + (size_t)(a0 + j) * 3 * P.F + g * P.F + f0 + 4 * q4);
    }
    return v;
}
template <int BATCH>
__device__ __forceinline__ void stage_sources(const T5Params& P, float* s_src, int phase, int a0, int n, int f0, int tid,
                                              int nthreads, int first = 0) {
    float4* s_src4 = reinterpret_cast<float4*>(s_src);
    constexpr int Q = T5_SF / 4;
    const int F = P.F;
    if (phase == 0) {
        const int total = (n + 1) * 2 * Q;
        for (int i0 = first + tid; i0 < total; i0 += BATCH * nthreads) {
            float4 v[BATCH];
#pragma unroll
            for (int k = 0; k < BATCH; ++k) v[k] = stage_load_a(P, i0 + k * nthreads, total, a0, n, f0);
#pragma unroll
            for (int k = 0; k < BATCH; ++k)
                if (i0 + k * nthreads < total) s_src4[i0 + k * nthreads] = v[k];
        }
    } else {
        const int total = (n + 1) * 3 * Q;
        for (int i0 = first + tid; i0 < total; i0 += BATCH * nthreads) {
            float4 a[BATCH], w[BATCH];
#pragma unroll
            for (int k = 0; k < BATCH; ++k) {
                const int i = i0 + k * nthreads;
                a[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                w[k] = a[k];
                if (i < total) {
                    const int j = i / (3 * Q), r = i - j * 3 * Q;
                    const int c = r / Q, q4 = r - c * Q;
                    if (j < n) {
                        a[k] = *reinterpret_cast<const float4*>(P.xh + (size_t)(a0 + j) * 3 * F + F + f0 + 4 * q4);
                        w[k] = *reinterpret_cast<const float4*>(P.vec_in + ((size_t)(a0 + j) * 3 + c) * F + f0 + 4 * q4);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < BATCH; ++k)
                if (i0 + k * nthreads < total)
                    s_src4[i0 + k * nthreads] = make_float4(a[k].x * w[k].x, a[k].y * w[k].y, a[k].z * w[k].z, a[k].w * w[k].w);
        }
    }
}
#if T5_WTMEM
// weight rows of one phase: global -> registers (eight 16-byte loads), registers -> TMEM (see the kernel's set-up)
__device__ __forceinline__ void load_weight_rows(const T5Params& P, uint4* w4, int phase, int warp, int lane, int f0) {
    const int r = (warp & 3) * 32 + lane, wg = (warp - T5_EPI_WARP0) >> 2;
    const int g = phase == 0 ? (r < 64 ? 0 : 2) : 1;
    const uint4* wrow = reinterpret_cast<const uint4*>(P.w_split + (((size_t)(wg >> 1) * 3 + g) * P.F + f0 + (r & 63)) * P.R) + (wg & 1) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) w4[i] = __ldg(wrow + i);
}
__device__ __forceinline__ void store_weight_rows(uint32_t tmem_base, const uint4* w4, int warp) {
    const int wg = (warp - T5_EPI_WARP0) >> 2;
    const uint32_t t_w = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + T5_W_TMEM_COL + (uint32_t)((wg >> 1) * 64 + (wg & 1) * 32);
    tmem_st16(t_w, reinterpret_cast<const uint32_t*>(w4));
    tmem_st16(t_w + 16u, reinterpret_cast<const uint32_t*>(w4 + 4));
    tmem_st_wait();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
#endif
constexpr int T5_STAGE_PRE = 3;   // phase-A items per thread whose loads are in flight during the CTA's set-up

template <int T5_BBUFS>
__global__ void __launch_bounds__(T5_THREADS, 1)
message_t5_kernel(const __grid_constant__ CUtensorMap tmW, T5Params P) {   // (tmW: shared-memory weight variant only)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t w_smem = base;   // (shared-memory weight variant only)
    const uint32_t b_smem = base + (T5_WTMEM ? 0u : T5_W_BYTES);
    const uint32_t meta_smem = b_smem + T5_BBUFS * T5_B_BYTES;
    const uint32_t ctl = meta_smem + T5_DSLOTS * T5_META_BYTES;     // 512-byte control block
    uint8_t* ctl_g = gbase + (ctl - base);
    const uint32_t w_full = ctl;
    static_assert(T5_BBUFS <= 4 && T5_DSLOTS <= 4, "control block layout");
    static_assert(T5_BBUFS % T5_GEN_GROUPS == 0, "each generator group owns T5_BBUFS / T5_GEN_GROUPS operand buffers");
    auto b_full = [&](int s) { return ctl + 8u + 8u * s; };
    auto b_empty = [&](int s) { return ctl + 40u + 8u * s; };
    auto d_full = [&](int s) { return ctl + 72u + 8u * s; };
    auto d_empty = [&](int s) { return ctl + 104u + 8u * s; };
    auto m_full = [&](int s) { return ctl + 136u + 8u * s; };
    const uint32_t tmem_slot = ctl + 168u;
    int* s_item = reinterpret_cast<int*>(ctl_g + 176);             // [T5_BBUFS][4]: kbase, nks, last, -
    int* s_win = reinterpret_cast<int*>(ctl_g + 240);              // [groups 2][parity 2][4 warps][2]: (kmin, kmax)
    int* s_ntiles = reinterpret_cast<int*>(ctl_g + 496);
    uint32_t* s_tiles = reinterpret_cast<uint32_t*>(ctl_g + 512);
    int16_t* s_order = reinterpret_cast<int16_t*>(ctl_g + 512 + T5_MAX_TILES * 4);
    int* s_rs = reinterpret_cast<int*>(ctl_g + 512 + T5_MAX_TILES * 4 + 2 * T5_MAX_ATOMS + 512);  // by rank: first CSR slot
    int* s_rd = s_rs + T5_MAX_ATOMS;                                                             // by rank: in-degree
    const uint32_t src_smem = ctl + 512u + T5_MAX_TILES * 4u + 2u * T5_MAX_ATOMS + 512u + 8u * T5_MAX_ATOMS;
    float* s_src = reinterpret_cast<float*>(gbase + (src_smem - base));

    const int warp = adk::warp_id(), lane = adk::lane_id();
    const int f0 = blockIdx.x * T5_SF;
    const int b = blockIdx.y;
    const int a0 = P.atom_off[b], n = P.atom_off[b + 1] - a0;
    const int F = P.F, R = P.R;
    const bool has_vec = P.vec_in != nullptr;
    if (n > P.n_max) return;   // larger than the staging area this launch was sized for: another kernel owns it (row_sel == 0 there)
    T5_CTA_STAMP(0);
    // phase A's sources: the first T5_STAGE_PRE items of every thread are requested now and stored after the set-up
    // below (row ranking, tile list), which hides one of the two staging round trips behind the other work
    float4 pre[T5_STAGE_PRE];
#pragma unroll
    for (int k = 0; k < T5_STAGE_PRE; ++k)
        pre[k] = stage_load_a(P, (int)threadIdx.x + k * T5_THREADS, (n + 1) * 2 * (T5_SF / 4), a0, n, f0);


    // ---- one-time setup ----------------------------------------------------------------------------------
    if (warp == 0) {
        if (lane == 0) {
#if !T5_WTMEM
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
            mbar_init(w_full, 1);
#endif
            for (int s = 0; s < T5_BBUFS; ++s) { mbar_init(b_full(s), 4); mbar_init(b_empty(s), 1); }
            for (int s = 0; s < T5_DSLOTS; ++s) { mbar_init(d_full(s), 1); mbar_init(d_empty(s), 4 * T5_TEAM_WG); mbar_init(m_full(s), 4); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // selected rows ranked by (degree desc, index asc): a permutation; similar rows share a tile, and the first row
    // of a group bounds the group's number of chunks
    for (int r = threadIdx.x; r < n; r += T5_THREADS) {
        if (P.row_sel && P.row_sel[a0 + r] != 1) continue;
        const int dr = P.row_deg[a0 + r];
        int rank = 0;
        for (int u = 0; u < n; ++u) {
            if (P.row_sel && P.row_sel[a0 + u] != 1) continue;
            const int du = P.row_deg[a0 + u];
            rank += (du > dr || (du == dr && u < r)) ? 1 : 0;
        }
        s_order[rank] = (int16_t)r;
        s_rs[rank] = P.row_start[a0 + r];
        s_rd[rank] = dr;
    }
    if (warp == 1) {   // number of selected rows
        int c = 0;
        for (int r = lane; r < n; r += 32) c += (!P.row_sel || P.row_sel[a0 + r] == 1) ? 1 : 0;
        c = __reduce_add_sync(ADK_FULL_MASK, c);
        if (lane == 0) s_ntiles[1] = c;
    }
    // pass-through rows (row_sel == 2): this slice of vec_in is copied to vec_out (+ its operand planes), so that
    // everything downstream of an unselected row still sees values of the network's own scale
    if (P.row_sel && blockIdx.z == 0) {
        for (int i = threadIdx.x; i < n * 3 * T5_SF; i += T5_THREADS) {
            const int r = i / (3 * T5_SF), rem = i - r * 3 * T5_SF;
            if (P.row_sel[a0 + r] != 2) continue;
            const int c = rem / T5_SF, f = rem - c * T5_SF;
            const size_t off = ((size_t)(a0 + r) * 3 + c) * P.F + f0 + f;
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
#pragma unroll
    for (int k = 0; k < T5_STAGE_PRE; ++k)
        if ((int)threadIdx.x + k * T5_THREADS < (n + 1) * 2 * (T5_SF / 4))
            reinterpret_cast<float4*>(s_src)[threadIdx.x + k * T5_THREADS] = pre[k];
    stage_sources<2>(P, s_src, 0, a0, n, f0, (int)threadIdx.x, T5_THREADS, T5_STAGE_PRE * T5_THREADS);   // (systems > 83 atoms)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (threadIdx.x == 0) {
        // tile list: the row groups of a "super group" (one per epilogue team) alternate, chunk by chunk, so that the
        // teams work on different tiles at the same time; gridDim.z > 1 (a handful of systems only): the super groups
        // of a system are dealt to several CTAs.  entry: row group | chunk << 8 | team << 16 | last chunk << 24
        const int nsel0 = s_ntiles[1];
        const int ngroups = (nsel0 + T5_ROWS - 1) / T5_ROWS;
        int nt = 0;
        for (int sg = (int)blockIdx.z; sg * T5_TEAMS < ngroups; sg += (int)gridDim.z) {
            int tmax = 0, tg[T5_TEAMS];
            for (int w = 0; w < T5_TEAMS; ++w) {
                const int g = sg * T5_TEAMS + w;
                tg[w] = g < ngroups ? max(1, (s_rd[g * T5_ROWS] + T5_SLOT - 1) / T5_SLOT) : 0;
                tmax = max(tmax, tg[w]);
            }
            for (int t = 0; t < tmax; ++t)
                for (int w = 0; w < T5_TEAMS; ++w)
                    if (t < tg[w] && nt < T5_MAX_TILES)
                        s_tiles[nt++] = (uint32_t)(sg * T5_TEAMS + w) | ((uint32_t)t << 8) | ((uint32_t)w << 16) |
                                        ((t == tg[w] - 1) ? (1u << 24) : 0u);
        }
        *s_ntiles = nt;
    }
    __syncthreads();
    const int ntiles = *s_ntiles;
    const int nsel = s_ntiles[1];
    const int nphases = has_vec ? 2 : 1;
    T5_CTA_STAMP(1);

    // Every role runs the same phase skeleton -- stage the sources, __syncthreads, its loop over the tiles,
    // __syncthreads -- inside its own branch, which opens with its own setmaxnreg (so ptxas applies that budget there).
    if (warp < T5_GEN_WARP0) {
        // ===================== warpgroup 0: TMA (weights) + MMA issue =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        uint32_t item_par = 0u;   // bit b: parity of the number of operand tiles consumed from buffer b
        for (int phase = 0; phase < nphases; ++phase) {
#if !T5_WTMEM
            if (warp == 0 && lane == 0) {
                mbar_expect_tx(w_full, T5_W_BYTES);
                const int g_lo = phase == 0 ? 0 : 1, g_hi = phase == 0 ? 2 : 1;   // row halves: [m1 | m3] or [m2 | m2]
#pragma unroll
                for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint32_t dst = w_smem + pl * T5_W_PLANE + kb * T5_W_KBLOCK;
                        tma_load_2d(dst, &tmW, w_full, kb * 64, pl * 3 * F + g_lo * F + f0);
                        tma_load_2d(dst + 64 * 128, &tmW, w_full, kb * 64, pl * 3 * F + g_hi * F + f0);
                    }
            }
#endif
            // (phase A's sources were staged during the set-up, phase B's are staged by the other roles' 768 threads:
            // this warpgroup runs on 40 registers)
            __syncthreads();
            T5_CTA_STAMP(2 + 3 * phase);
            if (warp == 0 && lane == 0) {
                const uint32_t tile0 = (uint32_t)phase * (uint32_t)ntiles;
                const uint32_t idesc = (1u << 4) | ((uint32_t)(T5_TILE >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
#if T5_WTMEM
                // the weight planes were written to TMEM by an epilogue warpgroup before the barrier above
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t wh_col = tmem_base + T5_W_TMEM_COL;
#else
                const uint32_t wh_lo = desc_lo(w_smem);
                mbar_wait(w_full, phase & 1);
#endif
                for (int ti = 0; ti < ntiles; ++ti) {
                    const uint32_t tau = tile0 + ti;
                    const int ds = tau % T5_DSLOTS;
                    mbar_wait(d_empty(ds), ((tau / T5_DSLOTS) & 1) ^ 1);
                    T5_STAMP(ti, 5);
                    const uint32_t d_tmem = tmem_base + (uint32_t)ds * T5_TILE;
                    bool first = true, last = false;
                    const int buf = ti % T5_BBUFS;
                    while (!last) {
                        mbar_wait(b_full(buf), (item_par >> buf) & 1u);
                        T5_STAMP(ti, 6);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const int kbase = s_item[buf * 4], nks = s_item[buf * 4 + 1];
                        last = s_item[buf * 4 + 2] != 0;
                        // descriptors differ in their low word only: (address >> 4) | constant bits
                        const uint32_t bh_lo = desc_lo(b_smem + buf * T5_B_BYTES), bl_lo = bh_lo + (T5_B_PLANE >> 4);
                        // All descriptor words first (independent integer work), then the tcgen05.mma of the item
                        // in straight-line code per k-step count: the issuing thread, not the tensor pipe, paces a
                        // tile (~100 cycles per MMA when descriptor arithmetic and branches sit between them).
                        const uint32_t acc0 = first ? 0u : 1u;
                        // corrections first (while the accumulator is small), then the hi x hi products
#if T5_WTMEM
                        // A operand in TMEM: k-step s of the window starts (kbase + 16 s) / 2 columns into a plane
                        const uint32_t a_col = wh_col + (uint32_t)(kbase >> 1);
#define T5_MMA_CORR(s, acc)                                                                                         \
    umma_f16_ts(d_tmem, a_col + 8u * (s), mk_desc(bl_lo + 2 * (s)), idesc, acc);                                    \
    umma_f16_ts(d_tmem, a_col + 64u + 8u * (s), mk_desc(bh_lo + 2 * (s)), idesc, 1u);
#define T5_MMA_MAIN(s) umma_f16_ts(d_tmem, a_col + 8u * (s), mk_desc(bh_lo + 2 * (s)), idesc, 1u);
#else
                        uint32_t a_lo[4];
#pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            const int k = kbase + 16 * s;
                            a_lo[s] = wh_lo + (uint32_t)(k >> 6) * (T5_W_KBLOCK >> 4) + (uint32_t)((k >> 3) & 7);
                        }
#define T5_MMA_CORR(s, acc)                                                                                         \
    umma_f16(d_tmem, mk_desc(a_lo[s]), mk_desc(bl_lo + 2 * (s)), idesc, acc);                                      \
    umma_f16(d_tmem, mk_desc(a_lo[s] + (T5_W_PLANE >> 4)), mk_desc(bh_lo + 2 * (s)), idesc, 1u);
#define T5_MMA_MAIN(s) umma_f16(d_tmem, mk_desc(a_lo[s]), mk_desc(bh_lo + 2 * (s)), idesc, 1u);
#endif
                        if (T5_EXP == 4 || T5_EXP == 6) {
                        } else if (nks == 4) {
                            T5_MMA_CORR(0, acc0) T5_MMA_CORR(1, 1u) T5_MMA_CORR(2, 1u) T5_MMA_CORR(3, 1u)
                            T5_MMA_MAIN(0) T5_MMA_MAIN(1) T5_MMA_MAIN(2) T5_MMA_MAIN(3)
                        } else if (nks == 3) {
                            T5_MMA_CORR(0, acc0) T5_MMA_CORR(1, 1u) T5_MMA_CORR(2, 1u)
                            T5_MMA_MAIN(0) T5_MMA_MAIN(1) T5_MMA_MAIN(2)
                        } else if (nks == 2) {
                            T5_MMA_CORR(0, acc0) T5_MMA_CORR(1, 1u)
                            T5_MMA_MAIN(0) T5_MMA_MAIN(1)
                        } else {
                            T5_MMA_CORR(0, acc0)
                            T5_MMA_MAIN(0)
                        }
#undef T5_MMA_CORR
#undef T5_MMA_MAIN
                        umma_commit(b_empty(buf));
                        first = false;
                        item_par ^= 1u << buf;
                    }
                    umma_commit(d_full(ds));
                    T5_STAMP(ti, 7);
                }
            }
            T5_CTA_STAMP(3 + 3 * phase);
            __syncthreads();   // everybody is done with this phase's weights, sources and partial vec_out
            T5_CTA_STAMP(4 + 3 * phase);
        }
    } else if (warp < T5_EPI_WARP0) {
        // ===================== generator warpgroups, one column each =====================
        // A tile's operand rows are one warp's serial work of a few hundred instructions (~4000 cycles at single-warp
        // issue rates): two groups make alternate tiles, each into its own operand buffer.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        const int gg = (warp - T5_GEN_WARP0) >> 2;      // generator group = its operand buffer
        const int gw = (warp - T5_GEN_WARP0) & 3;
        const int c = gw * 32 + lane;
        const int slot = c >> 4, j = c & 15;
        const uint32_t row_off = (uint32_t)(c >> 3) * 1024u + (uint32_t)(c & 7) * 128u;
        const float rmax = (float)(R - 1);
        const float h0 = 1.0f / rmax;   // centre spacing (the host checked that rbf_offset is linspace(0, 1, R))
        // per operand buffer b (tile ti uses buffer ti % T5_BBUFS; this group owns those with b % T5_GEN_GROUPS == gg):
        uint32_t item_par = 0u;          // bit b: parity of the number of operand tiles written to it
        uint32_t dirty = 0xffffffffu;    // byte b: 16-byte chunks of MY row of it that may hold non-zeros (all, at first)
        // the CSR record of this column for tile ti (padding: source row n = zeros)
        auto fetch = [&](int ti, int& src, float4& geo, int& deg_out, int& row_out) {
            src = n;
            geo = make_float4(0.f, 0.f, 0.f, 0.f);
            deg_out = 0;
            row_out = -1;
            if (ti < ntiles) {
                const uint32_t tl = s_tiles[ti];
                const int ridx = tile_group(tl) * T5_ROWS + slot;
                if (ridx < nsel) {
                    deg_out = s_rd[ridx];
                    row_out = a0 + (int)s_order[ridx];
                    const int rank = tile_chunk(tl) * T5_SLOT + j;
                    if (rank < deg_out) {
                        const int e = s_rs[ridx] + rank;
                        src = P.e_src[e] - a0;
                        geo = P.e_geo[e];
                    }
                }
            }
        };
        for (int phase = 0; phase < nphases; ++phase) {
            // the record of this group's NEXT tile is in flight while the current one is worked on; the first one is
            // requested before the staging and the barrier (a global round trip costs ~3 000 cycles in a busy machine)
            int src1, deg1, row1;
            float4 geo1;
            fetch(gg, src1, geo1, deg1, row1);
            if (phase > 0) stage_sources<4>(P, s_src, phase, a0, n, f0, (int)threadIdx.x - 128, T5_THREADS - 128);
            __syncthreads();
            const uint32_t tile0 = (uint32_t)phase * (uint32_t)ntiles;
            const int pitch = phase == 0 ? (int)T5_SRC_PITCH_A : (int)T5_SRC_PITCH_B;
            for (int ti = gg; ti < ntiles; ti += T5_GEN_GROUPS) {
                const uint32_t tau = tile0 + ti, tl = s_tiles[ti];
                const int ds = tau % T5_DSLOTS;
                const int buf = ti % T5_BBUFS;
                const int src = src1, deg = deg1, row = row1;
                const float4 geo = geo1;
                fetch(ti + T5_GEN_GROUPS, src1, geo1, deg1, row1);
                if (gw == 0) T5_STAMP(ti, 0);
                const bool valid = src != n;
                const float s = geo.x * P.inv_cutoff;
                float sp;
                if (P.env_p == 5) {
                    const float s2 = s * s;
                    sp = s2 * s2 * s;
                } else {
                    sp = s;
                    for (int q = 1; q < P.env_p; ++q) sp *= s;
                }
                float env = 1.0f + P.env_a * sp;
                sp *= s; env += P.env_b * sp;
                sp *= s; env += P.env_c * sp;
                env = (valid && s < 1.0f) ? env * T5_RBF_SCALE : 0.0f;
                int klo = (int)floorf(s * rmax) - 7;
                klo = max(0, min(klo, R - 16));
                // union window of the tile
                int kmin = valid ? klo : 0x7fffffff, kmax = valid ? klo : -1;
                kmin = __reduce_min_sync(ADK_FULL_MASK, kmin);
                kmax = __reduce_max_sync(ADK_FULL_MASK, kmax);
                int* win = s_win + gg * 16 + ((ti / T5_GEN_GROUPS) & 1) * 8;
                if (lane == 0) { win[gw * 2] = kmin; win[gw * 2 + 1] = kmax; }
                named_bar_sync(1 + gg, 128);
                kmin = min(min(win[0], win[2]), min(win[4], win[6]));
                kmax = max(max(win[1], win[3]), max(win[5], win[7]));
                if (kmax < 0) { kmin = 0; kmax = 0; }   // a tile without edges still initialises its accumulator
                const int kbase = kmin & ~15;
                const int nks_total = (kmax + 16 - kbase + 15) >> 4;    // 1..8
                const int nsub = nks_total > 4 ? 2 : 1;
                // The 24 centres of the three aligned 8-blocks that cover this column's 16-tap window [klo, klo + 16).
                // Centres more than 7 spacings away need no mask: 1024 * exp(-d^2 / 2) is below half the smallest fp16
                // subnormal from d = 7 on, so both planes hold exact zeros there.
                const int k8 = min(klo & ~7, R - 24);
                const float d0 = fmaf(-(float)k8, h0, s);   // s - mu_k8, one rounding
                uint32_t hi[12], lo[12];
#pragma unroll
                for (int i = 0; i < (T5_EXP >= 5 ? 0 : 12); ++i) {
                    const float t0 = fmaf(-(float)(2 * i), h0, d0) * P.coeff_sqrt;
                    const float t1 = fmaf(-(float)(2 * i + 1), h0, d0) * P.coeff_sqrt;
                    const float g0 = env * ex2_approx(-t0 * t0), g1 = env * ex2_approx(-t1 * t1);
                    const __half2 hh = __floats2half2_rn(g0, g1);
                    const float2 back = __half22float2(hh);
                    const __half2 ll = __floats2half2_rn(g0 - back.x, g1 - back.y);
                    hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
                    lo[i] = *reinterpret_cast<const uint32_t*>(&ll);
                }
                // per-column metadata for the epilogue, in the ring slot of this tile's accumulator
                if (gw == 0) T5_STAMP(ti, 1);
                mbar_wait(d_empty(ds), ((tau / T5_DSLOTS) & 1) ^ 1);
                if (gw == 0) T5_STAMP(ti, 2);
                {
                    uint8_t* mg = gbase + (meta_smem - base) + ds * T5_META_BYTES;
                    reinterpret_cast<int*>(mg + T5_META_SRC)[c] = src * pitch;
                    float* rh = reinterpret_cast<float*>(mg + T5_META_RHAT) + (c >> 3) * 24 + (c & 7);   // [x0..7 | y0..7 | z0..7]
                    rh[0] = geo.y; rh[8] = geo.z; rh[16] = geo.w;
                    if (j == 0) {
                        reinterpret_cast<int*>(mg + T5_META_CNT)[slot] = max(0, min(T5_SLOT, deg - tile_chunk(tl) * T5_SLOT));
                        reinterpret_cast<int*>(mg + T5_META_ROW)[slot] = row;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(m_full(ds));
                }
                for (int sub = 0; sub < nsub; ++sub) {
                    const int kb = kbase + 64 * sub;
                    const int nks = min(4, nks_total - 4 * sub);
                    mbar_wait(b_empty(buf), ((item_par >> buf) & 1u) ^ 1u);
                    if (gw == 0 && sub == 0) T5_STAMP(ti, 3);
                    const uint32_t bb = b_smem + buf * T5_B_BYTES + row_off;
                    const int q0 = (kb - k8) >> 3;   // my 8-block qq sits in chunk qq - q0 of this operand row
                    // my three 8-blocks (static register indices) ...
                    uint32_t valued = 0u;
#pragma unroll
                    for (int qq = 0; qq < 3; ++qq) {
                        const int ch = qq - q0;
                        if (T5_EXP < 5 && ch >= 0 && ch < 2 * nks) {
                            valued |= 1u << ch;
                            const uint32_t addr = bb + (uint32_t)((ch ^ (c & 7)) * 16);
                            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hi[4 * qq]), "r"(hi[4 * qq + 1]), "r"(hi[4 * qq + 2]), "r"(hi[4 * qq + 3]) : "memory");
                            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr + T5_B_PLANE), "r"(lo[4 * qq]), "r"(lo[4 * qq + 1]), "r"(lo[4 * qq + 2]), "r"(lo[4 * qq + 3]) : "memory");
                        }
                    }
                    // ... and zeros only where this row of the buffer still holds values of an earlier tile (the 16
                    // unconditional zero stores per tile were a tenth of the kernel's shared-memory wavefronts)
                    const uint32_t dirty_b = (dirty >> (8 * buf)) & 0xffu;
                    const uint32_t stale = dirty_b & ~valued & ((1u << (2 * nks)) - 1u);
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        if ((stale >> ch) & 1u) {
                            const uint32_t addr = bb + (uint32_t)((ch ^ (c & 7)) * 16);
                            asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
                            asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr + T5_B_PLANE), "r"(0u) : "memory");
                        }
                    }
                    dirty = (dirty & ~(0xffu << (8 * buf))) | (((dirty_b & ~stale) | valued) << (8 * buf));
                    if (c == 0) {
                        s_item[buf * 4] = kb;
                        s_item[buf * 4 + 1] = nks;
                        s_item[buf * 4 + 2] = (sub == nsub - 1) ? 1 : 0;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> MMA reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(b_full(buf));
                    if (gw == 0) T5_STAMP(ti, 4);
                    item_par ^= 1u << buf;
                }
            }
            __syncthreads();
        }
    } else {
        // ===================== epilogue warpgroups =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 80;");
        const int wg = (warp - T5_EPI_WARP0) >> 2;
        const int team = wg / T5_TEAM_WG, mem = wg % T5_TEAM_WG;   // team = which tiles, member = which rows of them
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int fl = (q & 1) * 32 + lane;           // feature inside the slice
        const int f = f0 + fl;
        const bool upper = q >= 2;                    // lanes 64-127
        const float2 scale2 = make_float2(P.acc_scale, P.acc_scale);
        const float inv_sqrt_h = 1.0f / sqrtf((float)F);
        const float inv_sqrt_3 = 0.57735026918962576451f;
        for (int phase = 0; phase < nphases; ++phase) {
#if T5_WTMEM
            {
                // This phase's weight rows into TMEM (all MMAs of the previous phase have retired: its last
                // __syncthreads came after every epilogue warp's last d_full wait).  Thread = TMEM lane = weight row:
                // phase A [m1 slice | m3 slice], phase B [m2 slice | m2 slice]; a row is 128 fp16 = 64 columns per
                // plane, and each of the four epilogue warpgroups writes half a plane.
                static_assert(T5_EPI_WG == 4, "one (plane, half) per epilogue warpgroup");
                uint4 wb[8];
                load_weight_rows(P, wb, phase, warp, lane, f0);
                store_weight_rows(tmem_base, wb, warp);
            }
#endif
            if (phase > 0) stage_sources<4>(P, s_src, phase, a0, n, f0, (int)threadIdx.x - 128, T5_THREADS - 128);
            __syncthreads();
            const uint32_t tile0 = (uint32_t)phase * (uint32_t)ntiles;
            // phase A: this warpgroup's rows are slots 4 mem .. 4 mem + 3 (lanes 0-63: dx, lanes 64-127: m3 r_hat);
            // phase B: two rows, slots 2 mem, 2 mem + 1 for lanes 0-63 and 4 + 2 mem, 5 + 2 mem for lanes 64-127
            const int s0 = phase == 0 ? 4 * mem : (upper ? 4 + 2 * mem : 2 * mem);
            float acc[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) acc[i] = 0.f;
            const int g = phase == 0 ? (upper ? 2 : 0) : 1;
            const float bias = P.b_rbf[g * F + f];
            const float2 bias2 = make_float2(bias, bias);
            const uint8_t* src_lane = reinterpret_cast<const uint8_t*>(s_src) + (size_t)fl * 4 + ((phase == 0 && upper) ? T5_SF * 4 : 0);
            for (int ti = 0; ti < ntiles; ++ti) {
                const uint32_t tl = s_tiles[ti];
                if (tile_wg(tl) != team) continue;
                const uint32_t tau = tile0 + ti;
                const int ds = tau % T5_DSLOTS;
                const uint32_t par = (tau / T5_DSLOTS) & 1;
                const uint8_t* meta_g = gbase + (meta_smem - base) + ds * T5_META_BYTES;
                if (warp == T5_EPI_WARP0 + 2) T5_STAMP(ti, 8);
                mbar_wait(m_full(ds), par);
                mbar_wait(d_full(ds), par);
                if (warp == T5_EPI_WARP0 + 2) T5_STAMP(ti, 9);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ds * T5_TILE;
                if (T5_EXP == 3 || T5_EXP >= 5) {
                } else if (phase == 0) {
                    if (!upper) tile_body<0, 4>(t_addr, meta_g, src_lane, scale2, bias2, s0, acc);
                    else tile_body<1, 4>(t_addr, meta_g, src_lane, scale2, bias2, s0, acc);
                } else {
                    tile_body<2, 2>(t_addr, meta_g, src_lane, scale2, bias2, s0, acc);
                }
                if (warp == T5_EPI_WARP0 + 2) T5_STAMP(ti, 10);
                // The accumulator and the tile's metadata have been consumed (row sums are in registers): hand the slot back
                // BEFORE the write-out of a finished row group, whose global round trip (~2 300 cycles) used to keep
                // the slot -- and with it the generator and the MMA of this team's next tile -- waiting.
                const bool group_done = tile_last(tl);
                int rr[4] = {-1, -1, -1, -1};
                if (group_done) {
                    const int* rows = reinterpret_cast<const int*>(meta_g + T5_META_ROW);
                    const int nr = phase == 0 ? 4 : 2;
#pragma unroll
                    for (int i = 0; i < 4; ++i) rr[i] = i < nr ? rows[s0 + (i < nr ? i : 0)] : -1;
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(d_empty(ds));
                if (group_done) {
                    // the row group is complete: residuals, write-out, reset
                    if (phase == 0 && !upper) {
                        float xin[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) xin[i] = rr[i] >= 0 ? P.x_io[(size_t)rr[i] * F + f] : 0.f;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (rr[i] >= 0) P.x_io[(size_t)rr[i] * F + f] = (xin[i] + acc[i]) * 0.70710678118654752440f;
                    } else {
                        const float w = phase == 0 ? inv_sqrt_h : inv_sqrt_h * inv_sqrt_3;
                        const bool final_pass = phase == nphases - 1;
                        const float* base_ptr = phase == 0 ? P.vec_in : P.vec_out;   // (phase 0 without vec_in: zeros)
                        // all residual loads first, then the stores (interleaved, every store made the next load wait)
                        float basev[12];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc)
                                basev[3 * i + cc] = (rr[i] >= 0 && base_ptr) ? base_ptr[((size_t)rr[i] * 3 + cc) * F + f] : 0.f;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (rr[i] >= 0) {
#pragma unroll
                                for (int cc = 0; cc < 3; ++cc) {
                                    const size_t off = ((size_t)rr[i] * 3 + cc) * F + f;
                                    const float vo = basev[3 * i + cc] + acc[3 * i + cc] * w;
                                    P.vec_out[off] = vo;
                                    if (final_pass && P.vsplit) {   // operand planes of the vec_proj GEMM that follows
                                        __half h, l;
                                        bool overflow = false;
                                        adk::split_f16x2(vo, P.vsplit_scale, h, l, overflow);
                                        P.vsplit[off] = h;
                                        P.vsplit[P.vsplit_plane + off] = l;
                                        if (overflow && P.status) atomicOr(P.status, ADK_STATUS_F16_OVERFLOW);
                                    }
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 12; ++i) acc[i] = 0.f;
                }
                if (warp == T5_EPI_WARP0 + 2) T5_STAMP(ti, 11);
            }
            __syncthreads();
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    T5_CTA_STAMP(8);
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace

extern "C" int64_t adk_message_t5_smem_bytes(int R, int n_max) {
    if (R != 128 || n_max <= 0 || n_max > T5_MAX_ATOMS) return ADK_ERANGE;
    const int nb = t5_pick_bufs(n_max);
    return nb == 0 ? (int64_t)ADK_ERANGE : (int64_t)t5_smem_bytes(n_max, nb);
}

extern "C" int adk_message_t5(const int32_t* atom_off, int B, int n_max, const int32_t* row_sel, const int32_t* row_start,
                              const int32_t* row_deg,
                              const int32_t* e_src, const float* e_geo, const float* xh, const float* vec_in,
                              const void* w_rbf_split, float w_scale, const float* b_rbf, const float* rbf_offset,
                              int F, int R, float cutoff, int envelope_exponent, float comp, float* x_io,
                              float* vec_out, void* vec_split, int64_t split_rows, float split_scale,
                              uint32_t* status, void* stream) {
    if (!atom_off || !row_start || !row_deg || !e_src || !e_geo || !xh || !w_rbf_split || !b_rbf || !rbf_offset ||
        !x_io || !vec_out || B <= 0 || n_max <= 0)
        return ADK_EINVAL;
    if (F % T5_SF != 0 || R != 128 || envelope_exponent < 1 || vec_in == vec_out || B > 65535) return ADK_EINVAL;
    if (n_max > T5_MAX_ATOMS) return ADK_ERANGE;
    const int nbufs = t5_pick_bufs(n_max);
    if (nbufs == 0) return ADK_ERANGE;
    const size_t smem = t5_smem_bytes(n_max, nbufs);
    alignas(64) CUtensorMap tmW;
#if T5_WTMEM
    memset(&tmW, 0, sizeof(tmW));
#else
    int rc = make_map_f16(&tmW, w_rbf_split, 2 * 3 * (uint64_t)F, (uint64_t)R, 64, 64);
    if (rc != 0) return rc;
#endif
    T5Params P;
    P.w_split = reinterpret_cast<const __half*>(w_rbf_split);
    P.atom_off = atom_off; P.row_sel = row_sel; P.row_start = row_start; P.row_deg = row_deg; P.e_src = e_src;
    P.e_geo = reinterpret_cast<const float4*>(e_geo);
    P.xh = xh; P.vec_in = vec_in; P.b_rbf = b_rbf; P.rbf_offset = rbf_offset;
    P.F = F; P.R = R; P.n_max = n_max;
    P.inv_cutoff = (float)(1.0 / (double)cutoff);
    const double spacing = 1.0 / (double)(R - 1);
    const double p = (double)envelope_exponent;
    P.env_p = envelope_exponent;
    P.env_a = (float)(-(p + 1) * (p + 2) / 2);
    P.env_b = (float)(p * (p + 2));
    P.env_c = (float)(-p * (p + 1) / 2);
    P.coeff_sqrt = (float)(sqrt(0.5 * 1.4426950408889634) / spacing);   // exp(coeff d^2) = 2^-(coeff_sqrt d)^2
    P.acc_scale = (1.0f + comp) / (T5_RBF_SCALE * w_scale);
    P.x_io = x_io; P.vec_out = vec_out;
    P.vsplit = reinterpret_cast<__half*>(vec_split); P.vsplit_plane = split_rows * (int64_t)F;
    P.vsplit_scale = split_scale; P.status = status;
    // A handful of systems: deal the row super-groups (16 rows) of a system to up to eight CTAs per slice, as long as
    // every CTA is resident at once.  The row sums do not depend on the split (one thread, same order).
    int z = 1;
    while (z < 8 && (long long)B * (F / T5_SF) * z * 2 <= g_num_sms) z *= 2;
    // slices of one system are adjacent in launch order: they run together and share its CSR records in L2
    if (nbufs == T5_BBUFS_MAX)
        message_t5_kernel<T5_BBUFS_MAX><<<dim3(F / T5_SF, B, z), T5_THREADS, smem, adk::as_stream(stream)>>>(tmW, P);
    else
        message_t5_kernel<2><<<dim3(F / T5_SF, B, z), T5_THREADS, smem, adk::as_stream(stream)>>>(tmW, P);
    ADK_LAUNCH_CHECK();
    return 0;
}

#ifdef T5_TRACE
extern "C" int adk_message_t5_trace(long long* host_out, int n) {
    return (int)cudaMemcpyFromSymbol(host_out, g_t5_trace, sizeof(long long) * n);
}
#endif

int adk_message_t5_set_attrs() {
    int rc = (int)cudaFuncSetAttribute(message_t5_kernel<T5_BBUFS_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (rc == 0 && T5_BBUFS_MAX != 2)
        rc = (int)cudaFuncSetAttribute(message_t5_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    return rc;
}
