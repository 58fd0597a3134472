/*
 * adsorbdiff_b200 -- C ABI of the B200-native PaiNN denoising hot path.
 *
 * Drop-in boundary for the reference's (pure-Python) PaiNN score-model forward and the
 * per-step SE(3) update of its reverse-diffusion sampler.  The reference has no FFI of its
 * own (no native code at all), so each entry point below names the
 * reference Python function(s) whose arithmetic it replaces (file:line under
 * /root/reference/adsorbdiff/).  INTEGRATION.md shows the ctypes binding a maintainer of
 * the reference would add.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer on the current CUDA device unless marked "host";
 *   - the caller owns every buffer; nothing is allocated, freed or retained by the library;
 *   - launches are asynchronous on `stream` (a cudaStream_t passed as void*), no internal
 *     synchronisation, safe to capture into a CUDA graph after adk_init();
 *   - return value: 0 on success, a negative ADK_E* code on argument errors, or the positive
 *     cudaError_t of a failed launch.  Data-dependent failures (a system without neighbours,
 *     an in-degree above the staging capacity) are reported through the device-side `status`
 *     word (bit mask ADK_STATUS_*), read by the host whenever it chooses to synchronise.
 *   - node features: x[N][F], vec[N][3][F] (node, xyz, feature), fp32, F = hidden_channels;
 *     weights in torch.nn.Linear layout [out][in], fp32.
 */
#ifndef ADSORBDIFF_B200_H
#define ADSORBDIFF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADK_ABI_VERSION 3

#define ADK_EINVAL (-22)   /* bad argument (null pointer, unsupported size) */
#define ADK_ERANGE (-34)   /* size beyond a compiled capacity (images, atoms per system) */

#define ADK_STATUS_EMPTY_SYSTEM 1u  /* a system has zero neighbours: reference raises ValueError
                                       (models/painn/painn_denoising.py:370-375) */
#define ADK_STATUS_ROW_OVERFLOW 2u  /* an atom's in-degree exceeds ADK_MAX_ROW_DEGREE */

#define ADK_STATUS_BAD_ELEMENT 8u   /* an atomic number outside [1, num_elements] (nn.Embedding would raise) */
#define ADK_STATUS_F16_OVERFLOW 4u  /* a value left the fp16 range while being split for the tensor-core GEMM */

#define ADK_MAX_IMAGES 2048        /* periodic images enumerated per system */
#define ADK_MAX_ROW_DEGREE 512     /* in-edges per atom after symmetrisation */
#define ADK_MAX_ATOMS_PER_SYSTEM 1024

/* activation codes for adk_linear */
#define ADK_ACT_NONE 0
#define ADK_ACT_SSILU 1            /* ScaledSiLU: silu(x)/0.6 (models/gemnet_oc/layers/base_layers.py:65-72) */

int adk_abi_version(void);

/* One-time per process and device: opts kernels into their dynamic shared-memory sizes.
 * Must be called before any launch and outside stream capture. */
int adk_init(void);

/* Bytes of dynamic shared memory adk_neighbors needs for (n_max atoms, num_images, max_nbrs);
 * negative ADK_ERANGE if it does not fit one CTA. */
int64_t adk_neighbors_smem_bytes(int n_max, int num_images, int max_nbrs);

/*
 * Periodic-boundary neighbour search + per-atom top-k + PBC distances + edge symmetrisation.
 * Replaces: radius_graph_pbc (utils/utils.py:556-730), get_max_neighbors_mask (utils.py:733-853,
 * enforce_max_strictly=True, ties broken by enumeration order), get_pbc_distances
 * (utils.py:513-553), PaiNN.generate_graph_values / symmetrize_edges
 * (models/painn/painn_denoising.py:353-400, 262-327).
 *
 *   pos[N][3], cell[B][3][3] (rows = lattice vectors), atom_off[B+1] (prefix sum of natoms),
 *   rep[3] = image repeats per lattice direction (host ints; 0 for a non-periodic direction),
 *   cutoff2 = fp32(radius*radius), max_nbrs = k.
 * Outputs (in-edge CSR by target atom; system b owns edge slots [2k*atom_off[b], 2k*atom_off[b+1])):
 *   row_start[N], row_deg[N]  absolute slot of each atom's first in-edge, and its in-degree
 *   e_src[2kN], e_tgt[2kN]    global index of the source / target atom of each in-edge
 *   e_geo[2kN][4]             (d, rx, ry, rz): clamped distance and unit vector target->source image
 *   kept_pack[N][k], kept_cnt[N]  the directed half (j<i rule) in reference order, packed
 *                             (j_local<<16 | image_index), for adk_export_edges
 *   sys_counts[B][2]          (raw top-k edge count, kept half count) per system
 *   status                    ADK_STATUS_* bits are OR-ed in
 * Rows are ordered by (d, source, image): deterministic, no atomics on floating point.
 */
int adk_neighbors(const float* pos, const float* cell, const int32_t* atom_off, int B, int n_max,
                  const int32_t rep[3] /* host */, float cutoff2, int max_nbrs,
                  int32_t* row_start, int32_t* row_deg, int32_t* e_src, int32_t* e_tgt, float* e_geo,
                  uint32_t* kept_pack, int32_t* kept_cnt, int32_t* sys_counts,
                  uint32_t* status, void* stream);

/*
 * Materialise the edge list exactly as PaiNN.generate_graph_values returns it
 * (painn_denoising.py:394-400): per system [kept..., reversed...], kept ordered by (i, j, image).
 *   edge_index[2][E_cap] int64 (row 0 = source, row 1 = target), cell_offsets[E_cap][3],
 *   dist[E_cap], unit_vec[E_cap][3], neighbors[B] int64 (= 2 * kept), sys_edge_off[B+1] scratch.
 * The number of valid edges is sys_edge_off[B] (device).  Needed for API parity and tests only;
 * the message kernel consumes the CSR of adk_neighbors directly.
 */
int adk_export_edges(const float* pos, const float* cell, const int32_t* atom_off, int B,
                     const int32_t rep[3] /* host */, int max_nbrs,
                     const uint32_t* kept_pack, const int32_t* kept_cnt, const int32_t* sys_counts,
                     int32_t* sys_edge_off, int64_t* edge_index, int64_t e_cap, float* cell_offsets,
                     float* dist, float* unit_vec, int64_t* neighbors, void* stream);

/* x[n] = emb[z[n]-1] ; vec = 0.  Replaces AtomEmbedding.forward
 * (models/gemnet_oc/layers/embedding_block.py:35-43) and painn_denoising.py:425-426.
 * An atomic number outside [1, num_elements] (nn.Embedding raises IndexError on it) sets
 * ADK_STATUS_BAD_ELEMENT in *status (may be NULL). */
int adk_embed(const int64_t* z, const float* emb, int num_elements, int N, int F,
              float* x, float* vec, uint32_t* status, void* stream);

/* y = LayerNorm(x) * gamma + beta over the last dim (eps 1e-5).  Replaces
 * PaiNNMessage.x_layernorm (painn_denoising.py:517,531). */
int adk_layernorm(const float* x, const float* gamma, const float* beta, int M, int F, float eps,
                  float* y /* may be NULL */, void* y_split /* fp16 [2][split_rows][F] or NULL */,
                  int64_t split_rows, float split_scale, uint32_t* status, void* stream);

/* C[M][N] = act(A[M][K] . W[N][K]^T + bias[N]); lda/ldc in elements; bias may be NULL.
 * The dense contractions of PaiNNMessage.x_proj (painn_denoising.py:508-512,531),
 * PaiNNUpdate.vec_proj/xvec_proj (:580-587,602-613) and GatedEquivariantBlock (:667-676,689-693). */
int adk_linear(const float* A, int64_t lda, const float* W, const float* bias, int M, int N, int K,
               int act, float* C, int64_t ldc, void* stream);

/*
 * Tensor-core variant of adk_linear (tcgen05.mma kind::f16, TMEM accumulators, TMA operand loads),
 * fp32-parity through the "fp16x2 split": an fp32 operand x is held as two fp16 planes
 * hi = fp16(s*x), lo = fp16(s*x - hi) (s a power of two), and A.W^T is accumulated in fp32 as
 * Ah.Wh + Ah.Wl + Al.Wh.  Error vs fp64 ~1e-7 of the output scale (tests/test_gpu_linear_tc.py).
 *
 * adk_split_f16: src fp32 [M][K] (row stride ld) -> dst fp16 [2][plane_rows][K]; rows >= M untouched.
 *   ADK_STATUS_F16_OVERFLOW is OR-ed into *status if |s*x| > 65504.
 * adk_linear_tc: C = act(acc_scale * A.W^T + bias), A = a_split [2][a_plane_rows][K] (a_plane_rows a
 *   multiple of 128, >= M), W = w_split [2][N][K]; N % 16 == 0 (tiles are 256 wide; a ragged last tile is masked), K % 64 == 0; acc_scale = 1/(s_A*s_W).
 *   Outputs: out_f32 [M][ldc] and/or out_split [2][out_plane_rows][N] (scaled by out_split_scale),
 *   either may be NULL (not both).
 */
int adk_split_f16(const float* src, int64_t ld, int M, int K, float scale, void* dst, int64_t plane_rows,
                  uint32_t* status, void* stream);
/*
 * Training-side operand preparation (csrc/train_ops.cu): prescales found on the device, per call, no host round trip.
 *   adk_amax_scale:     rec[0] = s, rec[1] = 1/s with s the power of two that puts max|src| into [target/2, target]
 *                       (s = 1 for an all-zero tensor); scratch = 2 zero-initialised uint32 (left zeroed again)
 *   adk_split_f16_dev:  adk_split_f16 with the scale read from rec[0]; rows [M, plane_rows) are written as zeros
 *   adk_split_f16_t_dev: planes of the TRANSPOSE of src[M][C] (row stride ld): dst [2][plane_rows][Kp], dst[c][m] =
 *                       split(src[m][c]), zeros for c >= C or m >= M (Kp >= M: the reduction length of the GEMM, a
 *                       multiple of 64; plane_rows >= C)
 *   adk_linear_tc_dev:  out_f32[M][ldc] = A . W^T * sa_rec[1] * sb_rec[1] + bias (bias may be NULL); operand layout
 *                       as for adk_linear_tc.  With these: Y = X W^T, dX = dY W, dW = dY^T X of torch.nn.Linear.
 */
int adk_amax_scale(const float* src, int64_t n, float target, float* rec, uint32_t* scratch, void* stream);
int adk_split_f16_dev(const float* src, int64_t ld, int M, int K, const float* rec, void* dst, int64_t plane_rows,
                      uint32_t* status, void* stream);
int adk_split_f16_t_dev(const float* src, int64_t ld, int M, int C, const float* rec, void* dst, int64_t plane_rows,
                        int64_t Kp, uint32_t* status, void* stream);
int adk_linear_tc_dev(const void* a_split, int64_t a_plane_rows, int M, const void* w_split, int N, int K,
                      const float* bias, const float* sa_rec, const float* sb_rec, float* out_f32, int64_t ldc,
                      uint32_t* status, void* stream);
/*
 * One call per torch.nn.Linear pass of the training step (the launches of the four entry points above, issued from C):
 *   adk_linear_train_fwd: y[M][N] = x[M][K] . w[N][K]^T + bias; recs[4] <- {s_x, 1/s_x, s_w, 1/s_w} for the backward
 *   adk_linear_train_bwd: dx[M][K] = g . w (if dx != NULL), dw[N][K] = g^T . x (if dw != NULL); rec_g[2] is scratch
 * ws: adk_linear_train_ws_bytes(M, K, N) bytes of device scratch (operand planes), scratch: the 2 zeroed uint32 of
 * adk_amax_scale.  K % 64 == 0, N % 64 == 0.
 */
int64_t adk_linear_train_ws_bytes(int M, int K, int N);
int adk_linear_train_fwd(const float* x, const float* w, const float* bias, int M, int K, int N, float target,
                         float* recs, void* ws, uint32_t* scratch, uint32_t* status, float* y, void* stream);
int adk_linear_train_bwd(const float* g, const float* x, const float* w, int M, int K, int N, float target,
                         const float* recs, float* rec_g, void* ws, uint32_t* scratch, uint32_t* status,
                         float* dx, float* dw, void* stream);
/*
 * Backward of adk_update_prep / adk_update_gate for the training step (reference: torch autograd through
 * PaiNNUpdate.forward, painn_denoising.py:601-623).
 *   adk_update_prep_bwd: from g_dot[N][F], g_cat[N][2F] -> g_x[N][F] (= g_cat[:, :F]) and g_vp[N][3][2F]
 *   adk_update_gate_bwd: from g_x_out[N][F], g_vec_out[N][3][F] -> g_x[N][F], g_h[N][3F], g_dot[N][F] and
 *                        g_vp[N][3][2F] (v1 half = c * g_vec_out, v2 half = 0); g_vec is g_vec_out itself
 */
int adk_update_prep_bwd(const float* vp, const float* g_dot, const float* g_cat, int N, int F, float* g_x, float* g_vp,
                        void* stream);
int adk_update_gate_bwd(const float* h, const float* dot, const float* vp, const float* scale, const float* g_x_out,
                        const float* g_vec_out, int N, int F, float* g_x, float* g_h, float* g_dot, float* g_vp,
                        void* stream);
/* Tuning knob: run wide GEMMs as cta_group::2 CTA pairs (one MMA over M = 256 rows, each CTA staging half of the
 * weight tile).  On by default (5 % faster at two k-blocks per promotion, bit-identical results; see csrc/linear_tc.cu);
 * ADK_TC_PAIR=0 in the environment also turns it off. */
int adk_set_tc_pair(int enable);

/* Same split for `count` contiguous tensors in one launch: table[i] = {const float* src; fp16* dst;
 * int64 n_elems; float scale; int32 pad} (device array of 32-byte records; dst planes are n_elems apart,
 * n_elems % 4 == 0; scale = that tensor's prescale, 0 = the `scale` argument).
 * Used to re-split every weight at the start of each forward, so in-place parameter updates
 * (EMA swaps through `.data`, optimizer steps) can never leave a stale packed copy behind. */
int adk_split_f16_multi(const void* table, int count, float scale, uint32_t* status, void* stream);
int adk_linear_tc(const void* a_split, int64_t a_plane_rows, int M, const void* w_split, int N, int K,
                  const float* bias, float acc_scale, int act, float* out_f32, int64_t ldc,
                  void* out_split, int64_t out_plane_rows, float out_split_scale, uint32_t* status,
                  void* stream);

/*
 * Fused edge featurisation + rbf projection + message + segmented reduction + residual.
 * Replaces RadialBasis.forward (models/gemnet_oc/layers/radial_basis.py:235-244; Gaussian basis
 * :64-82, polynomial envelope :18-43), PaiNNMessage.rbf_proj/message/aggregate
 * (painn_denoising.py:534-567) and the residual of PaiNN.forward (:443-445).
 *   xh[N][3F] (output of x_proj), vec_in[N][3][F] (NULL => all-zero, layer 0),
 *   w_rbf[3F][R], b_rbf[3F], rbf_offset[R] (Gaussian centres in scaled distance), R = num_rbf
 *   x_io[N][F]: in = x, out = (x + dx)/sqrt(2);  vec_out[N][3][F] = vec_in + dvec
 *   (vec_out must not alias vec_in: other rows still read it).
 * Row-tiled SIMT kernel, no per-system staging: the path for systems too large for adk_message_mma.
 *   row_sel: NULL, or [N]: only rows with row_sel == 1 are computed, every other row of x_io / vec_out is left
 *   untouched (a mixed batch is split between the three message kernels by system size, see PaiNN.plan).
 */
int adk_message(const int32_t* row_sel, const int32_t* row_start, const int32_t* row_deg, const int32_t* e_src,
                const float* e_geo, const float* xh, const float* vec_in, const float* w_rbf,
                const float* b_rbf, const float* rbf_offset, int N, int F, int R, float cutoff,
                int envelope_exponent, float* x_io, float* vec_out, void* stream);

/*
 * Warp-MMA variant of adk_message (csrc/message_mma.cu): one CTA per (system, 32-feature slice); the
 * system's xh / vec slices and the fp16x2 planes of the w_rbf slice are staged in shared memory, and
 * rbfh of 16 distance-sorted edges at a time is a mma.sync m16n8k16 micro-GEMM over the union RBF window.
 *   wt_split = adk_split_f16_transpose(w_rbf[3F][R], scale = w_scale) -> fp16 [2][R][3F]
 *   n_max = the system size the launch is sized for (ADK_ERANGE when its slices do not fit shared memory); systems
 *   with more atoms are skipped (their rows belong to another message kernel, row_sel == 0 there).  Same for adk_message_t5.
 */
int adk_split_f16_transpose(const float* w, int rows, int cols, float scale, void* dst, uint32_t* status,
                            void* stream);
/* Dynamic shared memory adk_message_mma needs for systems of up to n_max atoms (weight planes + the staged
 * source features of one system), or negative ADK_ERANGE if that exceeds the 227 KB of an sm_100a CTA (the
 * caller then uses adk_message for the batch). */
int64_t adk_message_mma_smem_bytes(int R, int n_max);
int adk_message_mma(const int32_t* atom_off, int B, int n_max,
                    const int32_t* row_sel /* NULL, or [N]: 1 = compute the row, 2 = pass it through (vec_out =
                                              vec_in, x untouched), 0 = leave it untouched (the sampler's tail) */,
                    const int32_t* row_start, const int32_t* row_deg, const int32_t* e_src, const float* e_geo, const float* xh,
                    const float* vec_in, const void* wt_split, float w_scale, const float* b_rbf,
                    const float* rbf_offset, int F, int R, float cutoff, int envelope_exponent,
                    float comp, float* x_io, float* vec_out,
                    void* vec_split /* fp16 [2][split_rows][F], row = atom*3+xyz, or NULL */, int64_t split_rows,
                    float split_scale, uint32_t* status, void* stream);

/*
 * tcgen05 variant of adk_message_mma, the default for full batches (csrc/message_t5.cu): one CTA per (system,
 * 64-feature slice); D[feature][edge] = W . rbf^T by tcgen05.mma into TMEM (fp16x2 split, banded K), the weight
 * tile as the TMEM-resident A operand (tcgen05.st once per phase), sources of the whole system staged in shared
 * memory, epilogue threads own one feature (= TMEM lane) and reduce the CSR rows in registers.  Same arithmetic
 * contract as adk_message; requires R == 128, F % 64 == 0 and a system that fits (adk_message_t5_smem_bytes > 0:
 * up to 199 atoms at F = 512; four operand buffers up to 114 atoms, two beyond), else the caller uses
 * adk_message_mma / adk_message.
 *   w_rbf_split = fp16 [2][3F][R] planes of w_rbf scaled by w_scale (adk_split_f16_multi)
 */
int64_t adk_message_t5_smem_bytes(int R, int n_max);
int adk_message_t5(const int32_t* atom_off, int B, int n_max,
                   const int32_t* row_sel /* NULL, or [N] with the meaning it has for adk_message_mma */,
                   const int32_t* row_start, const int32_t* row_deg, const int32_t* e_src, const float* e_geo, const float* xh, const float* vec_in,
                   const void* w_rbf_split, float w_scale, const float* b_rbf, const float* rbf_offset,
                   int F, int R, float cutoff, int envelope_exponent, float comp, float* x_io, float* vec_out,
                   void* vec_split /* fp16 [2][split_rows][F], row = atom*3+xyz, or NULL */, int64_t split_rows,
                   float split_scale, uint32_t* status, void* stream);

/*
 * Backward of the message op for the training step (csrc/message_bwd.cu; the reference differentiates
 * PaiNNMessage.message/aggregate with torch autograd, painn_denoising.py:534-567).  With g_dx[N][F] = dL/d dx and
 * g_dvec[N][3][F] = dL/d dvec (dx, dvec as defined for adk_message, before the residual / comp):
 *   d_xh[N][3F], d_vec[N][3][F] (message part only), d_w[3F][R], d_b[3F].
 * Uses the symmetry of the edge list (every edge has a mirror of equal length and negated unit vector), so the
 * transposed aggregation is a walk over the same in-edge CSR.  Exact fp32, deterministic.  F % 128 == 0.
 * scratch: adk_message_bwd_scratch_floats(N, F, R, NULL) floats.  vec_in may be NULL (first layer: vec == 0).
 * plan: the edges of each row chunk sorted by the 16-centre slot their Gaussian window starts in, with everything
 *   the weight-gradient pass needs per edge; built once per graph by adk_message_bwd_plan (R == 128) into
 *   adk_message_bwd_plan_ints(N, e_cap) int32 (e_cap >= number of edges, e.g. the capacity of e_src) and reused by
 *   every layer's backward.
 */
int64_t adk_message_bwd_scratch_floats(int N, int F, int R, int* chunks_out /* may be NULL */);
int64_t adk_message_bwd_plan_ints(int N, int64_t e_cap);
int adk_message_bwd_plan(const int32_t* row_start, const int32_t* row_deg, const int32_t* e_src, const float* e_geo,
                         int N, int R, float cutoff, int envelope_exponent, int32_t* plan, void* stream);
int adk_message_bwd(const int32_t* row_start, const int32_t* row_deg, const int32_t* e_src, const float* e_geo,
                    const int32_t* plan, const float* xh, const float* vec_in, const float* w_rbf, const float* b_rbf,
                    const float* rbf_offset, int N, int F, int R, float cutoff, int envelope_exponent,
                    const float* g_dx, const float* g_dvec, float* d_xh, float* d_vec, float* d_w, float* d_b,
                    float* scratch, void* stream);

/* From vp[N][3][2F] = vec_proj(vec) = (vec1|vec2): dot[N][F] = sum_xyz vec1*vec2 / sqrt(F),
 * cat[N][2F] = [x | sqrt(sum_xyz vec2^2 + 1e-8)].  PaiNNUpdate.forward (painn_denoising.py:602-613). */
int adk_update_prep(const float* x, const float* vp, int N, int F, float* dot, float* cat /* may be NULL */,
                    void* cat_split /* fp16 [2][split_rows][2F] or NULL */, int64_t split_rows, float split_scale,
                    uint32_t* status, void* stream);

/* h[N][3F] = (a|b|c): x = (x + (a + b*dot)/sqrt(2)) * scale ; vec += c * vec1 (vec1 = vp[:, :, :F]).
 * PaiNNUpdate.forward (:614-623), PaiNN.forward (:449-451), ScaleFactor.forward
 * (modules/scaling/scale_factor.py:157-172; scale == 0 means "not fitted": no multiply). */
int adk_update_gate(const float* h, const float* dot, const float* vp, const float* scale /* device scalar */,
                    int N, int F, float* x, float* vec,
                    void* vec_split /* fp16 [2][split_rows][F] planes of the new vec (row = atom*3+xyz), or NULL */,
                    int64_t split_rows, float split_scale, uint32_t* status, void* stream);

/* Row gather / scatter of fp32 matrices with W columns: dst[i] = src[rows[i]] / dst[rows[i]] = src[i],
 * i < M.  Used by the sampler, which needs the last layer and the output heads only for the adsorbate atoms
 * (Denoiser._get_ads_output, denoising_torch.py:460-467, averages the scores over tags == 2 and reads nothing else). */
int adk_gather_rows(const float* src, const int32_t* rows, int M, int W, float* dst, void* stream);
/* Rows a message layer has to compute so that the NEXT layer can be evaluated at the rows `sel` only: out[i] = 1 for
 * i in sel and for every source of an in-edge of a row in sel, else 2 (the row_sel codes of adk_message_mma). */
int adk_mark_sources(const int32_t* row_start, const int32_t* row_deg, const int32_t* e_src, const int32_t* sel,
                     int n_sel, int N, int32_t* out, void* stream);
int adk_scatter_rows(const float* src, const int32_t* rows, int M, int W, float* dst, void* stream);

/* GatedEquivariantBlock (painn_denoising.py:688-697), the parts around its linears:
 * prep: cat[N][2C] = [x | ||v1p||_xyz] from v1p[N][3][C] = vec1_proj(v);
 * gate: from u[N][2*Co] = (s|g): x_out[N][Co] = ssilu(s), v_out[N][3][Co] = g * v2p (v2p = vec2_proj(v)). */
int adk_head_prep(const float* x, const float* v1p, int N, int C, float* cat /* may be NULL */,
                  void* cat_split /* fp16 [2][split_rows][2C] or NULL */, int64_t split_rows, float split_scale,
                  uint32_t* status, void* stream);
int adk_head_gate(const float* u, const float* v2p, int N, int Co, float* x_out, float* v_out,
                  void* v_split /* fp16 [2][split_rows][Co] planes of v_out, Co % 4 == 0, or NULL */, int64_t split_rows,
                  float split_scale, uint32_t* status, void* stream);

/*
 * Initial placement: random in-plane centre of mass for the adsorbate (tags == 2), z kept.
 * Replaces Denoiser.reverse_sde_sampling_rot lines denoising_torch.py:215-232.
 * noise[B][3] = the torch.rand(B,3) draw.  pos is updated in place.
 */
int adk_init_placement(float* pos, const float* cell, const int32_t* atom_off, const int32_t* tags,
                       const float* noise, int B, void* stream);

/*
 * One reverse-diffusion step (ODE or SDE) on the rigid adsorbate of every system.
 * Replaces DiffTorchCalc.get_denoising_prediction (denoising_torch.py:491-500), _get_ads_output
 * (:460-467), the update / PBC wrap / rigid rotation of reverse_sde_sampling_rot (:266-353) and
 * axis_angle_to_matrix (utils/rot_utils.py:18-98).
 *   score_tr/score_rot[N][3] = the two model outputs.
 *   sched[S][ADK_SCHED_COLS] = per-step scalars, evaluated on the host with the reference's own expressions:
 *     [0] 0.5*tr_g^2*dt  [1] dt  [2] fp32(rot_g^2)  [3] tr_g^2*dt  [4] tr_g*sqrt(dt)  [5] fp32(rot_g*sqrt(dt));
 *   the row used is sched[*step], and *step (a device counter) is incremented afterwards, so a captured
 *   CUDA graph can be replayed for all S steps without host-side parameter updates.
 *   noise == NULL: ODE (`ode=True`, :269-272): delta COM = sched[0] * mean_tr (z zeroed), rotation vector =
 *   ((0.5*mean_rot)*dt)*rot_g2.
 *   noise[S][2][B][3]: SDE (`ode=False`, :273-295): noise[s][0] = tr_z and noise[s][1] = rot_z of step s, the
 *   standard-normal draws the reference makes with torch.normal; delta COM = sched[3]*mean_tr + sched[4]*tr_z
 *   (z zeroed), rotation vector = (mean_rot*dt)*rot_g2 + sched[5]*rot_z.
 *   In both cases the COM is wrapped into the cell and the adsorbate moved rigidly; pos updated in place;
 *   max_abs_upd[B] = max |delta COM| per system (for the reference's allclose early-stop, :312-320).
 */
#define ADK_SCHED_COLS 6
int adk_se3_step(float* pos, const float* cell, const int32_t* atom_off, const int32_t* tags,
                 const int32_t* fixed, const float* score_tr, const float* score_rot,
                 const float* sched, int32_t* step, int B, const float* noise /* NULL = ODE */,
                 float* max_abs_upd, const int32_t* stop /* adk_early_stop state or NULL */, void* stream);

/*
 * Device-side form of the reference's batch-wide early stop (denoising_torch.py:312-320): after a step,
 * allclose(delta COM, 0, rtol=1e-3, atol=1e-3) over all B systems (max_abs_upd from adk_se3_step) bumps
 * stop[0]; at the tenth hit the reference breaks BEFORE applying that step, so stop[1] is raised, stop[2]
 * records the number of applied steps and pos is restored from prev (the caller's copy of pos taken before
 * the step).  Once stop[1] is set adk_se3_step leaves pos and *step untouched, so a host that polls stop[1]
 * only every few steps still ends with exactly the reference's positions.  stop = int32[3], zeroed by the caller.
 */
int adk_early_stop(const float* max_abs_upd, int B, float atol, const int32_t* step, int32_t* stop,
                   float* pos, const float* prev, int64_t n_values, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADSORBDIFF_B200_H */
