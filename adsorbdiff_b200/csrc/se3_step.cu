// K4: the sampler's per-step SE(3) update of the rigid adsorbate, one warp per system.
// Reference (under /root/reference/adsorbdiff/): relaxation/diffusers/denoising_torch.py:215-232
// (initial placement), :266-353 (ODE update, PBC wrap of the centre of mass, rigid rotation),
// :460-467 (_get_ads_output = per-system mean over tags == 2), :491-500 (rotation score zeroed
// on fixed atoms); utils/rot_utils.py:18-98 (axis-angle -> quaternion -> matrix).
#include "common.cuh"

namespace {

struct Mean3 { float x, y, z; };

// per-system mean over adsorbate atoms (scatter(..., reduce="mean"): sum / max(count, 1))
__device__ __forceinline__ Mean3 ads_mean(const float* v, const int32_t* tags, const int32_t* fixed_mask,
                                          int a0, int n, int lane) {
    float sx = 0.f, sy = 0.f, sz = 0.f, cnt = 0.f;
    for (int i = lane; i < n; i += 32) {
        if (tags[a0 + i] == 2) {
            cnt += 1.f;
            if (fixed_mask == nullptr || fixed_mask[a0 + i] != 1) {
                sx += v[3 * (size_t)(a0 + i)];
                sy += v[3 * (size_t)(a0 + i) + 1];
                sz += v[3 * (size_t)(a0 + i) + 2];
            }
        }
    }
    sx = adk::warp_sum(sx); sy = adk::warp_sum(sy); sz = adk::warp_sum(sz); cnt = adk::warp_sum(cnt);
    cnt = fmaxf(cnt, 1.f);
    return Mean3{sx / cnt, sy / cnt, sz / cnt};
}

__device__ __forceinline__ float pymod1(float x) {
    // torch `x % 1`: result takes the sign of the divisor (in [0, 1))
    float r = fmodf(x, 1.0f);
    if (r != 0.f && r < 0.f) r += 1.0f;
    return r;
}

// Solve A f = v for a 3x3 system, Gaussian elimination with partial pivoting (as LAPACK gesv
// behind torch.linalg.solve does).  A row-major.
__device__ __forceinline__ void solve3(const float* A, const float* v, float* f) {
    float M[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r) { M[r][0] = A[3 * r]; M[r][1] = A[3 * r + 1]; M[r][2] = A[3 * r + 2]; M[r][3] = v[r]; }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int piv = c;
        float best = fabsf(M[c][c]);
#pragma unroll
        for (int r = c + 1; r < 3; ++r)
            if (fabsf(M[r][c]) > best) { best = fabsf(M[r][c]); piv = r; }
        if (piv != c) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { float t = M[c][q]; M[c][q] = M[piv][q]; M[piv][q] = t; }
        }
#pragma unroll
        for (int r = c + 1; r < 3; ++r) {
            float l = M[r][c] / M[c][c];
#pragma unroll
            for (int q = c; q < 4; ++q) M[r][q] -= l * M[c][q];
        }
    }
    f[2] = M[2][3] / M[2][2];
    f[1] = (M[1][3] - M[1][2] * f[2]) / M[1][1];
    f[0] = (M[0][3] - M[0][1] * f[1] - M[0][2] * f[2]) / M[0][0];
}

__global__ void init_placement_kernel(float* pos, const float* __restrict__ cell,
                                      const int32_t* __restrict__ atom_off, const int32_t* __restrict__ tags,
                                      const float* __restrict__ noise, int B) {
    const int b = blockIdx.x * (blockDim.x >> 5) + adk::warp_id();
    if (b >= B) return;
    const int lane = adk::lane_id();
    const int a0 = atom_off[b], n = atom_off[b + 1] - a0;
    const float* c = cell + 9 * (size_t)b;
    const float* u = noise + 3 * (size_t)b;
    Mean3 com = ads_mean(pos, tags, nullptr, a0, n, lane);
    // einsum("bi,bij->bj", noise, cell^T): out[j] = sum_i noise[i] * cell[j][i]
    float nx = u[0] * c[0] + u[1] * c[1] + u[2] * c[2];
    float ny = u[0] * c[3] + u[1] * c[4] + u[2] * c[5];
    float nz = com.z;  // keep the height
    for (int i = lane; i < n; i += 32) {
        if (tags[a0 + i] == 2) {
            float* p = pos + 3 * (size_t)(a0 + i);
            p[0] = (p[0] - com.x) + nx;
            p[1] = (p[1] - com.y) + ny;
            p[2] = (p[2] - com.z) + nz;
        }
    }
}

__global__ void se3_step_kernel(float* pos, const float* __restrict__ cell, const int32_t* __restrict__ atom_off,
                                const int32_t* __restrict__ tags, const int32_t* __restrict__ fixed,
                                const float* __restrict__ score_tr, const float* __restrict__ score_rot,
                                const float* __restrict__ sched, const int32_t* __restrict__ step, int B,
                                const float* __restrict__ noise, float* max_abs_upd,
                                const int32_t* __restrict__ stop) {
    const int b = blockIdx.x * (blockDim.x >> 5) + adk::warp_id();
    if (b >= B) return;
    if (stop && stop[1]) return;  // the run has stopped early (adk_early_stop): positions stay as they are
    const int s_idx = *step;
    const float* sc = sched + ADK_SCHED_COLS * (size_t)s_idx;
    const float c_tr = sc[0], dt = sc[1], rot_g2 = sc[2];
    const int lane = adk::lane_id();
    const int a0 = atom_off[b], n = atom_off[b + 1] - a0;
    const float* cl = cell + 9 * (size_t)b;

    Mean3 tr = ads_mean(score_tr, tags, nullptr, a0, n, lane);
    Mean3 rt = ads_mean(score_rot, tags, fixed, a0, n, lane);  // positions_free[fixed == 1] = 0
    Mean3 com = ads_mean(pos, tags, nullptr, a0, n, lane);

    // ODE step (:269-272): d_com = 0.5 g^2 dt * score ; rot_vec = ((0.5 * score) * dt) * g_rot^2
    float upd[3] = {c_tr * tr.x, c_tr * tr.y, 0.f};  // z component zeroed (:297)
    float rv[3] = {((0.5f * rt.x) * dt) * rot_g2, ((0.5f * rt.y) * dt) * rot_g2, ((0.5f * rt.z) * dt) * rot_g2};
    if (noise) {
        // SDE step (:273-295): d_com = (g^2 dt) * score + (g sqrt(dt)) * z_tr ;
        // rot_vec = (score * dt) * g_rot^2 + (g_rot sqrt(dt)) * z_rot, every product and the sum rounded to fp32
        // like the reference's tensor expression.  noise[step][2][B][3] holds this step's two normal draws.
        const float c_sde = sc[3], n_tr = sc[4], n_rot = sc[5];
        const float* zt = noise + ((size_t)s_idx * 2 * B + b) * 3;
        const float* zr = zt + (size_t)B * 3;
        upd[0] = __fadd_rn(__fmul_rn(c_sde, tr.x), __fmul_rn(n_tr, zt[0]));
        upd[1] = __fadd_rn(__fmul_rn(c_sde, tr.y), __fmul_rn(n_tr, zt[1]));
        rv[0] = __fadd_rn(__fmul_rn(__fmul_rn(rt.x, dt), rot_g2), __fmul_rn(n_rot, zr[0]));
        rv[1] = __fadd_rn(__fmul_rn(__fmul_rn(rt.y, dt), rot_g2), __fmul_rn(n_rot, zr[1]));
        rv[2] = __fadd_rn(__fmul_rn(__fmul_rn(rt.z, dt), rot_g2), __fmul_rn(n_rot, zr[2]));
    }

    // wrap the centre of mass into the cell: f = solve(cell, com + upd); f %= 1 (twice); back (:298-310)
    float target[3] = {com.x + upd[0], com.y + upd[1], com.z + upd[2]};
    float fr[3];
    solve3(cl, target, fr);
#pragma unroll
    for (int q = 0; q < 3; ++q) fr[q] = pymod1(pymod1(fr[q]));
    const float comv[3] = {com.x, com.y, com.z};
#pragma unroll
    for (int jx = 0; jx < 3; ++jx)
        upd[jx] = (fr[0] * cl[3 * jx] + fr[1] * cl[3 * jx + 1] + fr[2] * cl[3 * jx + 2]) - comv[jx];

    if (max_abs_upd && lane == 0) max_abs_upd[b] = fmaxf(fabsf(upd[0]), fmaxf(fabsf(upd[1]), fabsf(upd[2])));

    // axis-angle -> quaternion -> rotation matrix (rot_utils.py:50-98, 18-47)
    const float angle = sqrtf(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
    const float half = 0.5f * angle;
    const float kq = (fabsf(angle) < 1e-6f) ? (0.5f - (angle * angle) / 48.0f) : (sinf(half) / angle);
    const float qr = cosf(half), qi = rv[0] * kq, qj = rv[1] * kq, qk = rv[2] * kq;
    const float two_s = 2.0f / (qr * qr + qi * qi + qj * qj + qk * qk);
    const float R[9] = {1 - two_s * (qj * qj + qk * qk), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr),
                        two_s * (qi * qj + qk * qr), 1 - two_s * (qi * qi + qk * qk), two_s * (qj * qk - qi * qr),
                        two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi * qi + qj * qj)};

    // new = (p - com) @ R^T + upd + com   (:331-336)
    for (int i = lane; i < n; i += 32) {
        if (tags[a0 + i] == 2) {
            float* p = pos + 3 * (size_t)(a0 + i);
            const float rx = p[0] - com.x, ry = p[1] - com.y, rz = p[2] - com.z;
#pragma unroll
            for (int x = 0; x < 3; ++x)
                p[x] = ((rx * R[3 * x] + ry * R[3 * x + 1] + rz * R[3 * x + 2]) + upd[x]) + comv[x];
        }
    }
}

__global__ void bump_step_kernel(int32_t* step, const int32_t* stop) {
    if (!stop || !stop[1]) *step += 1;
}

// The reference's batch-wide convergence test (denoising_torch.py:312-320): allclose(delta COM, 0, rtol 1e-3,
// atol 1e-3) over ALL systems bumps a counter; at the tenth hit the loop breaks BEFORE that step's update is
// applied.  stop[0] = hit count, stop[1] = stopped flag, stop[2] = number of steps applied when it stopped.
__global__ void early_stop_decide_kernel(const float* __restrict__ max_abs_upd, int B, float atol,
                                         const int32_t* __restrict__ step, int32_t* stop) {
    __shared__ int s_any;
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    if (stop[1]) return;
    int bad = 0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) bad |= !(max_abs_upd[i] <= atol) ? 1 : 0;
    if (bad) s_any = 1;
    __syncthreads();
    if (threadIdx.x == 0 && !s_any) {
        stop[0] += 1;
        if (stop[0] == 10) {
            stop[1] = 1;
            stop[2] = *step - 1;  // adk_se3_step already bumped the counter for this step
        }
    }
}
__global__ void early_stop_rollback_kernel(float* pos, const float* __restrict__ prev, int64_t n,
                                           const int32_t* __restrict__ stop) {
    if (!stop[1]) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        pos[i] = prev[i];
}

}  // namespace

extern "C" int adk_init_placement(float* pos, const float* cell, const int32_t* atom_off, const int32_t* tags,
                                  const float* noise, int B, void* stream) {
    if (!pos || !cell || !atom_off || !tags || !noise || B <= 0) return ADK_EINVAL;
    init_placement_kernel<<<(B + 3) / 4, 128, 0, adk::as_stream(stream)>>>(pos, cell, atom_off, tags, noise, B);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_se3_step(float* pos, const float* cell, const int32_t* atom_off, const int32_t* tags,
                            const int32_t* fixed, const float* score_tr, const float* score_rot,
                            const float* sched, int32_t* step, int B, const float* noise, float* max_abs_upd,
                            const int32_t* stop, void* stream) {
    if (!pos || !cell || !atom_off || !tags || !fixed || !score_tr || !score_rot || !sched || !step || B <= 0)
        return ADK_EINVAL;
    se3_step_kernel<<<(B + 3) / 4, 128, 0, adk::as_stream(stream)>>>(pos, cell, atom_off, tags, fixed, score_tr,
                                                                   score_rot, sched, step, B, noise, max_abs_upd, stop);
    ADK_LAUNCH_CHECK();
    bump_step_kernel<<<1, 1, 0, adk::as_stream(stream)>>>(step, stop);
    ADK_LAUNCH_CHECK();
    return 0;
}

extern "C" int adk_early_stop(const float* max_abs_upd, int B, float atol, const int32_t* step, int32_t* stop,
                              float* pos, const float* prev, int64_t n_values, void* stream) {
    if (!max_abs_upd || !step || !stop || !pos || !prev || B <= 0 || n_values <= 0) return ADK_EINVAL;
    early_stop_decide_kernel<<<1, 256, 0, adk::as_stream(stream)>>>(max_abs_upd, B, atol, step, stop);
    ADK_LAUNCH_CHECK();
    const int blocks = (int)((n_values + 255) / 256 < 1184 ? (n_values + 255) / 256 : 1184);
    early_stop_rollback_kernel<<<blocks, 256, 0, adk::as_stream(stream)>>>(pos, prev, n_values, stop);
    ADK_LAUNCH_CHECK();
    return 0;
}
