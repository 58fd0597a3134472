// Shared helpers for the adsorbdiff_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "adsorbdiff_b200.h"

#define ADK_FULL_MASK 0xffffffffu

#define ADK_LAUNCH_CHECK()                         \
    do {                                           \
        cudaError_t e__ = cudaGetLastError();      \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

namespace adk {

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(ADK_FULL_MASK, v, o);
    return v;
}

// silu(x)/0.6  (reference: models/gemnet_oc/layers/base_layers.py:65-72)
__device__ __forceinline__ float ssilu(float x) {
    return (x / (1.0f + expf(-x))) * (1.0f / 0.6f);
}

// packed fp32x2 FMA (Blackwell FFMA2): d = a * b + c on both halves, one issue slot
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c);
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

// fp16x2 split of one value for the tensor-core GEMM operands: hi = fp16(s*v), lo = fp16(s*v - hi)
__device__ __forceinline__ void split_f16x2(float v, float scale, __half& hi, __half& lo, bool& overflow) {
    const float sv = v * scale;
    overflow |= !(fabsf(sv) <= 65504.0f);
    hi = __float2half_rn(sv);
    lo = __float2half_rn(sv - __half2float(hi));
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace adk

// per-translation-unit attribute setup, called by adk_init()
int adk_neighbors_set_attrs();
int adk_message_set_attrs();
int adk_linear_set_attrs();
int adk_linear_tc_set_attrs();
int adk_message_mma_set_attrs();
int adk_message_t5_set_attrs();
int adk_message_bwd_set_attrs();
int adk_message_bwd_nodes(const int32_t* row_start, const int32_t* row_deg, const int32_t* e_src, const float* e_geo,
                          const float* xh, const float* vec_in, const float* w_rbf, const float* b_rbf,
                          const float* rbf_offset, int N, int F, int R, float cutoff, int envelope_exponent,
                          const float* g_dx, const float* g_dvec, float* d_xh, float* d_vec, void* stream);
