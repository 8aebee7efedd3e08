// cmf_common.cuh -- shared helpers for libcmflow_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cmflow_b200.h"

void cmf_set_error(const char *fmt, ...);

#define CMF_REQUIRE(cond, msg)                                                    \
    do {                                                                          \
        if (!(cond)) {                                                            \
            cmf_set_error("%s: invalid argument: %s", __func__, msg);             \
            return CMF_ERR_INVALID;                                               \
        }                                                                         \
    } while (0)

#define CMF_CUDA(call)                                                            \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) {                                                  \
            cmf_set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e_)); \
            return CMF_ERR_CUDA;                                                  \
        }                                                                         \
    } while (0)

#define CMF_LAUNCH_CHECK()                                                        \
    do {                                                                          \
        cudaError_t e_ = cudaGetLastError();                                      \
        if (e_ != cudaSuccess) {                                                  \
            cmf_set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e_)); \
            return CMF_ERR_CUDA;                                                  \
        }                                                                         \
    } while (0)

static inline int cmf_divup(long long a, long long b) { return (int)((a + b - 1) / b); }

// (ax-bx)^2 + (ay-by)^2 + (az-bz)^2 in exactly the contraction nvcc -O2 gives the reference kernels
// (read from the PTX of lib/src/ball_query_gpu.cu:34, interpolate_gpu.cu:40,103, sampling_gpu.cu:135):
//   fma(dz,dz, fma(dx,dx, dy*dy)).  Written with intrinsics so no compiler flag can change it.
__device__ __forceinline__ float cmf_sqdist_ref(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// |p|^2 as torch.sum(p**2, -1) evaluates it: three rounded squares, added left to right, no fma
// (utils/model_utils/radarflow_util.py:27-28).
__device__ __forceinline__ float cmf_sqnorm3(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// square_distance's expanded form (radarflow_util.py:26-29); nq = |q|^2, nx = |x|^2.
__device__ __forceinline__ float cmf_sqdist_expanded(float qx, float qy, float qz, float nq,
                                                     float x, float y, float z, float nx) {
    float dot = __fmaf_rn(qz, z, __fmaf_rn(qy, y, __fmul_rn(qx, x)));
    float d = __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), nq), nx);
    return fmaxf(d, 0.0f);
}

// Coalesced copy of `count` floats global -> shared; 128-bit loads when the source is 16-byte aligned.
__device__ __forceinline__ void cmf_stage_floats(float *__restrict__ dst, const float *__restrict__ src, int count) {
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        int n4 = count >> 2;
        const float4 *s4 = reinterpret_cast<const float4 *>(src);
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        for (int i = threadIdx.x; i < n4; i += blockDim.x) d4[i] = __ldg(s4 + i);
        for (int i = (n4 << 2) + threadIdx.x; i < count; i += blockDim.x) dst[i] = __ldg(src + i);
    } else {
        for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = __ldg(src + i);
    }
}
