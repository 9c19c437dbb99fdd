// Shared helpers for the lfb200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/lfb200.h"

namespace lf {

void set_error(const char* fmt, ...);

#define LF_CHECK_ARG(cond, ...)                      \
    do {                                             \
        if (!(cond)) {                               \
            lf::set_error(__VA_ARGS__);              \
            return LF_EINVAL;                        \
        }                                            \
    } while (0)

#define LF_RETURN_LAUNCH()                                            \
    do {                                                              \
        cudaError_t e__ = cudaGetLastError();                         \
        if (e__ != cudaSuccess) {                                     \
            lf::set_error("CUDA error: %s", cudaGetErrorString(e__)); \
            return (int)e__;                                          \
        }                                                             \
        return LF_OK;                                                 \
    } while (0)

int sm_count();
// Tuning / A-B switches: read from the environment ONCE (first use), overridable through lf_set_option (tests, tools).
enum Option { OPT_TC_DC, OPT_TC_DEBUG, OPT_TC_NO_DUAL, OPT_RESAMPLE_KC, OPT_RESAMPLE_W, OPT_RESAMPLE_BRICK, OPT_BWDCAM, OPT_TC_NO_TRI, OPT_COUNT };
int option(Option o);

// torch.linspace(a, b, n)[i] exactly as ATen evaluates it (symmetric two-sided formula).
__device__ __forceinline__ float linspace_at(float a, float b, int n, int i) {
    float step = (b - a) / (float)(n - 1);
    return (i < n / 2) ? (a + step * (float)i) : (b - step * (float)(n - 1 - i));
}

// packed 2 x fp32 FMA (Blackwell FFMA2): d.xy = a * b.xy + c.xy with a scalar broadcast
__device__ __forceinline__ void ffma2_bcast(float a, float bx, float by, float& cx, float& cy) {
    unsigned long long ra, rb, rc, rd;
    ra = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
    rb = ((unsigned long long)__float_as_uint(by) << 32) | __float_as_uint(bx);
    rc = ((unsigned long long)__float_as_uint(cy) << 32) | __float_as_uint(cx);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    cx = __uint_as_float((unsigned)rd);
    cy = __uint_as_float((unsigned)(rd >> 32));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// streaming (write-once) 128-bit store: keep the producer's output from evicting the L2-resident cube
__device__ __forceinline__ void st4_stream(float* p, float4 v) {
    __stcs(reinterpret_cast<float4*>(p), v);
}

}  // namespace lf
