// Shared helpers for libsilent_b200: error plumbing and the canonical float32 arithmetic (DESIGN.md "Canonical order").
// Compiled with -fmad=false: every fused multiply-add in this library is an explicit fmaf()/fma().
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/silent_b200.h"

namespace silent {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int fail(int status, const char *fmt, ...);

#define SILENT_CUDA(call)                                                                                     \
    do {                                                                                                      \
        cudaError_t err__ = (call);                                                                           \
        if (err__ != cudaSuccess)                                                                             \
            return ::silent::fail(SILENT_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__),   \
                                  __FILE__, __LINE__);                                                        \
    } while (0)

#define SILENT_LAUNCH_CHECK(name)                                                                             \
    do {                                                                                                      \
        cudaError_t err__ = cudaGetLastError();                                                               \
        if (err__ != cudaSuccess)                                                                             \
            return ::silent::fail(SILENT_E_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(err__)); \
        ::silent::count_launch();                                                                             \
    } while (0)

constexpr int kTaps = 6;       // order-5 spline: 6 taps per axis
constexpr int kMaxChannels = 8;
constexpr int kMaxKernel = 7;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch: the kernels of one pipeline step are launched with the stream-serialisation
//      attribute, so a kernel's CTAs are scheduled while its predecessor drains and only its FIRST instruction waits for
//      the predecessor's memory (no launch gap between the step's 6 kernels; matters most at batch 1). A kernel that is
//      launched without the attribute passes both instructions at once.
__device__ __forceinline__ void pdl_enter()
{
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                    Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- canonical float32 element ops -----------------------------------------------------------------------------------

__device__ __forceinline__ float canon_relu(float v) { return v < 0.0f ? 0.0f : v; }   // NaN propagates
__device__ __forceinline__ float canon_clip_hi(float v, float hi) { return v > hi ? hi : v; }

// log2 / exp2 in double precision from +, *, /, fma only: bit-identical to oracle/silent_oracle.c on any IEEE machine.
__device__ inline double canon_log2(double x)
{
    unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    bits = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double f = __longlong_as_double((long long)bits);
    if (f > 1.4142135623730951) {
        f = f * 0.5;
        e += 1;
    }
    double t = (f - 1.0) / (f + 1.0);
    double t2 = t * t;
    double s = 1.0 / 27.0;
    s = fma(s, t2, 1.0 / 25.0);
    s = fma(s, t2, 1.0 / 23.0);
    s = fma(s, t2, 1.0 / 21.0);
    s = fma(s, t2, 1.0 / 19.0);
    s = fma(s, t2, 1.0 / 17.0);
    s = fma(s, t2, 1.0 / 15.0);
    s = fma(s, t2, 1.0 / 13.0);
    s = fma(s, t2, 1.0 / 11.0);
    s = fma(s, t2, 1.0 / 9.0);
    s = fma(s, t2, 1.0 / 7.0);
    s = fma(s, t2, 1.0 / 5.0);
    s = fma(s, t2, 1.0 / 3.0);
    s = fma(s, t2, 1.0);
    double lnf = (2.0 * t) * s;
    return fma(lnf, 1.4426950408889634, (double)e);
}

__device__ inline double canon_exp2(double y)
{
    double n = rint(y);
    double g = y - n;
    double z = g * 0.6931471805599453;
    double s = 1.0 / 87178291200.0;
    s = fma(s, z, 1.0 / 6227020800.0);
    s = fma(s, z, 1.0 / 479001600.0);
    s = fma(s, z, 1.0 / 39916800.0);
    s = fma(s, z, 1.0 / 3628800.0);
    s = fma(s, z, 1.0 / 362880.0);
    s = fma(s, z, 1.0 / 40320.0);
    s = fma(s, z, 1.0 / 5040.0);
    s = fma(s, z, 1.0 / 720.0);
    s = fma(s, z, 1.0 / 120.0);
    s = fma(s, z, 1.0 / 24.0);
    s = fma(s, z, 1.0 / 6.0);
    s = fma(s, z, 0.5);
    s = fma(s, z, 1.0);
    s = fma(s, z, 1.0);
    int ni = (int)n;
    ni = ni < -1000 ? -1000 : (ni > 1000 ? 1000 : ni);
    double scale = __longlong_as_double((long long)(ni + 1023) << 52);
    return s * scale;
}

// pow(x, r) of the regulator (tf.pow, util/regulator/gaussian_regulator_tensor.py:35).
__device__ inline float canon_pow(float x, float r)
{
    if (r == 0.0f) return 1.0f;
    if (x != x || r != r) return __int_as_float(0x7fc00000);
    if (x == 1.0f) return 1.0f;
    if (x < 0.0f) return __int_as_float(0x7fc00000);
    if (x == 0.0f) return r > 0.0f ? 0.0f : __int_as_float(0x7f800000);
    if (isinf(x)) return r > 0.0f ? __int_as_float(0x7f800000) : 0.0f;
    double y = (double)r * canon_log2((double)x);
    if (y > 999.0) return __int_as_float(0x7f800000);
    if (y < -999.0) return 0.0f;
    return (float)canon_exp2(y);
}

// gain of the regulator for a blurred value m: value / pow(min(m, 1), root); exactly `value` wherever m >= 1.
__device__ __forceinline__ float canon_gain(float m, float value, float root)
{
    float mm = m > 1.0f ? 1.0f : m;
    return __fdiv_rn(value, canon_pow(mm, root));
}

}  // namespace silent
