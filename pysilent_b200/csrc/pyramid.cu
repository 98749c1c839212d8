// K1: foveated pyramid build = zoom.from_image (reference util/zoom/from_image.py:48-64) for a batch of frames.
//
// One thread per output pixel (all colours). Canonical order: for each of the 6 y-taps, r_j = fmaf chain over the 6
// x-taps of the source row; then an fmaf chain over the r_j with the y weights. Tap tables come from the plan.
// Memory-wise a level-s tile reads a (scale^s)-times larger frame patch; rows are re-read by neighbouring threads
// through L1/L2, HBM sees each frame byte roughly once.
#include "plan.h"

namespace silent {

template <typename T>
__device__ __forceinline__ float load_sample(const T *p);
template <>
__device__ __forceinline__ float load_sample<uint8_t>(const uint8_t *p) { return (float)__ldg(p); }
template <>
__device__ __forceinline__ float load_sample<float>(const float *p) { return __ldg(p); }

template <typename T, int NC>
__global__ void __launch_bounds__(256) pyramid_kernel(const T *__restrict__ frames, float *__restrict__ out,
                                                      const int32_t *__restrict__ idx_y, const float *__restrict__ w_y,
                                                      const int32_t *__restrict__ idx_x, const float *__restrict__ w_x,
                                                      int L, int h, int w, int H, int W, int FC)
{
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y * blockDim.y + threadIdx.y;
    const int n = blockIdx.z;   // frame * L + level
    if (ox >= w || oy >= h) return;
    const int s = n % L, b = n / L;
    const int32_t *ty = idx_y + ((size_t)s * h + oy) * kTaps;
    const int32_t *tx = idx_x + ((size_t)s * w + ox) * kTaps;
    const float *gy = w_y + ((size_t)s * h + oy) * kTaps;
    const float *gx = w_x + ((size_t)s * w + ox) * kTaps;
    float *dst = out + (((size_t)n * h + oy) * w + ox) * NC;

    float acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0f;

    if (__ldg(ty) >= 0 && __ldg(tx) >= 0) {
        int cx[kTaps];
        float wx[kTaps];
#pragma unroll
        for (int i = 0; i < kTaps; ++i) {
            cx[i] = __ldg(tx + i) * FC;
            wx[i] = __ldg(gx + i);
        }
        const T *frame = frames + (size_t)b * H * W * FC;
#pragma unroll
        for (int j = 0; j < kTaps; ++j) {
            const T *row = frame + (size_t)__ldg(ty + j) * W * FC;
            const float wyj = __ldg(gy + j);
            float r[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) r[c] = 0.0f;
#pragma unroll
            for (int i = 0; i < kTaps; ++i) {
#pragma unroll
                for (int c = 0; c < NC; ++c) r[c] = fmaf(wx[i], load_sample<T>(row + cx[i] + c), r[c]);
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[c] = fmaf(wyj, r[c], acc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) dst[c] = acc[c];
}

template <typename T>
static int launch_pyramid(const silent_plan *plan, const T *frames, int batch, float *out, cudaStream_t stream)
{
    const int L = plan->levels, h = plan->h, w = plan->w;
    const silent_params &p = plan->params;
    dim3 block(32, 8);
    dim3 grid(ceil_div(w, 32), ceil_div(h, 8), batch * L);
#define SILENT_PYR_CASE(NC)                                                                                           \
    case NC:                                                                                                          \
        pyramid_kernel<T, NC><<<grid, block, 0, stream>>>(frames, out, plan->d_idx_y, plan->d_w_y, plan->d_idx_x,     \
                                                          plan->d_w_x, L, h, w, p.frame_h, p.frame_w, p.frame_c);     \
        break;
    switch (p.num_colors) {
        SILENT_PYR_CASE(1)
        SILENT_PYR_CASE(2)
        SILENT_PYR_CASE(3)
        SILENT_PYR_CASE(4)
        SILENT_PYR_CASE(5)
        SILENT_PYR_CASE(6)
        SILENT_PYR_CASE(7)
        SILENT_PYR_CASE(8)
        default:
            return fail(SILENT_E_SHAPE, "num_colors=%d not supported (1..8)", p.num_colors);
    }
#undef SILENT_PYR_CASE
    SILENT_LAUNCH_CHECK("pyramid_kernel");
    return SILENT_OK;
}

int pyramid_build(const silent_plan *plan, const void *frames_dev, int batch, float *pyramid_dev, cudaStream_t stream)
{
    if (!plan || !frames_dev || !pyramid_dev) return fail(SILENT_E_INVAL, "silent_pyramid_build: null argument");
    if (batch <= 0) return fail(SILENT_E_INVAL, "batch must be positive, got %d", batch);
    if (plan->levels == 0) return SILENT_OK;
    if (!plan->on_device) return fail(SILENT_E_CUDA, "plan was created without a CUDA device; no CPU fallback exists");
    if ((long long)batch * plan->levels > 65535) return fail(SILENT_E_SHAPE, "batch * levels must be <= 65535");
    if (plan->params.frame_dtype == SILENT_U8)
        return launch_pyramid<uint8_t>(plan, (const uint8_t *)frames_dev, batch, pyramid_dev, stream);
    return launch_pyramid<float>(plan, (const float *)frames_dev, batch, pyramid_dev, stream);
}

}  // namespace silent

extern "C" int silent_pyramid_build(const silent_plan *plan, const void *frames_dev, int batch, float *pyramid_dev,
                                    silent_stream stream)
{
    return silent::pyramid_build(plan, frames_dev, batch, pyramid_dev, (cudaStream_t)stream);
}
