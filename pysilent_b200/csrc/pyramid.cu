// K1: foveated pyramid build = zoom.from_image (reference util/zoom/from_image.py:48-64) for a batch of frames:
// per level, a centred crop resampled to h x w with the order-5 spline of scipy.ndimage.zoom(prefilter=False).
//
// Canonical order (oracle/silent_oracle.c:so_pyramid): separable, VERTICAL pass first -- for each of the 6 source
// columns an fmaf chain over the 6 y-taps -- then the fmaf chain over the 6 x-taps.
//
// Two kernels:
//  * pyramid_kernel      one thread per output pixel, NHWC float32 output: the stand-alone from_image operator, any
//                        channel count / frame dtype.
//  * pyramid_pair_kernel the pipeline's producer. One CTA = one tile of one level for TWO frames in lockstep (float2
//                        lanes, like the stack kernels). Phase V walks ALIGNED 32-bit words of the uint8 frame down the
//                        6 tap rows (coalesced, 4 byte-columns per load), converts with PRMT + FADD2 (the I2F pipe is
//                        16/clk/SM on B200: measured 2.4x slower), accumulates with FFMA2 and parks the column sums in
//                        shared memory; phase H gathers 6 of them per output sample. ~10x fewer instructions than the
//                        per-pixel kernel (which re-converts every tap for every pixel). Output is the pair-interleaved
//                        planar layout xpair[pair][c][y][x] = (frame A, frame B) that stack_a_kernel loads verbatim.
#include "plan.h"
#include "stack.h"

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
        const T *frame = frames + (size_t)b * H * W * FC;
        size_t row_off[kTaps];
        float wy[kTaps];
#pragma unroll
        for (int j = 0; j < kTaps; ++j) {
            row_off[j] = (size_t)__ldg(ty + j) * W * FC;
            wy[j] = __ldg(gy + j);
        }
#pragma unroll
        for (int i = 0; i < kTaps; ++i) {
            const int col = __ldg(tx + i) * FC;
            const float wxi = __ldg(gx + i);
            float t[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) t[c] = 0.0f;
#pragma unroll
            for (int j = 0; j < kTaps; ++j) {
#pragma unroll
                for (int c = 0; c < NC; ++c) t[c] = fmaf(wy[j], load_sample<T>(frame + row_off[j] + col + c), t[c]);
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[c] = fmaf(wxi, t[c], acc[c]);
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

// ---------------------------------------------------------------------------------------------------------------------
// frame-pair kernel
// ---------------------------------------------------------------------------------------------------------------------

typedef float2 f2;
constexpr int kPairThreads = 256;
constexpr int kVGroup = 18;   // float2 slots per group of 16 byte-columns in the column-sum buffer (16 + 2 of padding)

struct PairParams {
    const uint8_t *frames;
    f2 *xpair;
    const int32_t *idx_y, *idx_x;   // tables of THIS level: [h][6], [w][6]
    const float *w_y, *w_x;
    int H, row_bytes, FC;           // frame rows, bytes per frame row, interleaved channels
    int h, w, L, level, B;
    int th;                         // output rows per tile
    int vpitch;                     // float2 per row of the column-sum buffer
    int word_lo[kPairMaxTiles], nwords[kPairMaxTiles];   // per x tile: first 32-bit word of a frame row, word count
};

// byte k of `word` as an exact float: PRMT builds the bit pattern of 2^23 + byte, the caller subtracts 2^23 (FADD2)
__device__ __forceinline__ float magic_byte(uint32_t word, uint32_t selector)
{
    return __uint_as_float(__byte_perm(word, 0x4B000000u, selector));
}

__global__ void __launch_bounds__(kPairThreads) pyramid_pair_kernel(const __grid_constant__ PairParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    f2 *sV = reinterpret_cast<f2 *>(smem_raw);                          // [th][vpitch] column sums (frame A, frame B)
    int *sTy = reinterpret_cast<int *>(sV + (size_t)P.th * P.vpitch);   // [th][6] source rows (-1: row is zero)
    float *sWy = reinterpret_cast<float *>(sTy + P.th * kTaps);         // [th][6]

    const int tid = threadIdx.x;
    const int bx = blockIdx.x, by = blockIdx.y, q = blockIdx.z;
    const int th = P.th, h = P.h, w = P.w;
    const int oy0 = by * th;
    const size_t frame_bytes = (size_t)P.H * P.row_bytes;
    const uint8_t *frameA = P.frames + (size_t)(2 * q) * frame_bytes;
    const uint8_t *frameB = (2 * q + 1 < P.B) ? frameA + frame_bytes : frameA;

    for (int i = tid; i < th * kTaps; i += kPairThreads) {
        const int oy = oy0 + i / kTaps;
        const bool ok = oy < h && __ldg(P.idx_y + (size_t)oy * kTaps) >= 0;
        sTy[i] = ok ? __ldg(P.idx_y + (size_t)oy * kTaps + i % kTaps) : -1;
        sWy[i] = ok ? __ldg(P.w_y + (size_t)oy * kTaps + i % kTaps) : 0.0f;
    }
    __syncthreads();

    // ---- phase V: column sums over the 6 y-taps, 16 byte-columns (one aligned 128-bit load per tap row and frame) per
    //      task: 12 independent 16-byte loads in flight per thread cover the HBM/L2 latency ---------------------------
    const int nw = P.nwords[bx], wlo = P.word_lo[bx];   // multiples of 4 words
    const int nq = nw >> 2;
    const f2 bias = make_float2(-8388608.0f, -8388608.0f);
    for (int t = tid; t < th * nq; t += kPairThreads) {
        const int r = t / nq, qi = t - r * nq;
        f2 acc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = make_float2(0.0f, 0.0f);
        if (sTy[r * kTaps] >= 0) {
            uint4 qa[kTaps], qb[kTaps];
#pragma unroll
            for (int j = 0; j < kTaps; ++j) {
                const size_t off = (size_t)sTy[r * kTaps + j] * P.row_bytes;
                qa[j] = __ldg(reinterpret_cast<const uint4 *>(frameA + off) + (wlo >> 2) + qi);
                qb[j] = __ldg(reinterpret_cast<const uint4 *>(frameB + off) + (wlo >> 2) + qi);
            }
#pragma unroll
            for (int j = 0; j < kTaps; ++j) {
                const float wy = sWy[r * kTaps + j];
                const f2 wy2 = make_float2(wy, wy);
                const uint32_t wa[4] = {qa[j].x, qa[j].y, qa[j].z, qa[j].w}, wb[4] = {qb[j].x, qb[j].y, qb[j].z, qb[j].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    acc[4 * k + 0] = __ffma2_rn(wy2, __fadd2_rn(make_float2(magic_byte(wa[k], 0x7440), magic_byte(wb[k], 0x7440)), bias), acc[4 * k + 0]);
                    acc[4 * k + 1] = __ffma2_rn(wy2, __fadd2_rn(make_float2(magic_byte(wa[k], 0x7441), magic_byte(wb[k], 0x7441)), bias), acc[4 * k + 1]);
                    acc[4 * k + 2] = __ffma2_rn(wy2, __fadd2_rn(make_float2(magic_byte(wa[k], 0x7442), magic_byte(wb[k], 0x7442)), bias), acc[4 * k + 2]);
                    acc[4 * k + 3] = __ffma2_rn(wy2, __fadd2_rn(make_float2(magic_byte(wa[k], 0x7443), magic_byte(wb[k], 0x7443)), bias), acc[4 * k + 3]);
                }
            }
        }
        // 16 columns = 128 B per task: padded to 144 B so consecutive lanes start in consecutive 16-byte bank groups
        float4 *dst = reinterpret_cast<float4 *>(sV + (size_t)r * P.vpitch + kVGroup * qi);
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[k] = make_float4(acc[2 * k].x, acc[2 * k].y, acc[2 * k + 1].x, acc[2 * k + 1].y);
    }
    __syncthreads();

    // ---- phase H: chain over the 6 x-taps; a thread owns one output column and walks every 4th row of the tile --------
    constexpr int ROW_GROUPS = kPairThreads / kPairTileW;
    const int ox = bx * kPairTileW + (tid % kPairTileW);
    if (ox >= w) return;
    const int32_t *tx = P.idx_x + (size_t)ox * kTaps;
    const bool col_ok = __ldg(tx) >= 0;
    int bc[kTaps];
    f2 wx[kTaps];
#pragma unroll
    for (int i = 0; i < kTaps; ++i) {
        bc[i] = col_ok ? __ldg(tx + i) * P.FC - 4 * wlo : 0;
        const float v = col_ok ? __ldg(P.w_x + (size_t)ox * kTaps + i) : 0.0f;
        wx[i] = make_float2(v, v);
    }
    const size_t plane = (size_t)h * w;
    f2 *out = P.xpair + ((size_t)q * P.L + P.level) * 3 * plane + ox;
    for (int r = tid / kPairTileW; r < th; r += ROW_GROUPS) {
        const int oy = oy0 + r;
        if (oy >= h) break;
        const f2 *row = sV + (size_t)r * P.vpitch;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            f2 acc = make_float2(0.0f, 0.0f);
            if (col_ok) {
#pragma unroll
                for (int i = 0; i < kTaps; ++i) {
                    const int b = bc[i] + c;
                    acc = __ffma2_rn(wx[i], row[b + (kVGroup - 16) * (b >> 4)], acc);
                }
            }
            out[c * plane + (size_t)oy * w] = acc;
        }
    }
}

// Can the pair kernel serve this plan? (uint8 frames, 3 colours, word-aligned rows, table of tiles fits)
bool pyramid_pair_supported(const silent_plan *plan)
{
    const silent_params &p = plan->params;
    return plan->on_device && p.frame_dtype == SILENT_U8 && p.num_colors == 3 && (p.frame_w * p.frame_c) % 16 == 0 &&
           plan->pair_ok;
}

size_t pyramid_pair_bytes(const silent_plan *plan, int batch)
{
    return (size_t)((batch + 1) / 2) * plan->levels * 3 * plan->h * plan->w * sizeof(f2);
}

int pyramid_pair_build(const silent_plan *plan, const void *frames_dev, int batch, void *xpair_dev, cudaStream_t stream)
{
    if (!pyramid_pair_supported(plan)) return fail(SILENT_E_SHAPE, "frame-pair pyramid kernel does not support this plan");
    if (((uintptr_t)frames_dev & 15) != 0) return fail(SILENT_E_INVAL, "frames must be 16-byte aligned");
    const silent_params &p = plan->params;
    const int pairs = (batch + 1) / 2;
    if (pairs > 65535) return fail(SILENT_E_SHAPE, "at most 131070 frames per call");
    static bool configured = false;
    if (!configured) {
        SILENT_CUDA(cudaFuncSetAttribute(pyramid_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    for (int s = 0; s < plan->levels; ++s) {
        const PairLevel &pl = plan->pair[s];
        PairParams P;
        P.frames = (const uint8_t *)frames_dev;
        P.xpair = (f2 *)xpair_dev;
        P.idx_y = plan->d_idx_y + (size_t)s * plan->h * kTaps;
        P.w_y = plan->d_w_y + (size_t)s * plan->h * kTaps;
        P.idx_x = plan->d_idx_x + (size_t)s * plan->w * kTaps;
        P.w_x = plan->d_w_x + (size_t)s * plan->w * kTaps;
        P.H = p.frame_h;
        P.row_bytes = p.frame_w * p.frame_c;
        P.FC = p.frame_c;
        P.h = plan->h, P.w = plan->w, P.L = plan->levels, P.level = s, P.B = batch;
        P.th = pl.th;
        P.vpitch = pl.vpitch;
        for (int t = 0; t < pl.ntx; ++t) P.word_lo[t] = pl.word_lo[t], P.nwords[t] = pl.nwords[t];
        const size_t smem = (size_t)pl.th * pl.vpitch * sizeof(f2) + (size_t)pl.th * kTaps * 8;
        dim3 grid(pl.ntx, ceil_div(plan->h, pl.th), pairs);
        pyramid_pair_kernel<<<grid, kPairThreads, smem, stream>>>(P);
        SILENT_LAUNCH_CHECK("pyramid_pair_kernel");
    }
    return SILENT_OK;
}

}  // namespace silent

extern "C" int silent_pyramid_build(const silent_plan *plan, const void *frames_dev, int batch, float *pyramid_dev,
                                    silent_stream stream)
{
    return silent::pyramid_build(plan, frames_dev, batch, pyramid_dev, (cudaStream_t)stream);
}
