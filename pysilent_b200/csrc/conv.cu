// K2g: the stand-alone operators of the filter callables -- generic k x k SAME convolution (+relu, +clip), the
// regulator, the border mask and the channel mean. These serve rgc_filter / rgby_filter / orientation_filter /
// apply_filter / regulate_tensor / pad_inwards / get_value_from_color when they are called one by one; the fused
// stack (stack_fused.cu) is the hot path. HBM-bound: each launch reads its input tensor once and writes its output once.
#include "plan.h"

namespace silent {

enum PostOp { kPostNone = 0, kPostRelu = 1, kPostReluClip = 2, kPostRegulate = 3 };

constexpr int kConvTileW = 32, kConvTileH = 8;

// One thread per output pixel, all output channels. Input tile (+halo) and the filter are staged in shared memory.
// Canonical order (oracle/silent_oracle.c:so_conv2d): per output channel one fmaf chain in (ky, kx, ci) order; when
// every input-channel slice of the filter is bitwise identical the chain runs over the channel sum in (ky, kx) order.
__global__ void __launch_bounds__(kConvTileW *kConvTileH)
    conv2d_kernel(const float *__restrict__ x, int h, int w, int cin, const float *__restrict__ filt, int k, int cout,
                  int post, float clip_max, float reg_value, float reg_root, float *__restrict__ out)
{
    extern __shared__ float smem[];
    const int pad = (k - 1) / 2;
    const int tw = kConvTileW + k - 1, th = kConvTileH + k - 1;
    float *s_w = smem;                       // [k*k][cin][cout]
    float *s_x = smem + k * k * cin * cout;  // [th][tw][cin]
    const int tid = threadIdx.y * kConvTileW + threadIdx.x;
    const int nthreads = kConvTileW * kConvTileH;
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * kConvTileW - pad, y0 = blockIdx.y * kConvTileH - pad;
    const float *img = x + (size_t)n * h * w * cin;

    const int nw = k * k * cin * cout;
    for (int i = tid; i < nw; i += nthreads) s_w[i] = __ldg(filt + i);
    const int row_elems = tw * cin;
    for (int i = tid; i < th * row_elems; i += nthreads) {
        const int r = i / row_elems, e = i - r * row_elems;
        const int gy = y0 + r, gx = x0 + e / cin;
        float v = 0.0f;
        if (gy >= 0 && gy < h && gx >= 0 && gx < w) v = __ldg(img + ((size_t)gy * w + gx) * cin + (e % cin));
        s_x[i] = v;
    }
    __syncthreads();

    // Filter structure, decided by the whole CTA: "uniform-in" = every input-channel slice bitwise identical (stripe and
    // blur banks), "uniform-out" = additionally every output channel of a tap identical (blur_tensor). Uniform-in chains
    // run over the channel sum s = ((x0 + x1) + x2 ...), which is then staged ONCE per tile; with uniform-out all output
    // channels share one chain (identical bits), so it is evaluated once: 49 instead of 3136 multiply-adds per pixel for
    // the 8-channel regulator of config C4.
    int same_in = cin >= 2, same_out = 1;
    for (int i = tid; i < nw && (same_in || same_out); i += nthreads) {
        const int co = i % cout, t = i / (cin * cout);
        if (__float_as_uint(s_w[i]) != __float_as_uint(s_w[t * cin * cout + co])) same_in = 0;
        if (__float_as_uint(s_w[i]) != __float_as_uint(s_w[t * cin * cout])) same_out = 0;
    }
    const bool uniform_in = __syncthreads_and(same_in) != 0;
    const bool uniform_out = __syncthreads_and(same_out) != 0 && uniform_in;
    float *s_sum = s_x + th * tw * cin;   // [th][tw] channel sums (uniform-in only)
    if (uniform_in) {
        for (int i = tid; i < th * tw; i += nthreads) {
            const float *px = s_x + i * cin;
            float v = px[0];
            for (int ci = 1; ci < cin; ++ci) v = v + px[ci];
            s_sum[i] = v;
        }
        __syncthreads();
    }

    const int ox = blockIdx.x * kConvTileW + threadIdx.x, oy = blockIdx.y * kConvTileH + threadIdx.y;
    if (ox >= w || oy >= h) return;

    float acc[kMaxChannels];
#pragma unroll
    for (int co = 0; co < kMaxChannels; ++co) acc[co] = 0.0f;

    // chain order (ky, ci, kx): one input row-channel at a time, as in the register-blocked fused kernels
    if (uniform_out) {
        float a = 0.0f;
        for (int ky = 0; ky < k; ++ky)
            for (int kx = 0; kx < k; ++kx)
                a = fmaf(s_w[(ky * k + kx) * cin * cout], s_sum[(threadIdx.y + ky) * tw + threadIdx.x + kx], a);
#pragma unroll
        for (int co = 0; co < kMaxChannels; ++co) acc[co] = a;
    } else {
        for (int ky = 0; ky < k; ++ky) {
            if (uniform_in) {
                for (int kx = 0; kx < k; ++kx) {
                    const float sv = s_sum[(threadIdx.y + ky) * tw + threadIdx.x + kx];
                    const float *pw = s_w + (ky * k + kx) * cin * cout;
#pragma unroll
                    for (int co = 0; co < kMaxChannels; ++co)
                        if (co < cout) acc[co] = fmaf(pw[co], sv, acc[co]);
                }
            } else {
                for (int ci = 0; ci < cin; ++ci) {
                    for (int kx = 0; kx < k; ++kx) {
                        const float v = s_x[((threadIdx.y + ky) * tw + threadIdx.x + kx) * cin + ci];
                        const float *pw = s_w + ((ky * k + kx) * cin + ci) * cout;
#pragma unroll
                        for (int co = 0; co < kMaxChannels; ++co)
                            if (co < cout) acc[co] = fmaf(pw[co], v, acc[co]);
                    }
                }
            }
        }
    }

    float *dst = out + (((size_t)n * h + oy) * w + ox) * cout;
    const float *center = s_x + ((threadIdx.y + pad) * tw + threadIdx.x + pad) * cin;
#pragma unroll
    for (int co = 0; co < kMaxChannels; ++co) {
        if (co >= cout) break;
        float v = acc[co];
        if (post == kPostRelu || post == kPostReluClip) v = canon_relu(v);
        if (post == kPostReluClip) v = canon_clip_hi(v, clip_max);
        if (post == kPostRegulate) v = center[co] * canon_gain(v, reg_value, reg_root);
        dst[co] = v;
    }
}

static int launch_conv(const float *x, int n, int h, int w, int cin, const float *filt, int k, int cout, int post,
                       float clip_max, float reg_value, float reg_root, float *out, cudaStream_t stream,
                       const char *who)
{
    if (!x || !filt || !out) return fail(SILENT_E_INVAL, "%s: null argument", who);
    if (n <= 0 || h <= 0 || w <= 0) return fail(SILENT_E_INVAL, "%s: tensor shape must be positive", who);
    if (cin < 1 || cin > kMaxChannels || cout < 1 || cout > kMaxChannels)
        return fail(SILENT_E_SHAPE, "%s: channel counts must be in 1..%d (got %d -> %d)", who, kMaxChannels, cin, cout);
    if (k < 1 || k > kMaxKernel || (k % 2) == 0)
        return fail(SILENT_E_SHAPE, "%s: filter size must be odd and <= %d (got %d)", who, kMaxKernel, k);
    if (n > 65535) return fail(SILENT_E_SHAPE, "%s: at most 65535 levels per call", who);
    const size_t smem = ((size_t)k * k * cin * cout + (size_t)(kConvTileH + k - 1) * (kConvTileW + k - 1) * (cin + 1)) * 4;
    if (smem > 48 * 1024) {
        SILENT_CUDA(cudaFuncSetAttribute(conv2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 block(kConvTileW, kConvTileH);
    dim3 grid(ceil_div(w, kConvTileW), ceil_div(h, kConvTileH), n);
    conv2d_kernel<<<grid, block, smem, stream>>>(x, h, w, cin, filt, k, cout, post, clip_max, reg_value, reg_root, out);
    SILENT_LAUNCH_CHECK("conv2d_kernel");
    return SILENT_OK;
}

// pad_inwards (util/selection/isolate_rectangle.py:19-23): out = box * x with a 0/1 box.
__global__ void pad_inwards_kernel(const float *__restrict__ x, size_t total, int h, int w, int c, int top, int bottom,
                                   int left, int right, float *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t pix = i / c;
    const int xx = (int)(pix % w), y = (int)((pix / w) % h);
    const bool inside = y >= top && y < h - bottom && xx >= left && xx < w - right;
    const float v = x[i];
    out[i] = inside ? v : (v != v ? v : 0.0f * v);
}

// get_value_from_color (util/color/get_value.py:6-12): ((c0 + c1) + ...) * float32(1 / C).
__global__ void value_kernel(const float *__restrict__ x, size_t pixels, int c, float *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    const float div = __fdiv_rn(1.0f, (float)c);
    float s = x[i * c];
    for (int ch = 1; ch < c; ++ch) s = s + x[i * c + ch];
    out[i] = s * div;
}

}  // namespace silent

using namespace silent;

extern "C" {

int silent_conv2d(const float *x_dev, int n, int h, int w, int cin, const float *filter_hwio_dev, int k, int cout,
                  int post, float clip_max, float *out_dev, silent_stream stream)
{
    if (post < kPostNone || post > kPostReluClip) return fail(SILENT_E_INVAL, "silent_conv2d: post must be 0, 1 or 2");
    return launch_conv(x_dev, n, h, w, cin, filter_hwio_dev, k, cout, post, clip_max, 0.0f, 0.0f, out_dev,
                       (cudaStream_t)stream, "silent_conv2d");
}

int silent_regulate(const float *x_dev, int n, int h, int w, int c, const float *blur_hwio_dev, int k, float value,
                    float root, float *out_dev, silent_stream stream)
{
    return launch_conv(x_dev, n, h, w, c, blur_hwio_dev, k, c, kPostRegulate, 0.0f, value, root, out_dev,
                       (cudaStream_t)stream, "silent_regulate");
}

int silent_pad_inwards(const float *x_dev, int n, int h, int w, int c, int top, int bottom, int left, int right,
                       float *out_dev, silent_stream stream)
{
    if (!x_dev || !out_dev) return fail(SILENT_E_INVAL, "silent_pad_inwards: null argument");
    if (n <= 0 || h <= 0 || w <= 0 || c <= 0) return fail(SILENT_E_INVAL, "silent_pad_inwards: bad shape");
    if (top < 0 || bottom < 0 || left < 0 || right < 0) return fail(SILENT_E_INVAL, "silent_pad_inwards: negative padding");
    const size_t total = (size_t)n * h * w * c;
    pad_inwards_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x_dev, total, h, w, c, top,
                                                                                         bottom, left, right, out_dev);
    SILENT_LAUNCH_CHECK("pad_inwards_kernel");
    return SILENT_OK;
}

int silent_value_from_color(const float *x_dev, int n, int h, int w, int c, float *out_dev, silent_stream stream)
{
    if (!x_dev || !out_dev) return fail(SILENT_E_INVAL, "silent_value_from_color: null argument");
    if (n <= 0 || h <= 0 || w <= 0 || c <= 0) return fail(SILENT_E_INVAL, "silent_value_from_color: bad shape");
    const size_t pixels = (size_t)n * h * w;
    value_kernel<<<(unsigned)((pixels + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x_dev, pixels, c, out_dev);
    SILENT_LAUNCH_CHECK("value_kernel");
    return SILENT_OK;
}

}  // extern "C"
