// K2g: the stand-alone operators of the filter callables -- generic k x k SAME convolution (+relu, +clip), the
// regulator, the border mask and the channel mean. These serve rgc_filter / rgby_filter / orientation_filter /
// apply_filter / regulate_tensor / pad_inwards / get_value_from_color when they are called one by one; the fused
// stack (stack_fused.cu) is the hot path. HBM-bound: each launch reads its input tensor once and writes its output once.
#include "plan.h"

namespace silent {

enum PostOp { kPostNone = 0, kPostRelu = 1, kPostReluClip = 2, kPostRegulate = 3 };

constexpr int kConvTileW = 64, kConvTileH = 16;   // outputs per CTA
constexpr int kConvPX = 4;                        // consecutive output pixels per thread (register blocking)
constexpr int kConvThreads = (kConvTileW / kConvPX) * kConvTileH;

// A thread computes kConvPX consecutive pixels of one row for every output channel: per (ky, ci) it reads the K + 3 input
// values of the row once and feeds all kx taps, pixels and output channels from registers (the first version read one
// shared-memory word per multiply-add). Input tile (+halo) and the filter are staged in shared memory.
// Canonical order (oracle/silent_oracle.c:so_conv2d): per output channel one fmaf chain in (ky, ci, kx) order; when
// every input-channel slice of the filter is bitwise identical the chain runs over the channel sum in (ky, kx) order.
template <int K>
__global__ void __launch_bounds__(kConvThreads)
    conv2d_kernel(const float *__restrict__ x, int h, int w, int cin, const float *__restrict__ filt, int cout, int post,
                  float clip_max, float reg_value, float reg_root, float *__restrict__ out)
{
    extern __shared__ float smem[];
    constexpr int pad = (K - 1) / 2;
    constexpr int tw = kConvTileW + K - 1, th = kConvTileH + K - 1;
    float *s_w = smem;                       // [K*K][cin][cout]
    float *s_x = smem + K * K * cin * cout;  // [th][tw][cin]
    const int tid = threadIdx.x;
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * kConvTileW - pad, y0 = blockIdx.y * kConvTileH - pad;
    const float *img = x + (size_t)n * h * w * cin;

    const int nw = K * K * cin * cout;
    for (int i = tid; i < nw; i += kConvThreads) s_w[i] = __ldg(filt + i);
    // tile fill: a thread walks one tile row at a time (no division per element); the row's tw * cin floats are contiguous
    // in the NHWC tensor
    const int row_elems = tw * cin;
    for (int r = tid / 32; r < th; r += kConvThreads / 32) {
        const int gy = y0 + r;
        const bool row_ok = gy >= 0 && gy < h;
        const float *src = img + ((size_t)(row_ok ? gy : 0) * w + x0) * cin;
        const int lo = x0 < 0 ? -x0 * cin : 0, hi = min(row_elems, (w - x0) * cin);   // elements inside the image
        for (int e = tid % 32; e < row_elems; e += 32)
            s_x[r * row_elems + e] = (row_ok && e >= lo && e < hi) ? __ldg(src + e) : 0.0f;
    }
    __syncthreads();

    // Filter structure, decided by the whole CTA: "uniform-in" = every input-channel slice bitwise identical (stripe and
    // blur banks), "uniform-out" = additionally every output channel of a tap identical (blur_tensor). Uniform-in chains
    // run over the channel sum s = ((x0 + x1) + x2 ...), which is then staged ONCE per tile; with uniform-out all output
    // channels share one chain (identical bits), so it is evaluated once.
    int same_in = cin >= 2, same_out = 1;
    for (int i = tid; i < nw && (same_in || same_out); i += kConvThreads) {
        const int co = i % cout, t = i / (cin * cout);
        if (__float_as_uint(s_w[i]) != __float_as_uint(s_w[t * cin * cout + co])) same_in = 0;
        if (__float_as_uint(s_w[i]) != __float_as_uint(s_w[t * cin * cout])) same_out = 0;
    }
    const bool uniform_in = __syncthreads_and(same_in) != 0;
    const bool uniform_out = __syncthreads_and(same_out) != 0 && uniform_in;
    float *s_sum = s_x + th * tw * cin;   // [th][tw] channel sums (uniform-in only)
    if (uniform_in) {
        for (int i = tid; i < th * tw; i += kConvThreads) {
            const float *px = s_x + i * cin;
            float v = px[0];
            for (int ci = 1; ci < cin; ++ci) v = v + px[ci];
            s_sum[i] = v;
        }
        __syncthreads();
    }

    const int lx = (tid % (kConvTileW / kConvPX)) * kConvPX, ly = tid / (kConvTileW / kConvPX);
    const int ox = blockIdx.x * kConvTileW + lx, oy = blockIdx.y * kConvTileH + ly;
    if (ox >= w || oy >= h) return;

    float acc[kConvPX][kMaxChannels];
#pragma unroll
    for (int p = 0; p < kConvPX; ++p)
#pragma unroll
        for (int co = 0; co < kMaxChannels; ++co) acc[p][co] = 0.0f;

    if (uniform_out) {
        float a[kConvPX];
#pragma unroll
        for (int p = 0; p < kConvPX; ++p) a[p] = 0.0f;
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            float v[kConvPX + K - 1];
#pragma unroll
            for (int j = 0; j < kConvPX + K - 1; ++j) v[j] = s_sum[(ly + ky) * tw + lx + j];
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
                const float wt = s_w[(ky * K + kx) * cin * cout];
#pragma unroll
                for (int p = 0; p < kConvPX; ++p) a[p] = fmaf(wt, v[p + kx], a[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < kConvPX; ++p)
#pragma unroll
            for (int co = 0; co < kMaxChannels; ++co) acc[p][co] = a[p];
    } else {
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
            const int nci = uniform_in ? 1 : cin;
            for (int ci = 0; ci < nci; ++ci) {
                float v[kConvPX + K - 1];
#pragma unroll
                for (int j = 0; j < kConvPX + K - 1; ++j)
                    v[j] = uniform_in ? s_sum[(ly + ky) * tw + lx + j] : s_x[((ly + ky) * tw + lx + j) * cin + ci];
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    const float *pw = s_w + ((ky * K + kx) * cin + ci) * cout;
#pragma unroll
                    for (int co = 0; co < kMaxChannels; ++co) {
                        if (co < cout) {
                            const float wt = pw[co];
#pragma unroll
                            for (int p = 0; p < kConvPX; ++p) acc[p][co] = fmaf(wt, v[p + kx], acc[p][co]);
                        }
                    }
                }
            }
        }
    }

#pragma unroll
    for (int p = 0; p < kConvPX; ++p) {
        if (ox + p >= w) break;
        float *dst = out + (((size_t)n * h + oy) * w + ox + p) * cout;
        const float *center = s_x + ((ly + pad) * tw + lx + p + pad) * cin;
#pragma unroll
        for (int co = 0; co < kMaxChannels; ++co) {
            if (co >= cout) break;
            float v = acc[p][co];
            if (post == kPostRelu || post == kPostReluClip) v = canon_relu(v);
            if (post == kPostReluClip) v = canon_clip_hi(v, clip_max);
            if (post == kPostRegulate) v = center[co] * canon_gain(v, reg_value, reg_root);
            dst[co] = v;
        }
    }
}

template <int K>
static int launch_conv_k(const float *x, int n, int h, int w, int cin, const float *filt, int cout, int post,
                         float clip_max, float reg_value, float reg_root, float *out, cudaStream_t stream)
{
    const size_t smem = ((size_t)K * K * cin * cout + (size_t)(kConvTileH + K - 1) * (kConvTileW + K - 1) * (cin + 1)) * 4;
    SILENT_CUDA(cudaFuncSetAttribute(conv2d_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid(ceil_div(w, kConvTileW), ceil_div(h, kConvTileH), n);
    conv2d_kernel<K><<<grid, kConvThreads, smem, stream>>>(x, h, w, cin, filt, cout, post, clip_max, reg_value, reg_root, out);
    SILENT_LAUNCH_CHECK("conv2d_kernel");
    return SILENT_OK;
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
    switch (k) {
        case 1: return launch_conv_k<1>(x, n, h, w, cin, filt, cout, post, clip_max, reg_value, reg_root, out, stream);
        case 3: return launch_conv_k<3>(x, n, h, w, cin, filt, cout, post, clip_max, reg_value, reg_root, out, stream);
        case 5: return launch_conv_k<5>(x, n, h, w, cin, filt, cout, post, clip_max, reg_value, reg_root, out, stream);
        default: return launch_conv_k<7>(x, n, h, w, cin, filt, cout, post, clip_max, reg_value, reg_root, out, stream);
    }
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
