// K2: the fused stack S1-S7 of LineEndDisplayer.compile (reference recognition_testing.py:69-77) in ONE kernel:
//   rgc 3x3 + relu -> rgby 3x3 + relu -> stripe 3x3 + relu -> 7x7 blur regulator -> end 3x3 + relu + clip ->
//   border mask -> channel mean.
// Each CTA owns one TH x TW tile of one pyramid level; the input tile (+7 halo) is read from HBM once, every
// intermediate lives in shared memory, and only orient / line_end / gray are written back. All filter weights arrive as
// a kernel parameter (constant bank), so there is no global state and no weight traffic.
//
// Structure used (validated on the host, else SILENT_E_STRUCTURE): the stripe filter is identical over its input
// channels (3x3x3x3 -> three 3x3 kernels on the channel sum) and the blur filter is one 7x7 kernel in every slice
// (441 MAC -> 49). Both hold for every filter the reference's generators produce. Evaluation order = canonical order of
// oracle/silent_oracle.c, so results are bit-identical to it.
#include <cstring>

#include "plan.h"

namespace silent {

struct StackParams {
    float w1[9][3][3];   // rgc    [tap][ci][co]
    float w2[9][3][3];   // rgby
    float w3[9][3];      // stripe [tap][co]   (uniform over ci)
    float wb[49];        // blur   [tap]       (uniform over ci, co)
    float w5[9][3][3];   // end
    float reg_value, reg_root, clip_max;
    int border;
    int h, w;
};

constexpr int kStackThreads = 256;

template <int TH, int TW>
struct StackTile {
    static constexpr int XH = TH + 14, XW = TW + 14;   // input            (halo 7)
    static constexpr int AH = TH + 12, AW = TW + 12;   // rgc              (halo 6)
    static constexpr int BH = TH + 10, BW = TW + 10;   // rgby channel sum (halo 5)
    static constexpr int CH = TH + 8, CW = TW + 8;     // stripe           (halo 4)
    static constexpr int DH = TH + 2, DW = TW + 2;     // orient           (halo 1)
    static constexpr int kBufX = XH * XW * 3;          // X, later C
    static constexpr int kBufA = AH * AW * 3;          // A, later Csum + D
    static constexpr int kBufB = BH * BW;              // Bsum
    static constexpr size_t kSmemBytes = (size_t)(kBufX + kBufA + kBufB) * sizeof(float);
    static_assert(CH * CW * 3 <= kBufX, "stripe tile must fit in the input buffer");
    static_assert(CH * CW + DH * DW * 3 <= kBufA, "channel sum + orient tile must fit in the rgc buffer");
};

template <int TH, int TW>
__global__ void __launch_bounds__(kStackThreads, 2)
    stack_fused_kernel(const float *__restrict__ pyr, const __grid_constant__ StackParams P, float *__restrict__ orient,
                       float *__restrict__ line_end, float *__restrict__ gray)
{
    using T = StackTile<TH, TW>;
    extern __shared__ float smem[];
    float *sX = smem;                 // [XH][XW][3]
    float *sA = smem + T::kBufX;      // [AH][AW][3]
    float *sB = sA + T::kBufA;        // [BH][BW]
    float *sC = sX;                   // [CH][CW][3]   (X is dead once A exists)
    float *sCs = sA;                  // [CH][CW]      (A is dead once Bsum exists)
    float *sD = sA + T::CH * T::CW;   // [DH][DW][3]

    const int tid = threadIdx.x;
    const int n = blockIdx.z;
    const int ty0 = blockIdx.y * TH, tx0 = blockIdx.x * TW;
    const int h = P.h, w = P.w;
    const float *img = pyr + (size_t)n * h * w * 3;

    // ---- load input tile, zero outside the level (SAME padding of S1) -------------------------------------------------
    for (int i = tid; i < T::XH * T::XW * 3; i += kStackThreads) {
        const int r = i / (T::XW * 3), e = i - r * (T::XW * 3);
        const int gy = ty0 - 7 + r, gx = tx0 - 7 + e / 3;
        float v = 0.0f;
        if (gy >= 0 && gy < h && gx >= 0 && gx < w) v = __ldg(img + ((size_t)gy * w + gx) * 3 + (e % 3));
        sX[i] = v;
    }
    __syncthreads();

    // ---- S1: a = relu(conv3x3(x, rgc))                                                          filters/rgc.py:13-16
    for (int i = tid; i < T::AH * T::AW; i += kStackThreads) {
        const int r = i / T::AW, c = i - r * T::AW;
        const int gy = ty0 - 6 + r, gx = tx0 - 6 + c;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
        if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float *px = sX + ((r + t / 3) * T::XW + c + t % 3) * 3;
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const float v = px[ci];
                    a0 = fmaf(P.w1[t][ci][0], v, a0);
                    a1 = fmaf(P.w1[t][ci][1], v, a1);
                    a2 = fmaf(P.w1[t][ci][2], v, a2);
                }
            }
            a0 = canon_relu(a0), a1 = canon_relu(a1), a2 = canon_relu(a2);
        }
        sA[i * 3 + 0] = a0, sA[i * 3 + 1] = a1, sA[i * 3 + 2] = a2;
    }
    __syncthreads();

    // ---- S2: b = relu(conv3x3(a, rgby)); only the channel sum is needed downstream              filters/rgby.py:11-12
    for (int i = tid; i < T::BH * T::BW; i += kStackThreads) {
        const int r = i / T::BW, c = i - r * T::BW;
        const int gy = ty0 - 5 + r, gx = tx0 - 5 + c;
        float s = 0.0f;
        if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
            float b0 = 0.0f, b1 = 0.0f, b2 = 0.0f;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float *pa = sA + ((r + t / 3) * T::AW + c + t % 3) * 3;
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const float v = pa[ci];
                    b0 = fmaf(P.w2[t][ci][0], v, b0);
                    b1 = fmaf(P.w2[t][ci][1], v, b1);
                    b2 = fmaf(P.w2[t][ci][2], v, b2);
                }
            }
            s = (canon_relu(b0) + canon_relu(b1)) + canon_relu(b2);
        }
        sB[i] = s;
    }
    __syncthreads();

    // ---- S3: c = relu(conv3x3(b, stripe)) on the channel sum                              filters/orientation.py:24-29
    for (int i = tid; i < T::CH * T::CW; i += kStackThreads) {
        const int r = i / T::CW, c = i - r * T::CW;
        const int gy = ty0 - 4 + r, gx = tx0 - 4 + c;
        float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f;
        if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float v = sB[(r + t / 3) * T::BW + c + t % 3];
                c0 = fmaf(P.w3[t][0], v, c0);
                c1 = fmaf(P.w3[t][1], v, c1);
                c2 = fmaf(P.w3[t][2], v, c2);
            }
            c0 = canon_relu(c0), c1 = canon_relu(c1), c2 = canon_relu(c2);
        }
        sC[i * 3 + 0] = c0, sC[i * 3 + 1] = c1, sC[i * 3 + 2] = c2;
        sCs[i] = (c0 + c1) + c2;
    }
    __syncthreads();

    // ---- S4: d = c * (value / pow(min(blur7x7(c), 1), root))                 regulator/gaussian_regulator_tensor.py:34-36
    for (int i = tid; i < T::DH * T::DW; i += kStackThreads) {
        const int r = i / T::DW, c = i - r * T::DW;
        const int gy = ty0 - 1 + r, gx = tx0 - 1 + c;
        float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
        if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
            float m = 0.0f;
#pragma unroll
            for (int ky = 0; ky < 7; ++ky)
#pragma unroll
                for (int kx = 0; kx < 7; ++kx) m = fmaf(P.wb[ky * 7 + kx], sCs[(r + ky) * T::CW + c + kx], m);
            const float gain = canon_gain(m, P.reg_value, P.reg_root);
            const float *pc = sC + ((r + 3) * T::CW + c + 3) * 3;
            d0 = pc[0] * gain, d1 = pc[1] * gain, d2 = pc[2] * gain;
        }
        sD[i * 3 + 0] = d0, sD[i * 3 + 1] = d1, sD[i * 3 + 2] = d2;
    }
    __syncthreads();

    // ---- orient output (coalesced rows of the central TH x TW)
    if (orient) {
        float *dst = orient + (size_t)n * h * w * 3;
        for (int i = tid; i < TH * TW * 3; i += kStackThreads) {
            const int r = i / (TW * 3), e = i - r * (TW * 3);
            const int gy = ty0 + r, gx = tx0 + e / 3;
            if (gy < h && gx < w) dst[((size_t)gy * w + gx) * 3 + e % 3] = sD[((r + 1) * T::DW + 1) * 3 + e];
        }
    }

    // ---- S5-S7: e = clip(relu(conv3x3(d, end))); p = mask * e; g = mean(p)                recognition_testing.py:73-77
    for (int i = tid; i < TH * TW; i += kStackThreads) {
        const int r = i / TW, c = i - r * TW;
        const int gy = ty0 + r, gx = tx0 + c;
        if (gy >= h || gx >= w) continue;
        float e0 = 0.0f, e1 = 0.0f, e2 = 0.0f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float *pd = sD + ((r + t / 3) * T::DW + c + t % 3) * 3;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float v = pd[ci];
                e0 = fmaf(P.w5[t][ci][0], v, e0);
                e1 = fmaf(P.w5[t][ci][1], v, e1);
                e2 = fmaf(P.w5[t][ci][2], v, e2);
            }
        }
        e0 = canon_clip_hi(canon_relu(e0), P.clip_max);
        e1 = canon_clip_hi(canon_relu(e1), P.clip_max);
        e2 = canon_clip_hi(canon_relu(e2), P.clip_max);
        const bool inside = gy >= P.border && gy < h - P.border && gx >= P.border && gx < w - P.border;
        if (!inside) {
            e0 = e0 != e0 ? e0 : 0.0f * e0;
            e1 = e1 != e1 ? e1 : 0.0f * e1;
            e2 = e2 != e2 ? e2 : 0.0f * e2;
        }
        const size_t pix = ((size_t)n * h + gy) * w + gx;
        if (line_end) {
            line_end[pix * 3 + 0] = e0;
            line_end[pix * 3 + 1] = e1;
            line_end[pix * 3 + 2] = e2;
        }
        if (gray) gray[pix] = ((e0 + e1) + e2) * __fdiv_rn(1.0f, 3.0f);
    }
}

static bool bits_equal(float a, float b) { return std::memcmp(&a, &b, 4) == 0; }

// Validate the structure the fused kernel relies on and repack the HWIO filters.
int pack_stack_params(const silent_stack_weights *W, int h, int w, StackParams *P)
{
    if (!W) return fail(SILENT_E_INVAL, "null weights");
    for (int t = 0; t < 9; ++t)
        for (int ci = 0; ci < 3; ++ci)
            for (int co = 0; co < 3; ++co) {
                P->w1[t][ci][co] = W->rgc[(t * 3 + ci) * 3 + co];
                P->w2[t][ci][co] = W->rgby[(t * 3 + ci) * 3 + co];
                P->w5[t][ci][co] = W->end[(t * 3 + ci) * 3 + co];
                if (!bits_equal(W->stripe[(t * 3 + ci) * 3 + co], W->stripe[(t * 3) * 3 + co]))
                    return fail(SILENT_E_STRUCTURE, "stripe filter differs across input channels at tap %d; the fused "
                                                    "stack needs identical input slices (use the per-operator calls)", t);
                P->w3[t][co] = W->stripe[(t * 3) * 3 + co];
            }
    for (int t = 0; t < 49; ++t) {
        for (int s = 0; s < 9; ++s)
            if (!bits_equal(W->blur[t * 9 + s], W->blur[t * 9]))
                return fail(SILENT_E_STRUCTURE, "blur filter slices differ at tap %d; the fused stack needs one 7x7 "
                                                "kernel in every slice (use the per-operator calls)", t);
        P->wb[t] = W->blur[t * 9];
    }
    if (W->border < 0) return fail(SILENT_E_INVAL, "border must be >= 0");
    P->reg_value = W->regulation_value;
    P->reg_root = W->regulation_root;
    P->clip_max = W->clip_max;
    P->border = W->border;
    P->h = h;
    P->w = w;
    return SILENT_OK;
}

int stack_fused(const float *pyr, int n, int h, int w, const silent_stack_weights *W, float *orient, float *line_end,
                float *gray, cudaStream_t stream)
{
    if (!pyr) return fail(SILENT_E_INVAL, "silent_stack_fused: null pyramid");
    if (n <= 0 || h <= 0 || w <= 0) return fail(SILENT_E_INVAL, "silent_stack_fused: bad shape %dx%dx%d", n, h, w);
    if (n > 65535) return fail(SILENT_E_SHAPE, "silent_stack_fused: at most 65535 levels per call");
    StackParams P;
    int rc = pack_stack_params(W, h, w, &P);
    if (rc != SILENT_OK) return rc;
    constexpr int TH = 32, TW = 64;
    using T = StackTile<TH, TW>;
    static bool configured = false;
    if (!configured) {
        SILENT_CUDA(cudaFuncSetAttribute(stack_fused_kernel<TH, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)T::kSmemBytes));
        configured = true;
    }
    dim3 grid(ceil_div(w, TW), ceil_div(h, TH), n);
    stack_fused_kernel<TH, TW><<<grid, kStackThreads, T::kSmemBytes, stream>>>(pyr, P, orient, line_end, gray);
    SILENT_LAUNCH_CHECK("stack_fused_kernel");
    return SILENT_OK;
}

}  // namespace silent

extern "C" int silent_stack_fused(const float *pyramid_dev, int n, int h, int w, const silent_stack_weights *weights_host,
                                  float *orient_dev, float *line_end_dev, float *gray_dev, silent_stream stream)
{
    return silent::stack_fused(pyramid_dev, n, h, w, weights_host, orient_dev, line_end_dev, gray_dev,
                               (cudaStream_t)stream);
}
