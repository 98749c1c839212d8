/*
 * TEST INFRASTRUCTURE ONLY -- bit-defined float32 CPU restatement of pySILEnT's filter-pipeline hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may build, load or call this file. The product
 * (pysilent_b200 + libsilent_b200.so) never does.
 *
 * oracle/silent_oracle.py is the LITERAL oracle (dense weights, float64 accumulation). This file states the same
 * algorithm in the CANONICAL float32 EVALUATION ORDER (DESIGN.md "Canonical order"), which the CUDA kernels share, so
 * that kernel outputs can be compared BIT-EXACTLY (values and feature-point indices) instead of within a tolerance.
 * It is itself checked against the literal oracle (tests/test_oracle_c.py) to the north-star tolerance.
 * PARITY UNPINNED against real TensorFlow 1.x for the image-side operators (see silent_oracle.py header).
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math oracle/silent_oracle.c -lm   (oracle/build.py)
 *
 * Canonical order, summarised (references are to /root/reference/slam_recognition):
 *  - convolution (util/apply_filter.py:4-7, TF-1 conv2d SAME, cross-correlation): per output channel one fmaf chain
 *    acc = fmaf(W[ky][kx][ci][co], x[y+ky-p][x+kx-p][ci], acc) from +0 in (ky, ci, kx) order, out-of-bounds x = 0.
 *    If every input-channel slice of W is bitwise identical ("uniform-in", true for rgb_2d_stripe_tensors and
 *    blur_tensor) the chain runs over the channel sum s = ((x0 + x1) + x2 ...) with W[ky][kx][0][co] in (ky, kx) order.
 *  - relu: acc < 0 ? 0 : acc (NaN propagates).  clip: v > hi ? hi : v.
 *  - regulator (util/regulator/gaussian_regulator_tensor.py:34-36): m = conv(x, blur); mm = m > 1 ? 1 : m;
 *    gain = value / canon_pow(mm, root) (IEEE float32 divide); out = x * gain.
 *  - canon_pow: exact special cases, else float32(exp2(root * log2(x))) with the double-precision polynomial
 *    log2/exp2 below (only +, *, /, fma on doubles: identical on any IEEE machine, CPU or GPU).
 *  - pyramid (util/zoom/from_image.py:48-64 + scipy zoom order 5): separable, vertical pass first: for each of the 6
 *    source columns t_i = chain over the 6 y-taps (fmaf, from +0), then the chain over the 6 x-taps of t_i; float32 tap
 *    weights supplied by the caller.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------------------------- */
/* canon_pow                                                                                                        */
/* ---------------------------------------------------------------------------------------------------------------- */

static double canon_log2(double x) /* x > 0, finite, normal */
{
    uint64_t bits;
    memcpy(&bits, &x, 8);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    bits = (bits & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double f;
    memcpy(&f, &bits, 8);
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

static double canon_exp2(double y) /* |y| < 1000 */
{
    double n = nearbyint(y);
    double g = y - n;
    double z = g * 0.6931471805599453;
    double s = 1.0 / 87178291200.0;      /* 1/14! */
    s = fma(s, z, 1.0 / 6227020800.0);   /* 1/13! */
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
    if (ni < -1000) ni = -1000;
    if (ni > 1000) ni = 1000;
    uint64_t bits = (uint64_t)(ni + 1023) << 52;
    double scale;
    memcpy(&scale, &bits, 8);
    return s * scale;
}

float so_canon_pow(float xf, float rf)
{
    if (rf == 0.0f) return 1.0f;
    if (xf != xf || rf != rf) return NAN;
    if (xf == 1.0f) return 1.0f;
    if (xf < 0.0f) return NAN;
    if (xf == 0.0f) return rf > 0.0f ? 0.0f : INFINITY;
    if (isinf(xf)) return rf > 0.0f ? INFINITY : 0.0f;
    double y = (double)rf * canon_log2((double)xf);
    if (y > 999.0) return INFINITY;
    if (y < -999.0) return 0.0f;
    return (float)canon_exp2(y);
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* pyramid                                                                                                          */
/* ---------------------------------------------------------------------------------------------------------------- */

/*
 * frames: [B][H][W][FC] uint8 or float32. Tables per level s (L levels): iy/wy [L][h][6], oky [L][h]; ix/wx [L][w][6],
 * okx [L][w]; indices are absolute frame rows/columns. valid[L][2] = (rows, cols) actually written (rest 0).
 * out: [B*L][h][w][C] float32 taking frame channels 0..C-1 (from_image.py:54-64).
 */
void so_pyramid(const void *frames, int is_u8, int B, int H, int W, int FC, int L, int h, int w, int C, const int *iy,
                const float *wy, const uint8_t *oky, const int *ix, const float *wx, const uint8_t *okx,
                const int *valid, float *out)
{
    const uint8_t *f8 = (const uint8_t *)frames;
    const float *f32 = (const float *)frames;
    for (int b = 0; b < B; ++b)
        for (int s = 0; s < L; ++s)
            for (int oy = 0; oy < h; ++oy)
                for (int ox = 0; ox < w; ++ox)
                    for (int c = 0; c < C; ++c) {
                        float *dst = out + ((((size_t)b * L + s) * h + oy) * w + ox) * C + c;
                        if (oy >= valid[2 * s] || ox >= valid[2 * s + 1] || !oky[s * h + oy] || !okx[s * w + ox]) {
                            *dst = 0.0f;
                            continue;
                        }
                        const int *ty = iy + ((size_t)s * h + oy) * 6;
                        const int *tx = ix + ((size_t)s * w + ox) * 6;
                        const float *gy = wy + ((size_t)s * h + oy) * 6;
                        const float *gx = wx + ((size_t)s * w + ox) * 6;
                        /* vertical pass first: t_i = chain over the 6 y-taps of source column tx[i] ... */
                        float t[6];
                        for (int i = 0; i < 6; ++i) {
                            t[i] = 0.0f;
                            for (int j = 0; j < 6; ++j) {
                                size_t off = (((size_t)b * H + ty[j]) * W + tx[i]) * FC + c;
                                float v = is_u8 ? (float)f8[off] : f32[off];
                                t[i] = fmaf(gy[j], v, t[i]);
                            }
                        }
                        /* ... then the chain over the 6 x-taps */
                        float acc = 0.0f;
                        for (int i = 0; i < 6; ++i) acc = fmaf(gx[i], t[i], acc);
                        *dst = acc;
                    }
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* convolution and friends                                                                                          */
/* ---------------------------------------------------------------------------------------------------------------- */

static int bits_equal(float a, float b)
{
    uint32_t x, y;
    memcpy(&x, &a, 4);
    memcpy(&y, &b, 4);
    return x == y;
}

int so_uniform_in(const float *wt, int k, int cin, int cout)
{
    if (cin < 2) return 0;
    for (int t = 0; t < k * k; ++t)
        for (int ci = 1; ci < cin; ++ci)
            for (int co = 0; co < cout; ++co)
                if (!bits_equal(wt[(t * cin + ci) * cout + co], wt[(t * cin) * cout + co])) return 0;
    return 1;
}

/* post: 0 none, 1 relu, 2 relu then clip at clip_hi */
void so_conv2d(const float *x, int N, int h, int w, int cin, const float *wt, int kh, int kw, int cout, int post,
               float clip_hi, float *out)
{
    const int pt = (kh - 1) / 2, pl = (kw - 1) / 2;
    const int uni = (kh == kw) && so_uniform_in(wt, kh, cin, cout);
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < h; ++y)
            for (int xx = 0; xx < w; ++xx)
                for (int co = 0; co < cout; ++co) {
                    float acc = 0.0f;
                    /* chain order: ky outermost, then input channel, then kx (one input row-channel at a time) */
                    for (int ky = 0; ky < kh; ++ky)
                        for (int ci = 0; ci < (uni ? 1 : cin); ++ci)
                            for (int kx = 0; kx < kw; ++kx) {
                                int sy = y + ky - pt, sx = xx + kx - pl;
                                int inside = sy >= 0 && sy < h && sx >= 0 && sx < w;
                                const float *px =
                                    x + (((size_t)n * h + (inside ? sy : 0)) * w + (inside ? sx : 0)) * cin;
                                const float *wk = wt + ((size_t)(ky * kw + kx) * cin) * cout + co;
                                if (uni) {
                                    float s = inside ? px[0] : 0.0f;
                                    for (int c2 = 1; c2 < cin; ++c2) s = s + (inside ? px[c2] : 0.0f);
                                    acc = fmaf(wk[0], s, acc);
                                } else {
                                    acc = fmaf(wk[(size_t)ci * cout], inside ? px[ci] : 0.0f, acc);
                                }
                            }
                    if (post >= 1) acc = acc < 0.0f ? 0.0f : acc;
                    if (post >= 2) acc = acc > clip_hi ? clip_hi : acc;
                    out[(((size_t)n * h + y) * w + xx) * cout + co] = acc;
                }
}

/* util/regulator/gaussian_regulator_tensor.py:34-36. blur: [k][k][c][c]; tmp: scratch of N*h*w*c floats. */
void so_regulate(const float *x, int N, int h, int w, int c, const float *blur, int k, float value, float root,
                 float *tmp, float *out)
{
    so_conv2d(x, N, h, w, c, blur, k, k, c, 0, 0.0f, tmp);
    size_t total = (size_t)N * h * w * c;
    for (size_t i = 0; i < total; ++i) {
        float m = tmp[i];
        float mm = m > 1.0f ? 1.0f : m;
        float gain = value / so_canon_pow(mm, root);
        out[i] = x[i] * gain;
    }
}

/* util/selection/isolate_rectangle.py:19-23: multiply by a 0/1 box; border = 0 * x (NaN stays NaN). */
void so_pad_inwards(const float *x, int N, int h, int w, int c, int top, int bottom, int left, int right, float *out)
{
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < h; ++y)
            for (int xx = 0; xx < w; ++xx) {
                int inside = y >= top && y < h - bottom && xx >= left && xx < w - right;
                for (int ch = 0; ch < c; ++ch) {
                    size_t i = (((size_t)n * h + y) * w + xx) * c + ch;
                    float v = x[i];
                    out[i] = inside ? v : (v != v ? v : 0.0f * v);
                }
            }
}

/* util/color/get_value.py:6-12 */
void so_value_from_color(const float *x, size_t pixels, int c, float *out)
{
    const float div = 1.0f / (float)c;
    for (size_t i = 0; i < pixels; ++i) {
        float s = x[i * c];
        for (int ch = 1; ch < c; ++ch) s = s + x[i * c + ch];
        out[i] = s * div;
    }
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* feature-point emit  (util/selection/top_value_points.py:32-45)                                                   */
/* ---------------------------------------------------------------------------------------------------------------- */

static void same_geometry(int n, int k, int s, int *out, int *before)
{
    *out = (n + s - 1) / s;
    int total = (*out - 1) * s + k - n;
    if (total < 0) total = 0;
    *before = total / 2;
}

static int nearest_src(int dst, int n_in, int n_out)
{
    float scale = (float)((double)n_in / (double)n_out);
    int src = (int)floorf((float)dst * scale);
    return src < n_in - 1 ? src : n_in - 1;
}

/* value: [N][h][w]. Writes rows (n, y, x, 0) in row-major order; returns total count (may exceed capacity). */
int64_t so_max_value_indices_region(const float *value, int N, int h, int w, int region_h, int region_w,
                                    int64_t *out, int64_t capacity)
{
    int oh, ow, pt, pl;
    same_geometry(h, h, region_h, &oh, &pt);
    same_geometry(w, w, region_w, &ow, &pl);
    int64_t count = 0;
    float pooled[64 * 64];
    if (oh > 64 || ow > 64) return -1;
    for (int n = 0; n < N; ++n) {
        const float *v = value + (size_t)n * h * w;
        for (int i = 0; i < oh; ++i)
            for (int j = 0; j < ow; ++j) {
                int ya = i * region_h - pt, yb = ya + h, xa = j * region_w - pl, xb = xa + w;
                if (ya < 0) ya = 0;
                if (xa < 0) xa = 0;
                if (yb > h) yb = h;
                if (xb > w) xb = w;
                float best = -INFINITY;
                int seen_nan = 0;
                for (int y = ya; y < yb; ++y)
                    for (int xx = xa; xx < xb; ++xx) {
                        float t = v[(size_t)y * w + xx];
                        if (t != t) seen_nan = 1;
                        else if (t > best) best = t;
                    }
                pooled[i * ow + j] = seen_nan ? NAN : best;
            }
        for (int y = 0; y < h; ++y)
            for (int xx = 0; xx < w; ++xx) {
                float up = pooled[nearest_src(y, oh, h) * ow + nearest_src(xx, ow, w)];
                if (v[(size_t)y * w + xx] >= up) {
                    if (count < capacity) {
                        out[4 * count + 0] = n;
                        out[4 * count + 1] = y;
                        out[4 * count + 2] = xx;
                        out[4 * count + 3] = 0;
                    }
                    ++count;
                }
            }
    }
    return count;
}

/* util/selection/top_value_points.py:8-29 */
void so_top_value_points(const float *color, const float *value, int N, int h, int w, int c, double top_percent,
                         float *out)
{
    /* TF turns the Python floats (1.0 - top_percent) and top_percent into float32 constants (top_value_points.py:22) */
    const float keep_max = (float)(1.0 - top_percent), keep_min = (float)top_percent;
    for (int n = 0; n < N; ++n) {
        const float *v = value + (size_t)n * h * w;
        float mx = -INFINITY, mneg = -INFINITY;
        int seen_nan = 0;
        for (size_t i = 0; i < (size_t)h * w; ++i) {
            float t = v[i];
            if (t != t) seen_nan = 1;
            else {
                if (t > mx) mx = t;
                if (-t > mneg) mneg = -t;
            }
        }
        float mn = -1.0f * mneg;
        float thr = seen_nan ? NAN : keep_max * mx + keep_min * mn;
        for (size_t i = 0; i < (size_t)h * w; ++i) {
            float keep = v[i] >= thr ? 1.0f : 0.0f;
            for (int ch = 0; ch < c; ++ch) {
                size_t o = ((size_t)n * h * w + i) * c + ch;
                out[o] = color[o] * keep;
            }
        }
    }
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* composition  (recognition_testing.py:69-77)                                                                      */
/* ---------------------------------------------------------------------------------------------------------------- */

/* All buffers [N][h][w][3] except gray [N][h][w]; scratch: 2 * N*h*w*3 floats. */
void so_line_end_stack(const float *pyr, int N, int h, int w, const float *rgc, const float *rgby,
                       const float *stripe, const float *blur, int blur_k, const float *end, float *a, float *b,
                       float *c, float *orient, float *line_end, float *padded, float *gray, float *scratch)
{
    so_conv2d(pyr, N, h, w, 3, rgc, 3, 3, 3, 1, 0.0f, a);
    so_conv2d(a, N, h, w, 3, rgby, 3, 3, 3, 1, 0.0f, b);
    so_conv2d(b, N, h, w, 3, stripe, 3, 3, 3, 1, 0.0f, c);
    so_regulate(c, N, h, w, 3, blur, blur_k, 1.0f, 0.1f, scratch, orient);
    so_conv2d(orient, N, h, w, 3, end, 3, 3, 3, 2, 255.0f, line_end);
    so_pad_inwards(line_end, N, h, w, 3, 2, 2, 2, 2, padded);
    so_value_from_color(padded, (size_t)N * h * w, 3, gray);
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* FUSED canonical order: the composition above with the three structured convolutions evaluated the way the fused   */
/* stack kernels do (pysilent_b200/csrc/stack_fused.cu). The structures are properties of the reference's own weight */
/* generators and are detected BITWISE on the float32 weights; a filter without its structure    */
/* takes the generic (ky, ci, kx) chain of so_conv2d. Same math as the reference graph, fewer roundings.             */
/* ---------------------------------------------------------------------------------------------------------------- */

#define W4(wt, t, ci, co) ((wt)[((t) * 3 + (ci)) * 3 + (co)])

/* midget_rgc: only ci == co slices are non-zero (skipping exact zeros is a no-op for finite inputs) */
int so_depthwise3(const float *wt)
{
    for (int t = 0; t < 9; ++t)
        for (int ci = 0; ci < 3; ++ci)
            for (int co = 0; co < 3; ++co)
                if (ci != co && W4(wt, t, ci, co) != 0.0f) return 0;
    return 1;
}

/* rgby_3: off-centre taps couple every pair of DIFFERENT channels through one kernel S (doubled for the pair 1 <-> 2);
 * the centre tap holds d_i on the diagonal and e between channels 1 and 2 (cc/center_surround/rgby.py:36-56). */
int so_rgby_shared(const float *wt)
{
    if (!bits_equal(W4(wt, 4, 0, 1), 0.0f)) return 0;
    for (int t = 0; t < 9; ++t) {
        float sv = W4(wt, t, 0, 1);
        if (!bits_equal(W4(wt, t, 0, 2), sv) || !bits_equal(W4(wt, t, 1, 0), sv) || !bits_equal(W4(wt, t, 2, 0), sv))
            return 0;
        if (t != 4) {
            if (!bits_equal(W4(wt, t, 1, 2), 2.0f * sv) || !bits_equal(W4(wt, t, 2, 1), 2.0f * sv)) return 0;
            for (int c = 0; c < 3; ++c)
                if (!bits_equal(W4(wt, t, c, c), 0.0f)) return 0;
        }
    }
    return 1;
}

/* T_i = chain over the 8 off-centre taps of S on channel i (tap order ascending); u0 = d0 a0 + (T1 + T2);
 * u1 = d1 a1 + (e21 a2 + (2 T2 + T0)); u2 = d2 a2 + (e12 a1 + (2 T1 + T0)); each step one fmaf; relu. */
static void so_conv_rgby_shared(const float *a, int N, int h, int w, const float *wt, float *out)
{
    const float d0 = W4(wt, 4, 0, 0), d1 = W4(wt, 4, 1, 1), d2 = W4(wt, 4, 2, 2), e12 = W4(wt, 4, 1, 2),
                e21 = W4(wt, 4, 2, 1);
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                float T[3] = {0.0f, 0.0f, 0.0f};
                for (int ci = 0; ci < 3; ++ci)
                    for (int t = 0; t < 9; ++t) {
                        if (t == 4) continue;
                        int sy = y + t / 3 - 1, sx = x + t % 3 - 1;
                        int inside = sy >= 0 && sy < h && sx >= 0 && sx < w;
                        float v = inside ? a[(((size_t)n * h + sy) * w + sx) * 3 + ci] : 0.0f;
                        T[ci] = fmaf(W4(wt, t, 0, 1), v, T[ci]);
                    }
                const float *c = a + (((size_t)n * h + y) * w + x) * 3;
                float u0 = fmaf(d0, c[0], T[1] + T[2]);
                float u1 = fmaf(d1, c[1], fmaf(e21, c[2], fmaf(2.0f, T[2], T[0])));
                float u2 = fmaf(d2, c[2], fmaf(e12, c[1], fmaf(2.0f, T[1], T[0])));
                float *o = out + (((size_t)n * h + y) * w + x) * 3;
                o[0] = u0 < 0.0f ? 0.0f : u0;
                o[1] = u1 < 0.0f ? 0.0f : u1;
                o[2] = u2 < 0.0f ? 0.0f : u2;
            }
}

/* rgb_2d_stripe_tensors: identical over the input channel and K[ky][kx] == K[2-ky][2-kx] */
int so_stripe_sym180(const float *wt)
{
    if (!so_uniform_in(wt, 3, 3, 3)) return 0;
    for (int t = 0; t < 4; ++t)
        for (int co = 0; co < 3; ++co)
            if (!bits_equal(W4(wt, t, 0, co), W4(wt, 8 - t, 0, co))) return 0;
    return 1;
}

/* on the channel sum s = (b0 + b1) + b2: chain over (s[-1,-1] + s[1,1]), (s[-1,0] + s[1,0]), (s[-1,1] + s[1,-1]),
 * (s[0,-1] + s[0,1]), s[0,0] with the weights of taps 0..4; relu. */
static void so_conv_stripe_sym(const float *b, int N, int h, int w, const float *wt, float *out)
{
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                float s[9];
                for (int t = 0; t < 9; ++t) {
                    int sy = y + t / 3 - 1, sx = x + t % 3 - 1;
                    if (sy >= 0 && sy < h && sx >= 0 && sx < w) {
                        const float *px = b + (((size_t)n * h + sy) * w + sx) * 3;
                        s[t] = (px[0] + px[1]) + px[2];
                    } else {
                        s[t] = 0.0f;
                    }
                }
                const float p[5] = {s[0] + s[8], s[1] + s[7], s[2] + s[6], s[3] + s[5], s[4]};
                for (int co = 0; co < 3; ++co) {
                    float acc = 0.0f;
                    for (int t = 0; t < 5; ++t) acc = fmaf(W4(wt, t, 0, co), p[t], acc);
                    out[(((size_t)n * h + y) * w + x) * 3 + co] = acc < 0.0f ? 0.0f : acc;
                }
            }
}

/* rgb_2d_end_tensors: input channel ci reaches the two other output channels through the same kernel */
int so_end_ownoth(const float *wt)
{
    for (int t = 0; t < 9; ++t)
        for (int ci = 0; ci < 3; ++ci)
            if (!bits_equal(W4(wt, t, ci, (ci + 1) % 3), W4(wt, t, ci, (ci + 2) % 3))) return 0;
    return 1;
}

/* per input channel an "own" and an "other" chain over the 9 taps; e_co = (t_0 + t_1) + t_2, t_ci = own or other;
 * relu, clip at clip_hi. */
static void so_conv_end_ownoth(const float *d, int N, int h, int w, const float *wt, float clip_hi, float *out)
{
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                float own[3], oth[3];
                for (int ci = 0; ci < 3; ++ci) {
                    own[ci] = oth[ci] = 0.0f;
                    for (int t = 0; t < 9; ++t) {
                        int sy = y + t / 3 - 1, sx = x + t % 3 - 1;
                        int inside = sy >= 0 && sy < h && sx >= 0 && sx < w;
                        float v = inside ? d[(((size_t)n * h + sy) * w + sx) * 3 + ci] : 0.0f;
                        own[ci] = fmaf(W4(wt, t, ci, ci), v, own[ci]);
                        oth[ci] = fmaf(W4(wt, t, ci, (ci + 1) % 3), v, oth[ci]);
                    }
                }
                for (int co = 0; co < 3; ++co) {
                    float acc = co == 0 ? own[0] : oth[0];
                    acc = acc + (co == 1 ? own[1] : oth[1]);
                    acc = acc + (co == 2 ? own[2] : oth[2]);
                    acc = acc < 0.0f ? 0.0f : acc;
                    acc = acc > clip_hi ? clip_hi : acc;
                    out[(((size_t)n * h + y) * w + x) * 3 + co] = acc;
                }
            }
}

/* S1 + S2 as the fused kernels evaluate them (stack_a_kernel): a = relu(rgc * x), b = relu(rgby * a) with the shared-
 * surround order when both structures hold. Used by the orientation-bank composition (config C4). */
void so_rgc_rgby_fused(const float *pyr, int N, int h, int w, const float *rgc, const float *rgby, float *a, float *b)
{
    so_conv2d(pyr, N, h, w, 3, rgc, 3, 3, 3, 1, 0.0f, a);
    if (so_depthwise3(rgc) && so_rgby_shared(rgby)) so_conv_rgby_shared(a, N, h, w, rgby, b);
    else so_conv2d(a, N, h, w, 3, rgby, 3, 3, 3, 1, 0.0f, b);
}

/* Same buffers as so_line_end_stack. Which stage takes its structured order mirrors the kernels' dispatch:
 * S2 shared needs a depthwise rgc as well (one stack_a variant), S3 symmetric and S5 own/other go together. */
void so_line_end_stack_fused(const float *pyr, int N, int h, int w, const float *rgc, const float *rgby,
                             const float *stripe, const float *blur, int blur_k, const float *end, float *a, float *b,
                             float *c, float *orient, float *line_end, float *padded, float *gray, float *scratch)
{
    so_conv2d(pyr, N, h, w, 3, rgc, 3, 3, 3, 1, 0.0f, a);
    if (so_depthwise3(rgc) && so_rgby_shared(rgby)) so_conv_rgby_shared(a, N, h, w, rgby, b);
    else so_conv2d(a, N, h, w, 3, rgby, 3, 3, 3, 1, 0.0f, b);
    const int structured = so_stripe_sym180(stripe) && so_end_ownoth(end);
    if (structured) so_conv_stripe_sym(b, N, h, w, stripe, c);
    else so_conv2d(b, N, h, w, 3, stripe, 3, 3, 3, 1, 0.0f, c);
    so_regulate(c, N, h, w, 3, blur, blur_k, 1.0f, 0.1f, scratch, orient);
    if (structured) so_conv_end_ownoth(orient, N, h, w, end, 255.0f, line_end);
    else so_conv2d(orient, N, h, w, 3, end, 3, 3, 3, 2, 255.0f, line_end);
    so_pad_inwards(line_end, N, h, w, 3, 2, 2, 2, 2, padded);
    so_value_from_color(padded, (size_t)N * h * w, 3, gray);
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* "next" rows: centroids (util/centroids.py:21-71), nearest resize, boosting (util/energy/boosting.py)  */
/*                                                                                                                  */
/* Canonical order: block sums are float32 add chains from +0 over the window in (ky, kx) order (out-of-image taps    */
/* contribute nothing); biased = index * value is one float32 multiply; centroid = sum / total and distances are      */
/* IEEE float32 divide / subtract / fabsf, summed as (|dx| + |dy|); the boosting update is the float32 sequence       */
/* ((e * 255 - exhaustion) + recovery) / 255 then max(lo) then min(hi), NaN-propagating; pow is canon_pow.            */
/* ---------------------------------------------------------------------------------------------------------------- */

static int nearest_src_f(int dst, int n_in, int n_out)
{
    float scale = (float)((double)n_in / (double)n_out);
    int s = (int)floorf((float)dst * scale);
    return s < n_in - 1 ? s : n_in - 1;
}

/* corrected: [N][oh][ow][2] = (cx, cy); total: [N][oh][ow]; centroids: [N][h][w] (may be NULL) */
void so_get_centroids(const float *value, int N, int h, int w, int rh, int rw, float *corrected, float *total,
                      float *centroids)
{
    int oh, ow, pt, pl;
    same_geometry(h, rh, rh, &oh, &pt);
    same_geometry(w, rw, rw, &ow, &pl);
    for (int n = 0; n < N; ++n) {
        const float *v = value + (size_t)n * h * w;
        for (int i = 0; i < oh; ++i)
            for (int j = 0; j < ow; ++j) {
                float sx = 0.0f, sy = 0.0f, st = 0.0f;
                for (int ky = 0; ky < rh; ++ky) {
                    int y = i * rh - pt + ky;
                    if (y < 0 || y >= h) continue;
                    for (int kx = 0; kx < rw; ++kx) {
                        int x = j * rw - pl + kx;
                        if (x < 0 || x >= w) continue;
                        float val = v[(size_t)y * w + x];
                        sx = sx + (float)x * val;
                        sy = sy + (float)y * val;
                        st = st + val;
                    }
                }
                size_t o = ((size_t)n * oh + i) * ow + j;
                corrected[2 * o] = sx / st;
                corrected[2 * o + 1] = sy / st;
                total[o] = st;
            }
        if (!centroids) continue;
        for (int y = 0; y < h; ++y) {
            int i = nearest_src_f(y, oh, h);
            for (int x = 0; x < w; ++x) {
                int j = nearest_src_f(x, ow, w);
                size_t o = ((size_t)n * oh + i) * ow + j;
                float dx = fabsf(corrected[2 * o] - (float)x), dy = fabsf(corrected[2 * o + 1] - (float)y);
                centroids[((size_t)n * h + y) * w + x] = dx + dy;
            }
        }
    }
}

void so_resize_nearest(const float *x, int N, int h, int w, int c, int oh, int ow, float *out)
{
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < oh; ++y) {
            int sy = nearest_src_f(y, h, oh);
            for (int xx = 0; xx < ow; ++xx) {
                int sx = nearest_src_f(xx, w, ow);
                for (int ch = 0; ch < c; ++ch)
                    out[(((size_t)n * oh + y) * ow + xx) * c + ch] = x[(((size_t)n * h + sy) * w + sx) * c + ch];
            }
        }
}

static float nan_max(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

/* recovery_mode: 1 constant (10), 2 input based (0.8 * fire strength), 3 both (max). energy is updated in place. */
void so_get_boosting(const float *inp, float *energy, int N, int h, int w, float exhaustion_max, float excitation_max,
                     int recovery_mode, float *fired, float *biased_scratch)
{
    size_t plane = (size_t)h * w;
    for (size_t i = 0; i < (size_t)N * plane; ++i) biased_scratch[i] = so_canon_pow(inp[i], energy[i]);
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const float *b = biased_scratch + (size_t)n * plane;
                size_t o = (size_t)n * plane + (size_t)y * w + x;
                int first = 1;
                float m = 0.0f;
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        int yy = y + dy, xx = x + dx;
                        if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                        float v = b[(size_t)yy * w + xx];
                        m = first ? v : nan_max(m, v);
                        first = 0;
                    }
                float f = (b[(size_t)y * w + x] == m) ? 1.0f : 0.0f;
                float strength = f * inp[o];
                float exhaustion = f * 255.0f;
                float recovery = 10.0f;
                if (recovery_mode == 2) recovery = strength * 0.8f;
                if (recovery_mode == 3) recovery = nan_max(strength * 0.8f, 10.0f);
                float t = energy[o] * 255.0f;
                t = t - exhaustion;
                t = t + recovery;
                t = t / 255.0f;
                t = t < -exhaustion_max ? -exhaustion_max : t;
                t = t > excitation_max ? excitation_max : t;
                fired[o] = f;
                energy[o] = t;
            }
}

/* display arithmetic of recognition_testing.py:79-100. kind: 0 x / 255; 1 255 - x * 255; 2 clip(x * 63.75, 1, 256) - 1;
 * 3 x * 255; 4 x * 127.5 + 127.5 (two roundings); 5 a * b (fired * importance) with b = second input */
void so_pointwise(const float *x, const float *y, size_t count, int kind, float *out)
{
    for (size_t i = 0; i < count; ++i) {
        float v = x[i], r;
        switch (kind) {
            case 0: r = v / 255.0f; break;
            case 1: { float t = v * 255.0f; r = 255.0f - t; } break;
            case 2: { float t = v * 63.75f; t = t < 1.0f ? 1.0f : t; t = t > 256.0f ? 256.0f : t; r = t - 1.0f; } break;
            case 3: r = v * 255.0f; break;
            case 4: { float t = v * 127.5f; r = t + 127.5f; } break;
            default: r = v * y[i]; break;
        }
        out[i] = r;
    }
}
