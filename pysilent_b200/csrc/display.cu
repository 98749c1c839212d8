// The display tensors that follow gray_line_end_tensor in LineEndDisplayer.compile (reference
// recognition_testing.py:79-100): block centroids (util/centroids.py:21-71), nearest-neighbour resize, the stateful
// boosting non-max (util/energy/boosting.py:10-42, recovery.py:4-22) and the scalar display arithmetic.
//
// All of them are tiny against the filter stack (a 3 x 3 block pool of a one-channel level, 64 x 96 per level for the
// boosting state): one thread per output element, coalesced one-channel reads, no shared memory. Arithmetic follows the
// canonical order of oracle/silent_oracle.c ("next rows" section), so results are bit-identical to it.
#include "common.cuh"

namespace silent {

__host__ __device__ inline void same_geometry(int n, int k, int s, int *out, int *before)
{
    *out = (n + s - 1) / s;
    int total = (*out - 1) * s + k - n;
    if (total < 0) total = 0;
    *before = total / 2;
}

// tf.image.resize_nearest_neighbor, align_corners = False: min(floor(dst * float32(in / out)), in - 1)
__device__ __forceinline__ int nearest_src(int dst, int n_in, float scale)
{
    const int s = (int)floorf(__fmul_rn((float)dst, scale));
    return s < n_in - 1 ? s : n_in - 1;
}
static float nearest_scale(int n_in, int n_out) { return (float)((double)n_in / (double)n_out); }

// one thread per (image, block): value-weighted index sums over the region window, centroid = sum / total
__global__ void __launch_bounds__(256) centroid_blocks_kernel(const float *__restrict__ value, int n, int h, int w, int rh,
                                                              int rw, int oh, int ow, int pt, int pl,
                                                              float *__restrict__ corrected, float *__restrict__ total)
{
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= (size_t)n * oh * ow) return;
    const int j = (int)(o % ow), i = (int)((o / ow) % oh), img = (int)(o / ((size_t)ow * oh));
    const float *v = value + (size_t)img * h * w;
    float sx = 0.0f, sy = 0.0f, st = 0.0f;
    for (int ky = 0; ky < rh; ++ky) {
        const int y = i * rh - pt + ky;
        if (y < 0 || y >= h) continue;
        for (int kx = 0; kx < rw; ++kx) {
            const int x = j * rw - pl + kx;
            if (x < 0 || x >= w) continue;
            const float val = __ldg(v + (size_t)y * w + x);
            sx = __fadd_rn(sx, __fmul_rn((float)x, val));
            sy = __fadd_rn(sy, __fmul_rn((float)y, val));
            st = __fadd_rn(st, val);
        }
    }
    corrected[2 * o] = __fdiv_rn(sx, st);       // 0 / 0 = NaN on an empty block, like the reference graph
    corrected[2 * o + 1] = __fdiv_rn(sy, st);
    total[o] = st;
}

// one thread per pixel: L1 distance to the centroid of its (nearest-upsampled) block
__global__ void __launch_bounds__(256) centroid_distance_kernel(const float *__restrict__ corrected, int n, int h, int w,
                                                                int oh, int ow, float sy_scale, float sx_scale,
                                                                float *__restrict__ out)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (size_t)n * h * w) return;
    const int x = (int)(p % w), y = (int)((p / w) % h), img = (int)(p / ((size_t)w * h));
    const int i = nearest_src(y, oh, sy_scale), j = nearest_src(x, ow, sx_scale);
    const float2 c = __ldg(reinterpret_cast<const float2 *>(corrected) + ((size_t)img * oh + i) * ow + j);
    out[p] = __fadd_rn(fabsf(__fsub_rn(c.x, (float)x)), fabsf(__fsub_rn(c.y, (float)y)));
}

__global__ void __launch_bounds__(256) resize_nearest_kernel(const float *__restrict__ x, int n, int h, int w, int c, int oh,
                                                             int ow, float sy_scale, float sx_scale,
                                                             float *__restrict__ out)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)n * oh * ow * c) return;
    const int ch = (int)(e % c);
    const size_t p = e / c;
    const int xx = (int)(p % ow), yy = (int)((p / ow) % oh), img = (int)(p / ((size_t)ow * oh));
    const int sy = nearest_src(yy, h, sy_scale), sx = nearest_src(xx, w, sx_scale);
    out[e] = __ldg(x + (((size_t)img * h + sy) * w + sx) * c + ch);
}

__global__ void __launch_bounds__(256) boosting_bias_kernel(const float *__restrict__ inp, const float *__restrict__ energy,
                                                            size_t count, float *__restrict__ biased)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) biased[i] = canon_pow(inp[i], energy[i]);   // input ** exhaustion_tensor.value()  boosting.py:17
}

__device__ __forceinline__ float nan_max(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

// 3 x 3 max-pool equality test (the reference's non-max), fire strength, exhaustion / recovery, state update
__global__ void __launch_bounds__(256) boosting_update_kernel(const float *__restrict__ inp, const float *__restrict__ biased,
                                                              float *__restrict__ energy, int n, int h, int w,
                                                              float exhaustion_max, float excitation_max, int recovery_mode,
                                                              float *__restrict__ fired)
{
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= (size_t)n * h * w) return;
    const int x = (int)(o % w), y = (int)((o / w) % h);
    const float *b = biased + (o - (size_t)y * w - x);
    bool first = true;
    float m = 0.0f;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
            const float v = __ldg(b + (size_t)yy * w + xx);
            m = first ? v : nan_max(m, v);
            first = false;
        }
    const float f = (__ldg(b + (size_t)y * w + x) == m) ? 1.0f : 0.0f;                 // boosting.py:20-22
    const float strength = __fmul_rn(f, inp[o]);                                       // :24
    const float exhaustion = __fmul_rn(f, 255.0f);                                     // :26
    float recovery = 10.0f;                                                            // recovery.py:4-5
    if (recovery_mode == 2) recovery = __fmul_rn(strength, 0.8f);                      // recovery.py:8-9
    if (recovery_mode == 3) recovery = nan_max(__fmul_rn(strength, 0.8f), 10.0f);      // recovery.py:19
    float t = __fmul_rn(energy[o], 255.0f);                                            // boosting.py:30-33
    t = __fsub_rn(t, exhaustion);
    t = __fadd_rn(t, recovery);
    t = __fdiv_rn(t, 255.0f);
    t = t < -exhaustion_max ? -exhaustion_max : t;
    t = t > excitation_max ? excitation_max : t;
    fired[o] = f;
    energy[o] = t;
}

__global__ void __launch_bounds__(256) pointwise_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                        size_t count, int kind, float *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float v = x[i];
    float r;
    switch (kind) {
        case SILENT_PW_DIV255: r = __fdiv_rn(v, 255.0f); break;
        case SILENT_PW_INVERT255: r = __fsub_rn(255.0f, __fmul_rn(v, 255.0f)); break;
        case SILENT_PW_IMPORTANCE: {
            float t = __fmul_rn(v, 63.75f);
            t = t < 1.0f ? 1.0f : t;
            t = t > 256.0f ? 256.0f : t;
            r = __fsub_rn(t, 1.0f);
        } break;
        case SILENT_PW_MUL255: r = __fmul_rn(v, 255.0f); break;
        case SILENT_PW_ENERGY_DISPLAY: r = __fadd_rn(__fmul_rn(v, 127.5f), 127.5f); break;
        default: r = __fmul_rn(v, y[i]); break;
    }
    out[i] = r;
}

// rows [0, capacity) = the rank's points with the level id rebased to the global frame order (rows beyond the count are
// zeroed), row `capacity` = (count, 0, 0, 0): ONE buffer, so the exchange is a single all-gather
__global__ void __launch_bounds__(256) pack_points_kernel(const long long *__restrict__ points,
                                                          const long long *__restrict__ count, long long capacity,
                                                          long long level_offset, long long *__restrict__ packed)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > capacity) return;
    const long long n = *count < capacity ? *count : capacity;
    longlong2 a = make_longlong2(0, 0), b = make_longlong2(0, 0);
    if (i == capacity) {
        a.x = *count;
    } else if (i < n) {
        a = reinterpret_cast<const longlong2 *>(points)[2 * i];
        b = reinterpret_cast<const longlong2 *>(points)[2 * i + 1];
        a.x += level_offset;
    }
    reinterpret_cast<longlong2 *>(packed)[2 * i] = a;
    reinterpret_cast<longlong2 *>(packed)[2 * i + 1] = b;
}

// ---------------------------------------------------------------------------------------------------------------------
// The whole display chain of LineEndDisplayer.compile (reference recognition_testing.py:79-100) in TWO launches instead of
// seventeen. The operators above stay as the stand-alone API; these kernels apply the same roundings in the same order,
// so their results are bit-identical to the chain of operators (tests/test_gpu_parity.py compares both).
// ---------------------------------------------------------------------------------------------------------------------
struct DisplayGeom {
    int n, h, w;                 // gray levels
    int rh, rw;                  // centroid region
    int oh, ow, pt, pl;          // block grid of the full image (SAME geometry) ...
    float by, bx;                // ... and the nearest-neighbour scales that map a pixel to its block
    int h2, w2;                  // the e^-0.5 resize of gray (recognition_testing.py:82-83)
    float ry, rx;                // nearest-neighbour scales gray -> im2
    int oh2, ow2, pt2, pl2;      // block grid of im2
    float by2, bx2;
};

// value-weighted index sums of block (i, j); HALF reads gray through the nearest-neighbour map of im2. value = gray / 255.
template <bool HALF>
__device__ __forceinline__ void block_sums(const float *__restrict__ gray, const DisplayGeom &G, int i, int j, float &sx,
                                           float &sy, float &st)
{
    const int hh = HALF ? G.h2 : G.h, ww = HALF ? G.w2 : G.w, pt = HALF ? G.pt2 : G.pt, pl = HALF ? G.pl2 : G.pl;
    sx = sy = st = 0.0f;
    for (int ky = 0; ky < G.rh; ++ky) {
        const int y = i * G.rh - pt + ky;
        if (y < 0 || y >= hh) continue;
        const int gy = HALF ? nearest_src(y, G.h, G.ry) : y;
        for (int kx = 0; kx < G.rw; ++kx) {
            const int x = j * G.rw - pl + kx;
            if (x < 0 || x >= ww) continue;
            const int gx = HALF ? nearest_src(x, G.w, G.rx) : x;
            const float val = __fdiv_rn(__ldg(gray + (size_t)gy * G.w + gx), 255.0f);
            sx = __fadd_rn(sx, __fmul_rn((float)x, val));
            sy = __fadd_rn(sy, __fmul_rn((float)y, val));
            st = __fadd_rn(st, val);
        }
    }
}

// threads [0, n*h*w): 255 - centroids * 255; then [.., + n*h2*w2): the same on im2; then [.., + n*oh*ow): block importance
// and importance ** energy for the non-max of the second launch. A pixel recomputes the sums of its own block (9 cached
// loads) instead of waiting for a block pass: no dependency between threads, one launch.
__global__ void __launch_bounds__(256) display_centroids_kernel(const float *__restrict__ gray, const DisplayGeom G,
                                                                const float *__restrict__ energy,
                                                                float *__restrict__ cent_disp, float *__restrict__ cent2_disp,
                                                                float *__restrict__ importance, float *__restrict__ biased)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t full = (size_t)G.n * G.h * G.w, half = (size_t)G.n * G.h2 * G.w2, blocks = (size_t)G.n * G.oh * G.ow;
    float sx, sy, st;
    if (t < full) {
        const int x = (int)(t % G.w), y = (int)((t / G.w) % G.h), img = (int)(t / ((size_t)G.w * G.h));
        block_sums<false>(gray + (size_t)img * G.h * G.w, G, nearest_src(y, G.oh, G.by), nearest_src(x, G.ow, G.bx), sx, sy, st);
        const float d = __fadd_rn(fabsf(__fsub_rn(__fdiv_rn(sx, st), (float)x)), fabsf(__fsub_rn(__fdiv_rn(sy, st), (float)y)));
        cent_disp[t] = __fsub_rn(255.0f, __fmul_rn(d, 255.0f));
        return;
    }
    t -= full;
    if (t < half) {
        const int x = (int)(t % G.w2), y = (int)((t / G.w2) % G.h2), img = (int)(t / ((size_t)G.w2 * G.h2));
        block_sums<true>(gray + (size_t)img * G.h * G.w, G, nearest_src(y, G.oh2, G.by2), nearest_src(x, G.ow2, G.bx2), sx, sy, st);
        const float d = __fadd_rn(fabsf(__fsub_rn(__fdiv_rn(sx, st), (float)x)), fabsf(__fsub_rn(__fdiv_rn(sy, st), (float)y)));
        cent2_disp[t] = __fsub_rn(255.0f, __fmul_rn(d, 255.0f));
        return;
    }
    t -= half;
    if (t < blocks) {
        const int j = (int)(t % G.ow), i = (int)((t / G.ow) % G.oh), img = (int)(t / ((size_t)G.ow * G.oh));
        block_sums<false>(gray + (size_t)img * G.h * G.w, G, i, j, sx, sy, st);
        float v = __fmul_rn(st, 63.75f);                       // clip(importances * (255 / 4), 1, 256) - 1  (:81)
        v = v < 1.0f ? 1.0f : v;
        v = v > 256.0f ? 256.0f : v;
        v = __fsub_rn(v, 1.0f);
        importance[t] = v;
        biased[t] = canon_pow(v, energy[t]);                   // boosting.py:17
    }
}

// the non-max / exhaustion / recovery update of boosting_update_kernel plus the two display tensors made from it
// (boosting.py:35-40 with for_visualizing, recognition_testing.py:98-100): three equal channels each
__global__ void __launch_bounds__(256) display_boosting_kernel(const float *__restrict__ inp, const float *__restrict__ biased,
                                                               float *__restrict__ energy, int n, int h, int w,
                                                               float exhaustion_max, float excitation_max, int recovery_mode,
                                                               float normer, float centerer, float *__restrict__ fired_disp,
                                                               float *__restrict__ update_disp)
{
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= (size_t)n * h * w) return;
    const int x = (int)(o % w), y = (int)((o / w) % h);
    const float *b = biased + (o - (size_t)y * w - x);
    bool first = true;
    float m = 0.0f;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
            const float v = __ldg(b + (size_t)yy * w + xx);
            m = first ? v : nan_max(m, v);
            first = false;
        }
    const float f = (__ldg(b + (size_t)y * w + x) == m) ? 1.0f : 0.0f;
    const float strength = __fmul_rn(f, inp[o]);
    const float exhaustion = __fmul_rn(f, 255.0f);
    float recovery = 10.0f;
    if (recovery_mode == 2) recovery = __fmul_rn(strength, 0.8f);
    if (recovery_mode == 3) recovery = nan_max(__fmul_rn(strength, 0.8f), 10.0f);
    float t = __fmul_rn(energy[o], 255.0f);
    t = __fsub_rn(t, exhaustion);
    t = __fadd_rn(t, recovery);
    t = __fdiv_rn(t, 255.0f);
    t = t < -exhaustion_max ? -exhaustion_max : t;
    t = t > excitation_max ? excitation_max : t;
    energy[o] = t;
    const float fd = __fmul_rn(strength, 255.0f);                        // (has_fired * input) * 255
    const float ud = __fadd_rn(__fmul_rn(t, normer), centerer);          // update_energy * normer + centerer
    fired_disp[3 * o] = fired_disp[3 * o + 1] = fired_disp[3 * o + 2] = fd;
    update_disp[3 * o] = update_disp[3 * o + 1] = update_disp[3 * o + 2] = ud;
}

static unsigned blocks_for(size_t count) { return (unsigned)((count + 255) / 256); }

}  // namespace silent

using namespace silent;

extern "C" {

int silent_get_centroids(const float *value_dev, int n, int h, int w, int region_h, int region_w, float *corrected_dev,
                         float *total_dev, float *centroids_dev, silent_stream stream)
{
    if (!value_dev || !corrected_dev || !total_dev) return fail(SILENT_E_INVAL, "silent_get_centroids: null argument");
    if (n <= 0 || h <= 0 || w <= 0 || region_h <= 0 || region_w <= 0)
        return fail(SILENT_E_INVAL, "silent_get_centroids: bad shape %dx%dx%d region %dx%d", n, h, w, region_h, region_w);
    int oh, ow, pt, pl;
    same_geometry(h, region_h, region_h, &oh, &pt);
    same_geometry(w, region_w, region_w, &ow, &pl);
    cudaStream_t s = (cudaStream_t)stream;
    centroid_blocks_kernel<<<blocks_for((size_t)n * oh * ow), 256, 0, s>>>(value_dev, n, h, w, region_h, region_w, oh, ow,
                                                                           pt, pl, corrected_dev, total_dev);
    SILENT_LAUNCH_CHECK("centroid_blocks_kernel");
    if (centroids_dev) {
        centroid_distance_kernel<<<blocks_for((size_t)n * h * w), 256, 0, s>>>(corrected_dev, n, h, w, oh, ow,
                                                                              nearest_scale(oh, h), nearest_scale(ow, w),
                                                                              centroids_dev);
        SILENT_LAUNCH_CHECK("centroid_distance_kernel");
    }
    return SILENT_OK;
}

int silent_resize_nearest(const float *x_dev, int n, int h, int w, int c, int out_h, int out_w, float *out_dev,
                          silent_stream stream)
{
    if (!x_dev || !out_dev) return fail(SILENT_E_INVAL, "silent_resize_nearest: null argument");
    if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || out_h <= 0 || out_w <= 0)
        return fail(SILENT_E_INVAL, "silent_resize_nearest: bad shape");
    resize_nearest_kernel<<<blocks_for((size_t)n * out_h * out_w * c), 256, 0, (cudaStream_t)stream>>>(
        x_dev, n, h, w, c, out_h, out_w, nearest_scale(h, out_h), nearest_scale(w, out_w), out_dev);
    SILENT_LAUNCH_CHECK("resize_nearest_kernel");
    return SILENT_OK;
}

int silent_get_boosting(const float *input_dev, float *energy_dev, int n, int h, int w, float exhaustion_max,
                        float excitation_max, int recovery_mode, float *fired_dev, float *scratch_dev,
                        silent_stream stream)
{
    if (!input_dev || !energy_dev || !fired_dev || !scratch_dev)
        return fail(SILENT_E_INVAL, "silent_get_boosting: null argument");
    if (n <= 0 || h <= 0 || w <= 0) return fail(SILENT_E_INVAL, "silent_get_boosting: bad shape %dx%dx%d", n, h, w);
    if (recovery_mode < 1 || recovery_mode > 3)
        return fail(SILENT_E_INVAL, "You must choose a type of recovery");   /* recovery.py:21 */
    const size_t count = (size_t)n * h * w;
    cudaStream_t s = (cudaStream_t)stream;
    boosting_bias_kernel<<<blocks_for(count), 256, 0, s>>>(input_dev, energy_dev, count, scratch_dev);
    SILENT_LAUNCH_CHECK("boosting_bias_kernel");
    boosting_update_kernel<<<blocks_for(count), 256, 0, s>>>(input_dev, scratch_dev, energy_dev, n, h, w, exhaustion_max,
                                                            excitation_max, recovery_mode, fired_dev);
    SILENT_LAUNCH_CHECK("boosting_update_kernel");
    return SILENT_OK;
}

int silent_display_tensors(const float *gray_dev, int n, int h, int w, int region_h, int region_w, int half_h, int half_w,
                           float *energy_dev, float exhaustion_max, float excitation_max, int recovery_mode, float normer,
                           float centerer, float *centroids_disp_dev, float *centroids2_disp_dev, float *fired_disp_dev,
                           float *update_disp_dev, float *scratch_dev, silent_stream stream)
{
    if (!gray_dev || !energy_dev || !centroids_disp_dev || !centroids2_disp_dev || !fired_disp_dev || !update_disp_dev ||
        !scratch_dev)
        return fail(SILENT_E_INVAL, "silent_display_tensors: null argument");
    if (n <= 0 || h <= 0 || w <= 0 || region_h <= 0 || region_w <= 0 || half_h <= 0 || half_w <= 0)
        return fail(SILENT_E_INVAL, "silent_display_tensors: bad shape %dx%dx%d region %dx%d half %dx%d", n, h, w, region_h,
                    region_w, half_h, half_w);
    if (recovery_mode < 1 || recovery_mode > 3)
        return fail(SILENT_E_INVAL, "You must choose a type of recovery");   /* recovery.py:21 */
    DisplayGeom G;
    G.n = n, G.h = h, G.w = w, G.rh = region_h, G.rw = region_w, G.h2 = half_h, G.w2 = half_w;
    same_geometry(h, region_h, region_h, &G.oh, &G.pt);
    same_geometry(w, region_w, region_w, &G.ow, &G.pl);
    same_geometry(half_h, region_h, region_h, &G.oh2, &G.pt2);
    same_geometry(half_w, region_w, region_w, &G.ow2, &G.pl2);
    G.by = nearest_scale(G.oh, h), G.bx = nearest_scale(G.ow, w);
    G.by2 = nearest_scale(G.oh2, half_h), G.bx2 = nearest_scale(G.ow2, half_w);
    G.ry = nearest_scale(h, half_h), G.rx = nearest_scale(w, half_w);
    const size_t blocks = (size_t)n * G.oh * G.ow;
    float *importance = scratch_dev, *biased = scratch_dev + blocks;
    cudaStream_t s = (cudaStream_t)stream;
    display_centroids_kernel<<<blocks_for((size_t)n * h * w + (size_t)n * half_h * half_w + blocks), 256, 0, s>>>(
        gray_dev, G, energy_dev, centroids_disp_dev, centroids2_disp_dev, importance, biased);
    SILENT_LAUNCH_CHECK("display_centroids_kernel");
    display_boosting_kernel<<<blocks_for(blocks), 256, 0, s>>>(importance, biased, energy_dev, n, G.oh, G.ow, exhaustion_max,
                                                               excitation_max, recovery_mode, normer, centerer,
                                                               fired_disp_dev, update_disp_dev);
    SILENT_LAUNCH_CHECK("display_boosting_kernel");
    return SILENT_OK;
}

int silent_pack_points(const int64_t *points_dev, const int64_t *count_dev, int64_t capacity, int64_t level_offset,
                       int64_t *packed_dev, silent_stream stream)
{
    if (!points_dev || !count_dev || !packed_dev) return fail(SILENT_E_INVAL, "silent_pack_points: null argument");
    if (capacity <= 0) return fail(SILENT_E_INVAL, "silent_pack_points: capacity must be positive");
    pack_points_kernel<<<blocks_for((size_t)capacity + 1), 256, 0, (cudaStream_t)stream>>>(
        (const long long *)points_dev, (const long long *)count_dev, (long long)capacity, (long long)level_offset,
        (long long *)packed_dev);
    SILENT_LAUNCH_CHECK("pack_points_kernel");
    return SILENT_OK;
}

int silent_pointwise(const float *x_dev, const float *y_dev, size_t count, int kind, float *out_dev, silent_stream stream)
{
    if (!x_dev || !out_dev) return fail(SILENT_E_INVAL, "silent_pointwise: null argument");
    if (kind < SILENT_PW_DIV255 || kind > SILENT_PW_PRODUCT) return fail(SILENT_E_INVAL, "silent_pointwise: bad kind %d", kind);
    if (kind == SILENT_PW_PRODUCT && !y_dev) return fail(SILENT_E_INVAL, "silent_pointwise: product needs two inputs");
    if (count == 0) return SILENT_OK;
    pointwise_kernel<<<blocks_for(count), 256, 0, (cudaStream_t)stream>>>(x_dev, y_dev, count, kind, out_dev);
    SILENT_LAUNCH_CHECK("pointwise_kernel");
    return SILENT_OK;
}

}  // extern "C"
