"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of pySILEnT's filter-pipeline hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this
module; the product package ``pysilent_b200`` never does (its ops fail loudly without the CUDA library).

This is the LITERAL oracle: every operator is evaluated the way the reference's graph states it (dense convolutions with
the full ``[k,k,Cin,Cout]`` weights, float64 accumulation rounded once to float32 per operator, i.e. the correctly
rounded value of each TF-1 float32 op). ``oracle/silent_oracle.c`` is the second, bit-defined float32 restatement whose
evaluation order the CUDA kernels share.

Pinning status
  * weight generators: pinned by the reference's own tests and by outputs of the reference's generators executed in
    the build container (``tests/golden/generators.npz``);
  * pyramid (``from_image``): pinned against the reference's ``image_to_zoom_tensor`` executed in the build container
    on scipy 1.18 (``tests/golden/pyramid.npz``) and, where scipy is importable, against ``scipy.ndimage.zoom`` live;
  * conv / regulate / mask / value / emit: pinned against the reference's own ``filters/*.py`` + ``util/**`` executed
    on ``oracle/tf1_shim.py`` (``tests/golden/stack.npz``). Real TensorFlow 1.x is not installable, and no reference
    test pins results at that boundary: PARITY UNPINNED against TensorFlow itself.

Every function cites the reference lines it follows (paths relative to ``/root/reference/slam_recognition``).
"""
import math

import numpy as np

# ----------------------------------------------------------------------------------------------------------------------
# P1  pyramid  (util/zoom/from_image.py:10-69  +  scipy.ndimage.zoom(order=5, prefilter=False, mode='constant'))
# ----------------------------------------------------------------------------------------------------------------------


def bspline5(x):
    """Quintic cardinal B-spline."""
    x = np.abs(np.asarray(x, dtype=np.float64))
    x2 = x * x
    out = np.zeros_like(x)
    m = x < 1
    out = np.where(m, (66 - 60 * x2 + 30 * x2 * x2 - 10 * x2 * x2 * x) / 120.0, out)
    m2 = (x >= 1) & (x < 2)
    out = np.where(m2, (51 + 75 * x - 210 * x2 + 150 * x2 * x - 45 * x2 * x2 + 5 * x2 * x2 * x) / 120.0, out)
    m3 = (x >= 2) & (x < 3)
    t = 3 - x
    out = np.where(m3, t * t * t * t * t / 120.0, out)
    return out


def zoom_axis_table(n_in, n_out):
    """Per output index along one axis: 6 source indices, 6 float64 weights and an in-range flag.

    scipy ``ndimage.zoom`` (grid_mode=False): source coordinate ``cc = o * ((n_in - 1) / (n_out - 1))`` (factor 1 when
    ``n_out == 1``); taps ``floor(cc) - 2 .. floor(cc) + 3`` weighted by the quintic B-spline; taps outside
    ``[0, n_in)`` are MIRROR-reflected about the edge sample; a ``cc`` outside ``[0, n_in - 1]`` yields cval = 0.
    """
    idx = np.zeros((n_out, 6), dtype=np.int64)
    wts = np.zeros((n_out, 6), dtype=np.float64)
    ok = np.ones(n_out, dtype=bool)
    factor = (n_in - 1) / (n_out - 1) if n_out > 1 else 1.0
    for o in range(n_out):
        cc = float(o) * factor
        if cc < 0 or cc > n_in - 1:
            ok[o] = False
            continue
        base = int(math.floor(cc)) - 2
        for t in range(6):
            src = base + t
            wts[o, t] = float(bspline5(cc - src))
            if n_in == 1:
                src = 0
            else:
                period = 2 * n_in - 2
                src = src % period
                if src < 0:
                    src += period
                if src >= n_in:
                    src = period - src
            idx[o, t] = src
    return idx, wts, ok


def zoom_order5(plane, factor):
    """``scipy.ndimage.zoom(plane, factor, order=5, prefilter=False)`` for a 2-D float32 array (float64 accumulate)."""
    h, w = plane.shape
    oh, ow = int(round(h * factor)), int(round(w * factor))
    iy, wy, oky = zoom_axis_table(h, oh)
    ix, wx, okx = zoom_axis_table(w, ow)
    src = plane.astype(np.float64)
    rows = np.zeros((oh, w), dtype=np.float64)
    for t in range(6):
        rows += wy[:, t:t + 1] * src[iy[:, t], :]
    out = np.zeros((oh, ow), dtype=np.float64)
    for t in range(6):
        out += rows[:, ix[:, t]] * wx[np.newaxis, :, t]
    out[~oky, :] = 0.0
    out[:, ~okx] = 0.0
    return out.astype(plane.dtype)


def pyramid_levels(image_hw, center_wh, scale):
    """``num_scales`` of from_image.py:45-46."""
    dims = list(reversed(list(center_wh)))
    return int(math.ceil(max(math.log(i / c, scale) for i, c in zip(image_hw, dims))))


def level_crop(image_hw, center_wh, scale, s):
    """Crop bounds of level ``s`` (from_image.py:49-51), clamped the way Python slicing clamps."""
    dims = list(reversed(list(center_wh)))
    bounds = []
    for i, c in zip(image_hw, [c * (scale ** s) for c in dims]):
        lo, hi = int(max((i - c) / 2, 0)), int((i + c) / 2)
        bounds.append((min(lo, i), min(max(hi, 0), i)))
    return bounds


def from_image(image, num_colors, center_dimensions, scale):
    """Foveated pyramid ``[L, h, w, num_colors]`` float32 (from_image.py:10-69).

    Deviation (documented): rows/columns the reference leaves uninitialised (``np.empty``, ``:53``) are 0 here.
    ``center_dimensions`` is ``(w, h)`` and reversed internally (``:44``).
    """
    assert scale > 1, "Scale must be greater than one."
    assert num_colors > 0, "Number of colors must be greater than zero."
    for d in center_dimensions:
        assert d > 0, "Each dimension must be larger than zero."
    image = np.asarray(image, dtype=np.float32)
    hw = image.shape[:-1]
    h, w = list(reversed(list(center_dimensions)))
    levels = pyramid_levels(hw, center_dimensions, scale)
    out = np.zeros((max(levels, 0), h, w, num_colors), dtype=np.float32)
    for s in range(levels):
        (y0, y1), (x0, x1) = level_crop(hw, center_dimensions, scale, s)
        crop = image[y0:y1, x0:x1]
        for c in range(num_colors):
            z = zoom_order5(crop[:, :, c], 1.0 / (scale ** s))
            ym, xm = min(h, z.shape[0]), min(w, z.shape[1])
            out[s, :ym, :xm, c] = z[:ym, :xm]
    return out


# ----------------------------------------------------------------------------------------------------------------------
# F1-F9  image-side operators (TF-1 semantics restated)
# ----------------------------------------------------------------------------------------------------------------------


def conv2d_same(x, w):
    """``tf.nn.conv2d(x, w, [1,1,1,1], 'SAME')``: cross-correlation, zero pad ``(k-1)//2`` before (util/apply_filter.py:4-7)."""
    x = np.asarray(x, dtype=np.float32)
    w = np.asarray(w).astype(np.float32).astype(np.float64)
    n, h, wd, cin = x.shape
    kh, kw, wcin, cout = w.shape
    assert cin == wcin
    pt, pl = (kh - 1) // 2, (kw - 1) // 2
    xp = np.zeros((n, h + kh - 1, wd + kw - 1, cin), dtype=np.float64)
    xp[:, pt:pt + h, pl:pl + wd] = x
    acc = np.zeros((n, h, wd, cout), dtype=np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        for ky in range(kh):
            for kx in range(kw):
                acc += xp[:, ky:ky + h, kx:kx + wd, :] @ w[ky, kx]
    return acc.astype(np.float32)


def relu(x):
    """``tf.maximum(x, [0])`` (NaN-propagating)."""
    return np.maximum(x, np.float32(0))


def conv_relu(x, w):
    """filters/rgc.py:13-16, filters/rgby.py:11-12, filters/orientation.py:24-29."""
    return relu(conv2d_same(x, w))


def regulate_tensor(x, blur, regulation_value, regulation_root=0.5):
    """util/regulator/gaussian_regulator_tensor.py:34-36: ``x * (v / pow(min(conv(x, blur), 1), root))``."""
    m = np.minimum(conv2d_same(x, blur), np.float32(1))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        p = np.power(m.astype(np.float64), float(np.float32(regulation_root))).astype(np.float32)
        gain = (np.float32(regulation_value) / p).astype(np.float32)
        return (x * gain).astype(np.float32)


def pad_inwards(x, paddings):
    """util/selection/isolate_rectangle.py:19-23: multiply by a 0/1 box (so NaN * 0 stays NaN)."""
    box = np.zeros(x.shape, dtype=np.float32)
    sl = tuple(slice(p[0], x.shape[i] - p[1]) for i, p in enumerate(paddings))
    box[sl] = 1
    with np.errstate(invalid="ignore"):
        return (box * x).astype(np.float32)


def get_value_from_color(x):
    """util/color/get_value.py:6-12: ``((c0 + c1) + ...) * float32(1 / C)``, keepdims."""
    acc = x[..., 0].astype(np.float32)
    for c in range(1, x.shape[-1]):
        acc = acc + x[..., c]
    div = np.float32(1.0) / np.float32(x.shape[-1])
    return (acc * div)[..., np.newaxis].astype(np.float32)


def _same_geometry(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2


def window_max(value, region_hw):
    """``tf.nn.max_pool(value, ksize=(1,h,w,1), strides=(1,rh,rw,1), 'SAME')`` (top_value_points.py:39-41)."""
    n, h, w, _ = value.shape
    rh, rw = int(region_hw[0]), int(region_hw[1])
    oh, pt = _same_geometry(h, h, rh)
    ow, pl = _same_geometry(w, w, rw)
    out = np.empty((n, oh, ow, 1), dtype=np.float32)
    for i in range(oh):
        ya, yb = max(i * rh - pt, 0), min(i * rh - pt + h, h)
        for j in range(ow):
            xa, xb = max(j * rw - pl, 0), min(j * rw - pl + w, w)
            out[:, i, j, 0] = np.max(value[:, ya:yb, xa:xb, 0], axis=(1, 2))   # NaN-propagating
    return out


def nearest_index(n_out, n_in):
    """``resize_images(NEAREST_NEIGHBOR)``, align_corners=False: ``min(floor(dst * float32(in/out)), in-1)``."""
    return np.minimum(np.floor(np.arange(n_out, dtype=np.float32) * np.float32(n_in / n_out)).astype(np.int64),
                      n_in - 1)


def max_value_indices_region(color, region_shape, value=None):
    """util/selection/top_value_points.py:32-45 -> int64 ``[K, 4]`` rows ``(n, y, x, 0)`` in row-major order."""
    if value is None:
        value = get_value_from_color(color)
    pooled = window_max(value, (region_shape[1], region_shape[2]))
    n, h, w, _ = color.shape
    up = pooled[:, nearest_index(h, pooled.shape[1])][:, :, nearest_index(w, pooled.shape[2])]
    with np.errstate(invalid="ignore"):
        return np.argwhere(value >= up).astype(np.int64)


def top_value_points(color, top_percent=0.1, value=None):
    """util/selection/top_value_points.py:8-29: keep ``color`` where ``value >= (1-p)*max + p*min`` per level."""
    if value is None:
        value = get_value_from_color(color)
    mx = np.max(value, axis=(1, 2), keepdims=True)
    mn = np.float32(-1.0) * np.max(-value, axis=(1, 2), keepdims=True)
    thr = (np.float32(1.0 - top_percent) * mx + np.float32(top_percent) * mn).astype(np.float32)
    with np.errstate(invalid="ignore"):
        keep = np.where(value >= thr, np.float32(1), np.float32(0))
        return (color * keep).astype(np.float32)


# ----------------------------------------------------------------------------------------------------------------------
# D1  composition  (recognition_testing.py:60-100)
# ----------------------------------------------------------------------------------------------------------------------


def line_end_stack(pyramid, weights, region_divisor=2.0):
    """S1-S8 in the order ``LineEndDisplayer.compile`` builds them.

    :param weights: dict with float64/32 HWIO arrays ``rgc``, ``rgby``, ``stripe``, ``blur``, ``end``.
    :return: dict of every intermediate plus ``points``.
    """
    x = np.asarray(pyramid, dtype=np.float32)
    a = conv_relu(x, weights["rgc"])                                    # :69
    b = conv_relu(a, weights["rgby"])                                   # :70
    c = conv_relu(b, weights["stripe"])                                 # :71, filters/orientation.py:24-29
    d = regulate_tensor(c, weights["blur"], 1.0, .1)                    # filters/orientation.py:33
    e = np.minimum(relu(conv2d_same(d, weights["end"])), np.float32(255))   # :73-74
    p = pad_inwards(e, [[0, 0], [2, 2], [2, 2], [0, 0]])                # :75
    g = get_value_from_color(p)                                         # :77
    n, h, w, ch = x.shape
    region = [1, int(h / region_divisor), int(w / region_divisor), ch]  # :40
    pts = max_value_indices_region(p, region, g)                        # :90-91
    return dict(rgc=a, rgby=b, stripe=c, orient=d, line_end=e, padded=p, gray=g, points=pts)


# ----------------------------------------------------------------------------------------------------------------------
# "Next" rows: centroids, boosting, the six display tensors of LineEndDisplayer.compile
# ----------------------------------------------------------------------------------------------------------------------


def index_tensor(h, w):
    """util/index_tensor.py:7-20 with ``are_dimensions_reversed``: ``ind[y, x] = (x, y)``, cast to float32."""
    ind = np.empty((h, w, 2), dtype=np.float32)
    ind[..., 0] = np.arange(w, dtype=np.float32)[np.newaxis, :]
    ind[..., 1] = np.arange(h, dtype=np.float32)[:, np.newaxis]
    return ind


def _strided_window_sum(x, rh, rw):
    """``tf.nn.convolution(x, ones, strides=[rh, rw], 'SAME')`` per channel: float64 sums over each window."""
    n, h, w, c = x.shape
    oh, pt = _same_geometry(h, rh, rh)
    ow, pl = _same_geometry(w, rw, rw)
    out = np.zeros((n, oh, ow, c), dtype=np.float64)
    for i in range(oh):
        ya, yb = max(i * rh - pt, 0), min(i * rh - pt + rh, h)
        for j in range(ow):
            xa, xb = max(j * rw - pl, 0), min(j * rw - pl + rw, w)
            out[:, i, j, :] = x[:, ya:yb, xa:xb, :].astype(np.float64).sum(axis=(1, 2))
    return out


def get_centroids_array(value, region_shape):
    """util/centroids.py:49-71: per ``region`` block, the value-weighted mean index ``(x, y)``; also the block totals."""
    value = np.asarray(value, dtype=np.float32)
    n, h, w, _ = value.shape
    rh, rw = int(region_shape[1]), int(region_shape[2])
    ind = index_tensor(h, w)
    biased = (ind[np.newaxis] * value).astype(np.float32)                       # :35  ind * to_channels(value, 2)
    centroid_pool = _strided_window_sum(biased, rh, rw).astype(np.float32)      # :36-39 identity 'additive' filter
    # :40-43: every tap weighs both (identical) channels by 1/2 -> the plain window sum of the value
    total_pool = _strided_window_sum(value, rh, rw).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        corrected = (centroid_pool / total_pool).astype(np.float32)             # :44 (0 / 0 = NaN on empty blocks)
    return corrected, total_pool


def get_centroids(value, region_shape):
    """util/centroids.py:21-46 -> (L1 distance of every pixel to its block's centroid ``[N,h,w,1]``, block totals)."""
    value = np.asarray(value, dtype=np.float32)
    n, h, w, _ = value.shape
    corrected, total_pool = get_centroids_array(value, region_shape)
    resized = corrected[:, nearest_index(h, corrected.shape[1])][:, :, nearest_index(w, corrected.shape[2])]   # :45
    dist = np.abs(resized - index_tensor(h, w)[np.newaxis]).astype(np.float32)  # :46 math.equality_distance
    return (dist[..., 0] + dist[..., 1])[..., np.newaxis].astype(np.float32), total_pool


def resize_nearest_neighbor(x, size):
    """``tf.image.resize_nearest_neighbor`` (align_corners False)."""
    return x[:, nearest_index(int(size[0]), x.shape[1])][:, :, nearest_index(int(size[1]), x.shape[2])]


def max_pool_3x3(x):
    """``tf.nn.max_pool(x, (1,3,3,1), (1,1,1,1), 'SAME')``, NaN-propagating."""
    n, h, w, c = x.shape
    out = np.empty_like(x)
    for i in range(h):
        for j in range(w):
            out[:, i, j, :] = np.max(x[:, max(i - 1, 0):i + 2, max(j - 1, 0):j + 2, :], axis=(1, 2))
    return out


def get_boosting(inp, energy, exhaustion_max=1, excitation_max=1, input_based_recovery=False, constant_recovery=True):
    """util/energy/boosting.py:10-42 (+ recovery.py:4-22). Returns (has_fired, new_energy); the caller keeps the state."""
    inp = np.asarray(inp, dtype=np.float32)
    energy = np.asarray(energy, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        biased = np.power(inp.astype(np.float64), energy.astype(np.float64)).astype(np.float32)   # :17
    pooled = max_pool_3x3(biased)                                                                  # :18
    fired = np.where(biased == pooled, np.float32(1), np.float32(0))                               # :20-22
    strength = (fired * inp).astype(np.float32)                                                    # :24
    exhaustion = (fired * np.float32(255.0)).astype(np.float32)                                    # :26
    if input_based_recovery and not constant_recovery:
        recovery = (strength * np.float32(0.8)).astype(np.float32)
    elif constant_recovery and not input_based_recovery:
        recovery = np.full_like(strength, 10)
    elif input_based_recovery and constant_recovery:
        recovery = np.maximum((strength * np.float32(0.8)).astype(np.float32), np.full_like(strength, 10))
    else:
        raise ValueError("You must choose a type of recovery")
    t = (((energy * np.float32(255.0)).astype(np.float32) - exhaustion).astype(np.float32) + recovery).astype(np.float32)
    t = (t / np.float32(255.0)).astype(np.float32)                                                 # :30-33
    new_energy = np.minimum(np.maximum(t, np.float32(-exhaustion_max)), np.float32(excitation_max))
    return fired, new_energy.astype(np.float32)


def display_tensors(orient, padded, gray, energy, centroid_region=(1, 3, 3)):
    """recognition_testing.py:77-100: the six tensors ``run()`` fetches, and the boosting state after this frame."""
    scaled = (gray / np.float32(255.0)).astype(np.float32)
    centroids, importances = get_centroids(scaled, centroid_region)                                 # :79-80
    importances = (np.minimum(np.maximum((importances * np.float32(255 / 4.0)).astype(np.float32), np.float32(1)),
                              np.float32(256)) - np.float32(1)).astype(np.float32)                  # :81
    half = (np.asarray(gray.shape[1:3], dtype=np.float32) / np.float32(np.e ** .5)).astype(np.int32)   # :82
    im2 = resize_nearest_neighbor(gray, half)                                                       # :83
    centroids2, _ = get_centroids((im2 / np.float32(255.0)).astype(np.float32), centroid_region)    # :84
    fired, new_energy = get_boosting(importances, energy)                                           # :86-87
    fired_rgb = np.repeat((fired * importances).astype(np.float32), 3, axis=-1)                     # boosting.py:36
    update_rgb = np.repeat((new_energy * np.float32(127.5) + np.float32(127.5)).astype(np.float32), 3, axis=-1)   # :37-39
    outs = [orient, (np.float32(255) - centroids * np.float32(255)).astype(np.float32),
            (np.float32(255) - centroids2 * np.float32(255)).astype(np.float32),
            (fired_rgb * np.float32(255)).astype(np.float32), update_rgb, padded]                   # :98-100
    return outs, new_energy


# ----------------------------------------------------------------------------------------------------------------------
# The reference's CPU path as it runs: scipy's own zoom (from_image.py:55-59). bench.py times this (kind
# "reference-python"); tests check it against from_image above.
# ----------------------------------------------------------------------------------------------------------------------

def from_image_scipy(image, num_colors, center_dimensions, scale):
    """``zoom.from_image`` with the reference's own arithmetic: per level and channel one
    ``scipy.ndimage.zoom(crop, 1 / scale**s, order=5, prefilter=False)`` call (from_image.py:48-64), single-threaded
    like the reference. Undefined tail rows / columns are 0 (see ``from_image``)."""
    from scipy import ndimage
    image = np.asarray(image, dtype=np.float32)
    hw = image.shape[:-1]
    h, w = list(reversed(list(center_dimensions)))
    levels = pyramid_levels(hw, center_dimensions, scale)
    out = np.zeros((max(levels, 0), h, w, num_colors), dtype=np.float32)
    for s in range(levels):
        (y0, y1), (x0, x1) = level_crop(hw, center_dimensions, scale, s)
        crop = image[y0:y1, x0:x1]
        for c in range(num_colors):
            z = ndimage.zoom(crop[:, :, c], 1.0 / (scale ** s), order=5, prefilter=False)
            ym, xm = min(h, z.shape[0]), min(w, z.shape[1])
            out[s, :ym, :xm, c] = z[:ym, :xm]
    return out
