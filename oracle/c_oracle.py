"""TEST INFRASTRUCTURE ONLY -- ctypes front end of ``oracle/silent_oracle.c`` (the bit-defined float32 oracle).

``build()`` compiles the C file with gcc into ``oracle/_build/libsilent_oracle.so``; the wrappers take / return numpy
arrays. Only tests, ``__graft_entry__`` and ``bench.py``'s CPU-baseline legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import silent_oracle as lit

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "silent_oracle.c")


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            text = f.read()
        return " fma " in text, " avx2 " in text
    except OSError:
        return False, False


# With -mfma, fmaf() compiles to one vfmadd instruction instead of a libm call: same bits, ~20x faster; -O3 -mavx2 on top
# lets gcc vectorise the tap loops (no -ffast-math, no contraction: same bits again, checked by hash; 1.7x faster), so the
# CPU arm of the bench is not slower than it has to be. The flavour is part of the file name so a .so built on one host is
# never loaded on a CPU without the instructions it uses.
_FMA, _AVX2 = _cpu_flags()
_FLAVOUR = "_fma_avx2" if _FMA and _AVX2 else "_fma" if _FMA else ""
_OUT = os.path.join(_HERE, "_build", "libsilent_oracle%s.so" % _FLAVOUR)
_lib = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force=False):
    if not force and os.path.exists(_OUT) and os.path.getmtime(_OUT) >= os.path.getmtime(_SRC):
        return _OUT
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    subprocess.check_call(["gcc", "-O3" if _AVX2 else "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math"] +
                          (["-mfma"] if _FMA else []) + (["-mavx2"] if _FMA and _AVX2 else []) + ["-o", _OUT, _SRC, "-lm"])
    return _OUT


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        c_int, c_float, c_size = ctypes.c_int, ctypes.c_float, ctypes.c_size_t
        L.so_canon_pow.restype = c_float
        L.so_canon_pow.argtypes = [c_float, c_float]
        L.so_pyramid.restype = None
        L.so_pyramid.argtypes = [ctypes.c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _i32p,
                                 _f32p, _u8p, _i32p, _f32p, _u8p, _i32p, _f32p]
        L.so_conv2d.restype = None
        L.so_conv2d.argtypes = [_f32p, c_int, c_int, c_int, c_int, _f32p, c_int, c_int, c_int, c_int, c_float, _f32p]
        L.so_regulate.restype = None
        L.so_regulate.argtypes = [_f32p, c_int, c_int, c_int, c_int, _f32p, c_int, c_float, c_float, _f32p, _f32p]
        L.so_pad_inwards.restype = None
        L.so_pad_inwards.argtypes = [_f32p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _f32p]
        L.so_value_from_color.restype = None
        L.so_value_from_color.argtypes = [_f32p, c_size, c_int, _f32p]
        L.so_max_value_indices_region.restype = ctypes.c_int64
        L.so_max_value_indices_region.argtypes = [_f32p, c_int, c_int, c_int, c_int, c_int, _i64p, ctypes.c_int64]
        L.so_top_value_points.restype = None
        L.so_top_value_points.argtypes = [_f32p, _f32p, c_int, c_int, c_int, c_int, ctypes.c_double, _f32p]
        L.so_line_end_stack.restype = None
        L.so_line_end_stack.argtypes = [_f32p, c_int, c_int, c_int, _f32p, _f32p, _f32p, _f32p, c_int, _f32p] + \
            [_f32p] * 8
        L.so_line_end_stack_fused.restype = None
        L.so_line_end_stack_fused.argtypes = L.so_line_end_stack.argtypes
        for name in ("so_depthwise3", "so_rgby_shared", "so_stripe_sym180", "so_end_ownoth"):
            getattr(L, name).restype = c_int
            getattr(L, name).argtypes = [_f32p]
        L.so_rgc_rgby_fused.restype = None
        L.so_rgc_rgby_fused.argtypes = [_f32p, c_int, c_int, c_int, _f32p, _f32p, _f32p, _f32p]
        L.so_get_centroids.restype = None
        L.so_get_centroids.argtypes = [_f32p, c_int, c_int, c_int, c_int, c_int, _f32p, _f32p, _f32p]
        L.so_resize_nearest.restype = None
        L.so_resize_nearest.argtypes = [_f32p, c_int, c_int, c_int, c_int, c_int, c_int, _f32p]
        L.so_get_boosting.restype = None
        L.so_get_boosting.argtypes = [_f32p, _f32p, c_int, c_int, c_int, c_float, c_float, c_int, _f32p, _f32p]
        L.so_pointwise.restype = None
        L.so_pointwise.argtypes = [_f32p, _f32p, c_size, c_int, _f32p]
        _lib = L
    return _lib


def _c(a, dt=np.float32):
    return np.ascontiguousarray(np.asarray(a), dtype=dt)


def canon_pow(x, r):
    return lib().so_canon_pow(float(np.float32(x)), float(np.float32(r)))


def pyramid_tables(image_hw, center_wh, scale):
    """Tap tables for every level (float64 math of scipy's zoom, weights rounded to float32, absolute frame indices)."""
    H, W = image_hw
    h, w = list(reversed(list(center_wh)))
    L = max(lit.pyramid_levels(image_hw, center_wh, scale), 0)
    iy = np.zeros((L, h, 6), np.int32)
    wy = np.zeros((L, h, 6), np.float32)
    oky = np.zeros((L, h), np.uint8)
    ix = np.zeros((L, w, 6), np.int32)
    wx = np.zeros((L, w, 6), np.float32)
    okx = np.zeros((L, w), np.uint8)
    valid = np.zeros((L, 2), np.int32)
    for s in range(L):
        (y0, y1), (x0, x1) = lit.level_crop(image_hw, center_wh, scale, s)
        f = 1.0 / (scale ** s)
        oh, ow = int(round((y1 - y0) * f)), int(round((x1 - x0) * f))
        ty, gy, ky = lit.zoom_axis_table(y1 - y0, oh)
        tx, gx, kx = lit.zoom_axis_table(x1 - x0, ow)
        vh, vw = min(h, oh), min(w, ow)
        valid[s] = (vh, vw)
        iy[s, :vh] = ty[:vh] + y0
        wy[s, :vh] = gy[:vh]
        oky[s, :vh] = ky[:vh]
        ix[s, :vw] = tx[:vw] + x0
        wx[s, :vw] = gx[:vw]
        okx[s, :vw] = kx[:vw]
    return dict(L=L, h=h, w=w, iy=iy, wy=wy, oky=oky, ix=ix, wx=wx, okx=okx, valid=valid)


def from_image(frames, num_colors, center_wh, scale):
    """frames ``[B,H,W,FC]`` (or ``[H,W,FC]``) uint8 / float32 -> ``[B*L,h,w,num_colors]`` float32."""
    frames = np.asarray(frames)
    if frames.ndim == 3:
        frames = frames[np.newaxis]
    is_u8 = frames.dtype == np.uint8
    frames = np.ascontiguousarray(frames, dtype=np.uint8 if is_u8 else np.float32)
    B, H, W, FC = frames.shape
    t = pyramid_tables((H, W), center_wh, scale)
    out = np.zeros((B * t["L"], t["h"], t["w"], num_colors), np.float32)
    if t["L"] > 0:
        lib().so_pyramid(frames.ctypes.data, int(is_u8), B, H, W, FC, t["L"], t["h"], t["w"], num_colors, t["iy"],
                         t["wy"], t["oky"], t["ix"], t["wx"], t["okx"], t["valid"], out)
    return out


def conv2d(x, w, post=0, clip_hi=0.0):
    x, w = _c(x), _c(w)
    n, h, wd, cin = x.shape
    kh, kw, wc, cout = w.shape
    assert wc == cin
    out = np.empty((n, h, wd, cout), np.float32)
    lib().so_conv2d(x, n, h, wd, cin, w, kh, kw, cout, post, clip_hi, out)
    return out


def regulate_tensor(x, blur, value, root=0.5):
    x, blur = _c(x), _c(blur)
    n, h, w, c = x.shape
    out, tmp = np.empty_like(x), np.empty_like(x)
    lib().so_regulate(x, n, h, w, c, blur, blur.shape[0], value, root, tmp, out)
    return out


def pad_inwards(x, paddings):
    x = _c(x)
    n, h, w, c = x.shape
    out = np.empty_like(x)
    lib().so_pad_inwards(x, n, h, w, c, paddings[1][0], paddings[1][1], paddings[2][0], paddings[2][1], out)
    return out


def get_value_from_color(x):
    x = _c(x)
    out = np.empty(x.shape[:-1] + (1,), np.float32)
    lib().so_value_from_color(x, x.size // x.shape[-1], x.shape[-1], out)
    return out


def max_value_indices_region(value, region_hw, capacity=None):
    value = _c(value)
    n, h, w = value.shape[:3]
    cap = int(capacity if capacity is not None else n * h * w)
    out = np.zeros((max(cap, 1), 4), np.int64)
    count = lib().so_max_value_indices_region(value, n, h, w, int(region_hw[0]), int(region_hw[1]), out, cap)
    return out[:min(count, cap)], count


def top_value_points(color, value, top_percent=0.1):
    color, value = _c(color), _c(value)
    n, h, w, c = color.shape
    out = np.empty_like(color)
    lib().so_top_value_points(color, value, n, h, w, c, top_percent, out)
    return out


def line_end_stack(pyramid, weights, region_divisor=2.0, order="fused"):
    """S1-S8 on an NHWC pyramid. ``order``: "fused" = the canonical order of the fused stack kernels (structured
    convolutions evaluated through their shared sub-kernels, so_line_end_stack_fused); "operator" = every stage as the
    stand-alone operator evaluates it (one (ky, ci, kx) chain per output, so_line_end_stack)."""
    assert order in ("fused", "operator")
    x = _c(pyramid)
    n, h, w, ch = x.shape
    assert ch == 3
    W = {k: _c(v) for k, v in weights.items()}
    bufs = {k: np.empty_like(x) for k in ("rgc", "rgby", "stripe", "orient", "line_end", "padded")}
    gray = np.empty((n, h, w, 1), np.float32)
    scratch = np.empty((2,) + x.shape, np.float32)
    fn = lib().so_line_end_stack_fused if order == "fused" else lib().so_line_end_stack
    fn(x, n, h, w, W["rgc"], W["rgby"], W["stripe"], W["blur"], W["blur"].shape[0], W["end"],
       bufs["rgc"], bufs["rgby"], bufs["stripe"], bufs["orient"], bufs["line_end"], bufs["padded"], gray, scratch)
    region = (int(h / region_divisor), int(w / region_divisor))
    if min(region) >= 1:
        pts, count = max_value_indices_region(gray, region)
    else:   # a 1-pixel-high or -wide level has no valid region shape (the reference's max_pool would reject stride 0)
        pts = np.zeros((0, 4), np.int64)
    bufs.update(gray=gray, points=pts)
    return bufs


def bank_stack(pyramid, f, region_hw, order="fused"):
    """BASELINE config C4: S1-S2 on 3 channels, S3-S7 on the orientation bank ``f`` (dict rgc, rgby, stripe [3,3,3,C],
    blur [7,7,C,C], end [3,3,C,C]), composed like ``compile()`` (recognition_testing.py:69-77). ``order="fused"``: rgc /
    rgby as the fused kernels evaluate them (shared-surround order); every other stage is the per-operator chain (the
    bank kernel skips exact-zero weights only, which does not change a bit)."""
    x = _c(pyramid)
    n, h, w, _ = x.shape
    if order == "fused":
        a, b = np.empty_like(x), np.empty_like(x)
        lib().so_rgc_rgby_fused(x, n, h, w, _c(f["rgc"]), _c(f["rgby"]), a, b)
    else:
        b = conv2d(conv2d(x, f["rgc"], post=1), f["rgby"], post=1)
    orient = regulate_tensor(conv2d(b, f["stripe"], post=1), f["blur"], 1.0, .1)
    line_end = conv2d(orient, f["end"], post=2, clip_hi=255.0)
    padded = pad_inwards(line_end, [[0, 0], [2, 2], [2, 2], [0, 0]])
    gray = get_value_from_color(padded)
    points, _ = max_value_indices_region(gray, region_hw)
    return dict(orient=orient, padded=padded, gray=gray, points=points)


# ---- "next" rows: centroids, nearest resize, boosting, display arithmetic ------------------------------------------------

def _same_out(n, s):
    return -(-n // s)


def get_centroids(value, region_shape):
    """-> (centroids [N,h,w,1], total_pool [N,oh,ow,1], corrected [N,oh,ow,2]) (util/centroids.py:21-71)."""
    value = _c(value)
    n, h, w = value.shape[:3]
    rh, rw = int(region_shape[1]), int(region_shape[2])
    oh, ow = _same_out(h, rh), _same_out(w, rw)
    corrected = np.empty((n, oh, ow, 2), np.float32)
    total = np.empty((n, oh, ow, 1), np.float32)
    cent = np.empty((n, h, w, 1), np.float32)
    with np.errstate(all="ignore"):
        lib().so_get_centroids(value, n, h, w, rh, rw, corrected, total, cent)
    return cent, total, corrected


def resize_nearest(x, size):
    x = _c(x)
    n, h, w, c = x.shape
    out = np.empty((n, int(size[0]), int(size[1]), c), np.float32)
    lib().so_resize_nearest(x, n, h, w, c, int(size[0]), int(size[1]), out)
    return out


def get_boosting(inp, energy, exhaustion_max=1.0, excitation_max=1.0, recovery_mode=1):
    """-> (has_fired, new_energy); ``energy`` is not modified."""
    inp, energy = _c(inp), _c(energy).copy()
    n, h, w = inp.shape[:3]
    fired = np.empty_like(inp)
    scratch = np.empty_like(inp)
    lib().so_get_boosting(inp, energy, n, h, w, exhaustion_max, excitation_max, recovery_mode, fired, scratch)
    return fired, energy


def pointwise(x, kind, y=None):
    x = _c(x)
    out = np.empty_like(x)
    lib().so_pointwise(x, _c(y) if y is not None else x, x.size, kind, out)
    return out


def display_tensors(orient, padded, gray, energy):
    """recognition_testing.py:77-100 in the canonical order: the six fetched tensors and the new boosting state."""
    scaled = pointwise(gray, 0)
    centroids, total, _ = get_centroids(scaled, (1, 3, 3))
    importances = pointwise(total, 2)
    half = (np.asarray(gray.shape[1:3], dtype=np.float32) / np.float32(np.e ** .5)).astype(np.int32)
    im2 = resize_nearest(gray, half)
    centroids2, _, _ = get_centroids(pointwise(im2, 0), (1, 3, 3))
    fired, new_energy = get_boosting(importances, energy)
    fired_rgb = np.repeat(pointwise(pointwise(fired, 5, importances), 3), 3, axis=-1)
    update_rgb = np.repeat(pointwise(new_energy, 4), 3, axis=-1)
    return [orient, pointwise(centroids, 1), pointwise(centroids2, 1), fired_rgb, update_rgb, padded], new_energy
