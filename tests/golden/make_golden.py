#!/usr/bin/env python
"""Generate the golden fixtures in this directory by executing the REFERENCE's own code.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

What is executed, unmodified, from ``/root/reference/slam_recognition``:

* every numpy weight generator (``constant_convolutions/**``, ``util/normalize``, ``util/orientation``) ->
  ``generators.npz``;
* the pyramid builder ``util/zoom/from_image.py`` -> ``pyramid.npz``. The reference indexes arrays with LISTS of slices,
  which numpy >= 1.23 rejects, so the module's ``np`` / ``ndimage`` handles are replaced by thin proxies that hand out an
  ndarray subclass accepting list indices (the function body itself runs as written). ``np.empty`` is served
  zero-filled so the uninitialised tail rows/columns (``from_image.py:53``) are defined as 0;
* the filter callables and selection ops (``filters/*.py``, ``util/apply_filter.py``, ``util/regulator``,
  ``util/selection``, ``util/color/get_value.py``) composed in the order of ``recognition_testing.py:69-77,90-91`` ->
  ``stack.npz``; and ``util/centroids.py`` + ``util/energy/boosting.py`` composed as ``recognition_testing.py:49-57,
  77-100`` for three consecutive frames -> ``display.npz``. TensorFlow 1.x is not installable here, so these run on ``oracle/tf1_shim.py``, an eager numpy
  restatement of the ~25 TF-1 symbols they call. Results at that boundary are therefore "parity unpinned" against real
  TensorFlow; the composition, constants and weights are the reference's own.

No reference source is copied; only outputs are stored.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = "/root/reference"

sys.path.insert(0, ROOT)
from oracle import tf1_shim  # noqa: E402


def import_reference():
    tf1_shim.install()
    sys.path.insert(0, REFERENCE)
    return importlib.import_module("slam_recognition")


class _ListIndexArray(np.ndarray):
    """ndarray that accepts the pre-1.23 ``a[[slice, slice, ...]]`` spelling of ``a[(slice, slice, ...)]``."""

    @staticmethod
    def _fix(key):
        if isinstance(key, list) and any(isinstance(k, slice) or k is None for k in key):
            return tuple(key)
        return key

    def __getitem__(self, key):
        return super().__getitem__(self._fix(key))

    def __setitem__(self, key, value):
        return super().__setitem__(self._fix(key), value)


class _NumpyProxy:
    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def empty(shape, *a, **k):
        return np.zeros(shape, *a, **k).view(_ListIndexArray)


class _NdimageProxy:
    def __init__(self, real):
        self._real = real

    def zoom(self, *a, **k):
        return np.asarray(self._real.zoom(*a, **k)).view(_ListIndexArray)


def reference_from_image(sr, image, num_colors, center, scale):
    mod = importlib.import_module("slam_recognition.util.zoom.from_image")
    from scipy import ndimage
    mod.np, mod.ndimage = _NumpyProxy(), _NdimageProxy(ndimage)
    try:
        return np.asarray(mod.image_to_zoom_tensor(image.view(_ListIndexArray), num_colors, center, scale))
    finally:
        mod.np, mod.ndimage = np, ndimage


def frame(seed, h, w, c=3, kind="noise"):
    rs = np.random.RandomState(seed)
    if kind == "noise":
        return rs.randint(0, 256, size=(h, w, c)).astype(np.uint8)
    if kind == "flat":
        # noise frame with (i) a flat-40 patch: stripe response there is rounding noise -> blur << 1 -> large gain, and
        # (ii) a small black patch in one corner of the level-0 crop: everything exactly 0 -> blur 0 -> 0 * inf = NaN
        # (S4), which poisons only the region windows that contain it.
        img = rs.randint(0, 256, size=(h, w, c)).astype(np.uint8)
        img[(5 * h) // 7: (13 * h) // 14, w // 10: (9 * w) // 20, :] = 40
        img[(5 * h) // 14: (33 * h) // 70, (7 * w) // 20: (47 * w) // 100, :] = 0
        return img
    # "natural": smooth gradients + bars + flat regions; exercises the regulator's m < 1 and NaN paths
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, c))
    for ch in range(c):
        img[..., ch] = 96 + 80 * np.sin(xx / (9.0 + 3 * ch)) * np.cos(yy / (7.0 + 2 * ch))
    img[h // 5: h // 5 + 3, :, :] = 250
    img[:, w // 3: w // 3 + 2, 1] = 5
    img[h // 2:, w // 2:, :] = 40          # flat quadrant -> zero stripe response -> 0 * inf
    img[: h // 4, : w // 4, :] = 0
    img += rs.randint(0, 3, size=img.shape)
    img[h // 2 + 4:, w // 2 + 4:, :] = 40
    return np.clip(img, 0, 255).astype(np.uint8)


def make_generators(sr):
    cc = importlib.import_module("slam_recognition.constant_convolutions")
    cs = importlib.import_module("slam_recognition.constant_convolutions.center_surround.rgc")
    et = importlib.import_module("slam_recognition.constant_convolutions.edge_orientation_detector.edge_tensor")
    norm = importlib.import_module("slam_recognition.util.normalize").normalize_tensor_positive_negative
    ori = importlib.import_module("slam_recognition.util.orientation")
    out = {}
    out["cs_1d_test"] = cc.center_surround_tensor(1, [0, 1, 0], [1, 0, 0], [0, 0, 1], [1, 0, 0])
    out["cs_2d_test"] = cc.center_surround_tensor(2, [0, 1, 0], [1, 0, 0], [0, 0, 1], [1, 0, 0])
    out["cs_3d"] = cc.center_surround_tensor(3, [1, .5], [1, -2], [.25, 1], [-1, 3])
    out["midget_rgc_2"] = cc.midget_rgc(2)
    out["midget_rgc_1"] = cc.midget_rgc(1)
    out["midget_rgc_full_2"] = cs.midget_rgc_full(2)
    out["rgby_2"] = cc.rgby(2)
    out["rgby_3_2"] = cc.rgby_3(2)
    out["rgby_3_3"] = cc.rgby_3(3)
    out["rgb_2d_stripe"] = cc.rgb_2d_stripe_tensors()
    out["rgb_2d_stripe_in"] = cc.rgb_2d_stripe_tensors(in_channel=(1, .5, 0))
    out["stripe_3d"] = cc.stripe_tensor([0.0, 0.6, 0.8], [1, 0], [2, 1], [1, 1], [-1, .5])
    out["rgb_2d_edge"] = cc.rgb_2d_edge_tensors()
    out["rgb_2d_edge_time_diff"] = cc.rgb_2d_edge_tensors_time_diff()
    out["rgb_2d_end_7x7"] = et.rgb_2d_end_tensors()
    out["rgb_2d_end"] = cc.rgb_2d_end_tensors()
    out["blur_2_7"] = cc.blur_tensor(2, lengths=7)
    out["blur_2_default"] = cc.blur_tensor(2)
    out["blur_3_list"] = cc.blur_tensor(3, lengths=[3, 5, 3], channels_in=2, channels_out=1)
    # the 8-orientation bank BASELINE config C4 uses (general per-vector API)
    vecs = [np.array([np.cos(k * np.pi / 4), np.sin(k * np.pi / 4)]) for k in range(8)]
    eye = np.eye(8)
    spread = [1, 1, 1, 0, 0, 0, 0, 0]
    out["stripe_8"] = sum(cc.stripe_tensor(v, spread, list(eye[k] * 4), spread, list(-eye[k] * 4))
                          for k, v in enumerate(vecs))
    out["end_8"] = sum(cc.end_tensor(3 * v, list(eye[k]), list(.25 * eye[k]), list(eye[k]), list(.5 * eye[k]))
                       for k, v in enumerate(vecs))
    out["blur_8"] = cc.blur_tensor(2, 7, channels_in=8, channels_out=8)
    out["simplex_2"] = ori.simplex_coordinates(2)
    out["simplex_3"] = ori.simplex_coordinates(3)
    out["simplex_5"] = ori.simplex_coordinates(5)
    out["above_axis_simplex_3"] = ori.above_axis_simplex_coordinates(3)
    t = np.squeeze(cc.center_surround_tensor(2, [1], [1], [1], [-1]))
    out["norm_in"] = t.copy()
    out["norm_out"] = norm(t)
    t2 = np.random.RandomState(7).randn(4, 5, 3)
    out["norm_rand_in"] = t2.copy()
    out["norm_rand_out"] = norm(t2, 3.0, 0.5)
    np.savez_compressed(os.path.join(HERE, "generators.npz"), **out)
    return out


PYRAMID_CASES = [
    # name, H, W, C, center (w, h), scale, seed
    ("small_e", 70, 90, 3, (24, 16), float(np.e ** .5), 11),
    ("small_r2", 97, 131, 3, (24, 16), float(2 ** .5), 12),
    ("ragged", 61, 53, 3, (20, 12), 1.3, 13),
    ("gray", 50, 64, 1, (16, 16), 1.5, 14),
    ("vga_like", 120, 160, 3, (72, 48), 1.3, 15),
]


def make_pyramids(sr):
    out = {}
    for name, h, w, c, center, scale, seed in PYRAMID_CASES:
        img = frame(seed, h, w, c).astype(np.float32)
        pyr = reference_from_image(sr, img, c, list(center), scale)
        out[name + "_image"] = img.astype(np.uint8)
        out[name + "_pyramid"] = pyr.astype(np.float32)
        assert np.array_equal(pyr.astype(np.float32).astype(np.float64), pyr)  # holds fp32-rounded values
        out[name + "_params"] = np.array([center[0], center[1], scale], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "pyramid.npz"), **out)
    return out


def reference_stack(sr, pyramid, region_divisor=2.0):
    """recognition_testing.py:69-77 and :90-91, executed from the reference's own modules on the TF-1 shim."""
    tf = sys.modules["tensorflow"]
    filters = importlib.import_module("slam_recognition.filters")
    apply_filter = importlib.import_module("slam_recognition.util.apply_filter").apply_filter
    sel = importlib.import_module("slam_recognition.util.selection")
    get_value = importlib.import_module("slam_recognition.util.color.get_value").get_value_from_color
    end = importlib.import_module("slam_recognition.constant_convolutions.oriented_end_detector").rgb_2d_end_tensors()

    x = tf.constant(pyramid, dtype=tf.float32)
    rgc = filters.rgc_filter(x)
    rgby = filters.rgby_filter(rgc)
    orient = filters.orientation_filter(rgby)
    line_end = tf.maximum(apply_filter(orient, end), [0])
    line_end = tf.clip_by_value(line_end, 0, 255)
    padded = sel.pad_inwards(line_end, [[0, 0], [2, 2], [2, 2], [0, 0]])
    gray = get_value(padded)
    h, w, c = pyramid.shape[1:]
    region_shape = tf.TensorShape([1, h / region_divisor, w / region_divisor, c])   # recognition_testing.py:40
    points = sel.max_value_indices_region(padded, region_shape, gray)
    top = sel.top_value_points(padded, 0.1, gray)
    return dict(rgc=rgc.numpy(), rgby=rgby.numpy(), orient=orient.numpy(), line_end=line_end.numpy(),
                padded=padded.numpy(), gray=gray.numpy(), points=points.numpy(), top=top.numpy())


STACK_CASES = [
    # name, H, W, center (w,h), scale, seed, kind
    ("noise", 97, 131, (24, 16), float(2 ** .5), 21, "noise"),
    ("noise_odd", 83, 117, (36, 20), 1.3, 22, "noise"),
    ("natural", 140, 200, (48, 32), float(2 ** .5), 23, "natural"),
    ("flat", 140, 200, (48, 32), float(2 ** .5), 24, "flat"),
]


def make_stacks(sr):
    out = {}
    for name, h, w, center, scale, seed, kind in STACK_CASES:
        img = frame(seed, h, w, 3, kind)
        pyr = reference_from_image(sr, img.astype(np.float32), 3, list(center), scale).astype(np.float32)
        res = reference_stack(sr, pyr)
        out[name + "_image"] = img
        out[name + "_params"] = np.array([center[0], center[1], scale], dtype=np.float64)
        out[name + "_pyramid"] = pyr
        for k, v in res.items():
            out[name + "_" + k] = v
        print(name, "levels", pyr.shape[0], "points", res["points"].shape[0],
              "nan_frac(orient)", float(np.isnan(res["orient"]).mean()))
    np.savez_compressed(os.path.join(HERE, "stack.npz"), **out)
    return out


def reference_display(padded, gray, steps=3):
    """recognition_testing.py:49-57 (pre_compile: the boosting state) and :77-100 (compile) on the reference's own
    ``get_centroids`` / ``get_boosting``, fed the same padded/gray tensors ``steps`` times (a still camera): the six
    fetched tensors of every step and the state after it."""
    import math as m
    tf = sys.modules["tensorflow"]
    get_centroids = importlib.import_module("slam_recognition.util").get_centroids
    boosting = importlib.import_module("slam_recognition.util.energy.boosting")
    gray_t = tf.constant(gray, dtype=tf.float32)
    centroid_region_shape = [1, 3, 3]
    _, imp0 = get_centroids(gray_t / 255.0, centroid_region_shape, debug=True)
    energy_values = boosting.initialize_boosting(imp0 * 255)
    out = []
    for _ in range(steps):
        centroids, centroid_importances = get_centroids(gray_t / 255.0, centroid_region_shape, debug=True)
        centroid_importances = tf.clip_by_value(centroid_importances * (255 / 4.0), 1, 256) - 1
        half_shape = tf.cast(gray_t.shape[1:3], tf.float32) / tf.constant(m.e ** .5)
        im2 = tf.image.resize_nearest_neighbor(gray_t, tf.cast(half_shape, tf.int32))
        centroids2, _ = get_centroids(im2 / 255.0, centroid_region_shape, debug=True)
        fired_importants, update_importances = boosting.get_boosting(centroid_importances, energy_values,
                                                                     for_visualizing=True)
        out.append(dict(centroids=(255 - centroids * 255).numpy(), centroids2=(255 - centroids2 * 255).numpy(),
                        fired=(fired_importants * 255).numpy(), update=update_importances.numpy(),
                        energy=energy_values.numpy().copy()))
    return out


def make_display(stacks):
    out = {}
    for name in ("noise", "natural", "flat"):
        steps = reference_display(stacks[name + "_padded"], stacks[name + "_gray"])
        for i, st in enumerate(steps):
            for k, v in st.items():
                out["%s_step%d_%s" % (name, i, k)] = v.astype(np.float32)
        print("display", name, "fired pixels per step", [int((st["fired"] > 0).sum()) for st in steps])
    np.savez_compressed(os.path.join(HERE, "display.npz"), **out)
    return out


if __name__ == "__main__":
    sr = import_reference()
    g = make_generators(sr)
    print("generators:", len(g))
    p = make_pyramids(sr)
    print("pyramids:", [k for k in p if k.endswith("_pyramid")], [p[k].shape for k in p if k.endswith("_pyramid")])
    make_display(make_stacks(sr))
