"""Pins the two oracles: the literal numpy oracle against outputs of the reference's own code (golden fixtures) and
against scipy live; the bit-defined C oracle against the literal one to the north-star tolerance."""
import numpy as np
import pytest

from conftest import structured_frame, synthetic_frame
from oracle import silent_oracle as lit


def test_pyramid_oracle_equals_reference_from_image(goldens):
    P = goldens["pyramid"]
    names = [k[:-8] for k in P.files if k.endswith("_pyramid")]
    assert len(names) >= 5
    for name in names:
        img = P[name + "_image"]
        cw, ch, sc = P[name + "_params"]
        got = lit.from_image(img.astype(np.float32), img.shape[2], [int(cw), int(ch)], float(sc))
        assert np.array_equal(got, P[name + "_pyramid"]), name        # bit-for-bit what scipy produced for the reference


def test_zoom_restatement_against_scipy_live():
    ndimage = pytest.importorskip("scipy.ndimage")
    rs = np.random.RandomState(5)
    for (h, w, f) in [(40, 57, 1.0), (61, 90, 1 / 1.3), (200, 311, 1 / 2 ** 1.5), (33, 33, 1 / 5.65), (7, 9, 0.5)]:
        a = (rs.rand(h, w) * 255).astype(np.float32)
        want = ndimage.zoom(a, f, order=5, prefilter=False)
        got = lit.zoom_order5(a, f)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 1e-6 * 255, (h, w, f)


def test_level_geometry_matches_survey():
    # level counts of the BASELINE configs
    assert lit.pyramid_levels((480, 640), (288, 192), 1.3) == 4
    assert lit.pyramid_levels((480, 640), (288, 192), np.e ** .5) == 2
    assert lit.pyramid_levels((1080, 1920), (288, 192), 2 ** .5) == 6
    assert lit.pyramid_levels((2160, 3840), (288, 192), 2 ** .5) == 8
    assert lit.pyramid_levels((720, 1280), (288, 192), 2 ** .5) == 5
    assert lit.level_crop((1080, 1920), (288, 192), 2 ** .5, 5) == [(0, 1080), (145, 1774)]


def test_stack_oracle_equals_reference_code_on_shim(goldens, default_filters):
    S = goldens["stack"]
    for name in ("noise", "noise_odd", "natural", "flat"):
        r = lit.line_end_stack(S[name + "_pyramid"], default_filters)
        for key in ("rgc", "rgby", "orient", "line_end", "padded", "gray"):
            assert np.array_equal(r[key], S[name + "_" + key], equal_nan=True), (name, key)
        assert np.array_equal(r["points"], S[name + "_points"]), name
        top = lit.top_value_points(r["padded"], 0.1, r["gray"])
        assert np.array_equal(top, S[name + "_top"], equal_nan=True), name
    assert np.isnan(S["flat_orient"]).any()             # the 0 * inf path is covered
    assert len(S["flat_points"]) < len(S["noise_points"])


@pytest.mark.parametrize("order", ["fused", "operator"])
def test_c_oracle_against_literal(goldens, default_filters, c_oracle, order):
    """Both canonical orders of the bit-defined oracle against the reference's own code run on the shim."""
    S = goldens["stack"]
    for name in ("noise", "noise_odd", "natural", "flat"):
        pyr = S[name + "_pyramid"]
        r = c_oracle.line_end_stack(pyr, default_filters, order=order)
        for key in ("rgc", "rgby", "orient", "line_end", "padded", "gray"):
            a, b = r[key], S[name + "_" + key]
            assert np.array_equal(np.isnan(a), np.isnan(b)), (name, key)
            assert np.nanmax(np.abs(a - b)) <= 1e-5 * np.nanmax(np.abs(b)), (name, key)
        assert np.array_equal(r["points"], S[name + "_points"]), name


def test_c_oracle_fused_order_uses_the_structures(default_filters, c_oracle):
    """The reference's generators produce the structures the fused order relies on; perturbed filters fall back to the
    operator order bit for bit."""
    L = c_oracle.lib()
    W = {k: np.ascontiguousarray(v, np.float32) for k, v in default_filters.items()}
    assert L.so_depthwise3(W["rgc"]) and L.so_rgby_shared(W["rgby"]) and L.so_stripe_sym180(W["stripe"])
    assert L.so_end_ownoth(W["end"])
    pyr = np.random.RandomState(5).rand(2, 20, 28, 3).astype(np.float32) * 255
    fused = c_oracle.line_end_stack(pyr, default_filters)
    oper = c_oracle.line_end_stack(pyr, default_filters, order="operator")
    assert not np.array_equal(fused["orient"], oper["orient"])           # a different rounding order ...
    assert np.abs(fused["orient"] - oper["orient"]).max() <= 1e-5 * oper["orient"].max()   # ... of the same math
    broken = dict(default_filters)
    broken["rgby"] = np.array(default_filters["rgby"], copy=True)
    broken["rgby"][0, 0, 1, 2] *= 1.5
    broken["end"] = np.array(default_filters["end"], copy=True)
    broken["end"][2, 1, 0, 1] += 0.01
    Wb = {k: np.ascontiguousarray(v, np.float32) for k, v in broken.items()}
    assert not L.so_rgby_shared(Wb["rgby"]) and not L.so_end_ownoth(Wb["end"])
    a, b = c_oracle.line_end_stack(pyr, broken), c_oracle.line_end_stack(pyr, broken, order="operator")
    for key in ("orient", "padded", "gray"):
        assert np.array_equal(a[key], b[key], equal_nan=True), key


def test_c_oracle_pyramid_against_goldens(goldens, c_oracle):
    P = goldens["pyramid"]
    for name in [k[:-8] for k in P.files if k.endswith("_pyramid")]:
        img = P[name + "_image"]
        cw, ch, sc = P[name + "_params"]
        got = c_oracle.from_image(img, img.shape[2], [int(cw), int(ch)], float(sc))
        want = P[name + "_pyramid"]
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), name
        assert np.array_equal(got == 0, want == 0) or name == "ragged"    # same undefined-tail rows/cols are zero


def test_c_oracle_on_a_real_shape(c_oracle, default_filters):
    frame = structured_frame(3, 240, 320)
    pyr_c = c_oracle.from_image(frame, 3, (96, 64), 1.5)
    pyr_l = lit.from_image(frame, 3, (96, 64), 1.5)
    assert pyr_c.shape == pyr_l.shape == (4, 64, 96, 3)
    assert np.abs(pyr_c - pyr_l).max() <= 1e-5 * 255
    rc, rl = c_oracle.line_end_stack(pyr_l, default_filters), lit.line_end_stack(pyr_l, default_filters)
    assert np.isnan(rl["orient"]).any()
    assert np.array_equal(np.isnan(rc["orient"]), np.isnan(rl["orient"]))
    assert np.nanmax(np.abs(rc["padded"] - rl["padded"])) <= 1e-5 * 255
    assert np.array_equal(rc["points"], rl["points"])


def test_canon_pow(c_oracle):
    xs = np.float32(np.random.RandomState(0).rand(4000))
    got = np.array([c_oracle.canon_pow(x, 0.1) for x in xs], np.float32)
    want = np.power(xs.astype(np.float64), float(np.float32(0.1))).astype(np.float32)
    assert (got == want).mean() > 0.999 and np.max(np.abs(got - want) / want) < 2e-7
    assert c_oracle.canon_pow(0, .1) == 0 and c_oracle.canon_pow(1, .1) == 1 and c_oracle.canon_pow(.5, 0) == 1
    assert np.isnan(c_oracle.canon_pow(-1, .1)) and np.isnan(c_oracle.canon_pow(np.nan, .1))
    assert c_oracle.canon_pow(1e-45, .1) == pytest.approx(1.4012984643e-45 ** float(np.float32(.1)), rel=1e-6)
    assert c_oracle.canon_pow(0.25, .5) == 0.5 and c_oracle.canon_pow(0, -1.0) == np.inf


def test_emit_edge_cases(c_oracle):
    g = np.zeros((2, 8, 12, 1), np.float32)
    g[1, 3, 4, 0] = 5
    pts = lit.max_value_indices_region(np.repeat(g, 3, -1), [1, 4, 6, 3], g)
    assert (pts[:, 0] == 0).sum() == 96                     # all-zero level: every pixel qualifies
    # one bright pixel: it wins every window containing it; quadrants whose window max is 5 emit nothing else
    assert [tuple(r) for r in pts[pts[:, 0] == 1]] == [(1, 3, 4, 0)]
    c_pts, count = c_oracle.max_value_indices_region(g, (4, 6))
    assert count == len(pts) and np.array_equal(c_pts, pts)
    capped, count2 = c_oracle.max_value_indices_region(g, (4, 6), capacity=10)
    assert count2 == count and np.array_equal(capped, pts[:10])
    g[0, 0, 0, 0] = np.nan                                  # NaN poisons every window that contains it
    pts_nan = lit.max_value_indices_region(np.repeat(g, 3, -1), [1, 4, 6, 3], g)
    assert np.array_equal(c_oracle.max_value_indices_region(g, (4, 6))[0], pts_nan)
    assert (pts_nan[:, 0] == 0).sum() < 96


def test_synthetic_frames_are_deterministic():
    a, b = synthetic_frame(2, 7, 48, 64), synthetic_frame(2, 7, 48, 64)
    assert a.dtype == np.uint8 and np.array_equal(a, b) and not np.array_equal(a, synthetic_frame(2, 8, 48, 64))


# ---- centroids / boosting / display tensors --------------------------------------------------------

def test_display_oracles_equal_reference_centroid_and_boosting_code(goldens, c_oracle):
    """tests/golden/display.npz was produced by the reference's util/centroids.py + util/energy/boosting.py (on the TF-1
    shim) for three consecutive frames; both oracles reproduce it from the same gray / padded tensors: the literal one
    bit for bit, the bit-defined C one to the north-star tolerance with identical fired cells."""
    import os
    from conftest import GOLDEN
    S, D = goldens["stack"], np.load(os.path.join(GOLDEN, "display.npz"))
    for name in ("noise", "natural", "flat"):
        gray, padded, orient = S[name + "_gray"], S[name + "_padded"], S[name + "_orient"]
        n, h, w, _ = gray.shape
        e_c = np.full((n, -(-h // 3), -(-w // 3), 1), 8, np.float32)
        e_l = e_c.copy()
        for step in range(3):
            outs_c, e_c = c_oracle.display_tensors(orient, padded, gray, e_c)
            outs_l, e_l = lit.display_tensors(orient, padded, gray, e_l)
            for key, i in (("centroids", 1), ("centroids2", 2), ("fired", 3), ("update", 4)):
                gold = D["%s_step%d_%s" % (name, step, key)]
                assert np.array_equal(outs_l[i], gold, equal_nan=True), (name, step, key)
                assert np.array_equal(np.isnan(outs_c[i]), np.isnan(gold))
                scale = 255.0 * max(h, w) if key.startswith("centroids") else np.nanmax(np.abs(gold))
                assert np.nanmax(np.abs(outs_c[i] - gold)) <= 1e-5 * scale, (name, step, key)
            assert np.array_equal(outs_c[3] > 0, D["%s_step%d_fired" % (name, step)] > 0)
            assert np.array_equal(e_l, D["%s_step%d_energy" % (name, step)])
            assert np.abs(e_c - e_l).max() <= 1e-6
        assert (e_l < 1).any() and (e_l == 1).any()      # some cells fired and are exhausted, the rest recovered fully


def test_index_tensor_reference_kat():
    """The reference's tests/test_index_tensor.py:10-24: element [y, x] of the index tensor is (x, y)."""
    from pysilent_b200.util import index_tensor
    want = [[[0, 0], [1, 0]], [[0, 1], [1, 1]]]
    assert index_tensor.from_shape([4, 2, 2, 3]).tolist() == want
    assert index_tensor.from_tensor(np.ones((4, 2, 2, 3))).tolist() == want
    assert np.array_equal(lit.index_tensor(2, 2), np.asarray(want, np.float32))


def test_recovery_selection_like_reference():
    from pysilent_b200.util.energy import recovery_mode
    assert (recovery_mode(False, True), recovery_mode(True, False), recovery_mode(True, True)) == (1, 2, 3)
    with pytest.raises(ValueError):
        recovery_mode(False, False)


# ---- the TF-1 stand-in pinned by an independent implementation (torch on CPU) ---------------------------------------------

def _shim():
    from oracle import tf1_shim
    return tf1_shim


@pytest.mark.parametrize("h,w", [(9, 12), (10, 13), (16, 16), (1, 5)])
@pytest.mark.parametrize("k", [3, 7, 2, 4])
def test_shim_conv2d_same_matches_torch(h, w, k):
    """tf.nn.conv2d(SAME, stride 1) of the shim vs torch.nn.functional.conv2d: cross-correlation, HWIO filters, zero
    padding with the extra pad row / column at the END for even kernels (TF's pad_before = pad_total // 2)."""
    torch = pytest.importorskip("torch")
    F = torch.nn.functional
    rs = np.random.RandomState(100 * k + h)
    x = rs.randn(2, h, w, 3).astype(np.float32)
    wt = rs.randn(k, k, 3, 4).astype(np.float32)
    got = _shim().conv2d(input=x, filter=wt, strides=[1, 1, 1, 1], padding="SAME").numpy()
    total = k - 1
    xp = F.pad(torch.from_numpy(x).permute(0, 3, 1, 2).double(), (total // 2, total - total // 2, total // 2, total - total // 2))
    want = F.conv2d(xp, torch.from_numpy(wt).permute(3, 2, 0, 1).double()).permute(0, 2, 3, 1).numpy()
    assert got.shape == want.shape == (2, h, w, 4)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    if k % 2:   # torch's own 'same' padding agrees with TF's for odd kernels
        same = F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2).double(), torch.from_numpy(wt).permute(3, 2, 0, 1).double(),
                        padding="same").permute(0, 2, 3, 1).numpy()
        assert np.abs(got - same).max() <= 1e-5 * np.abs(same).max()


@pytest.mark.parametrize("h,w,rh,rw", [(8, 12, 4, 6), (9, 13, 4, 6), (192, 288, 96, 144), (7, 5, 3, 2), (6, 6, 6, 6)])
def test_shim_max_pool_whole_level_window_matches_torch(h, w, rh, rw):
    """max_value_indices_region's pooling (top_value_points.py:39): ksize = the whole level, stride = the region, SAME.
    torch: explicit -inf padding with TF's pad_before = pad_total // 2, then an unpadded max_pool2d."""
    torch = pytest.importorskip("torch")
    F = torch.nn.functional
    x = np.random.RandomState(h * w).randn(2, h, w, 1).astype(np.float32)
    got = _shim().max_pool(x, [1, h, w, 1], strides=[1, rh, rw, 1], padding="SAME").numpy()
    oh, ow = -(-h // rh), -(-w // rw)
    th, tw = max((oh - 1) * rh + h - h, 0), max((ow - 1) * rw + w - w, 0)
    xp = F.pad(torch.from_numpy(x).permute(0, 3, 1, 2), (tw // 2, tw - tw // 2, th // 2, th - th // 2), value=float("-inf"))
    want = F.max_pool2d(xp, kernel_size=(h, w), stride=(rh, rw)).permute(0, 2, 3, 1).numpy()
    assert got.shape == want.shape == (2, oh, ow, 1)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("h,w,oh,ow", [(2, 2, 192, 288), (2, 2, 9, 13), (3, 4, 10, 10), (64, 96, 192, 288), (5, 7, 5, 7)])
def test_shim_resize_nearest_matches_torch(h, w, oh, ow):
    """tf.image.resize_images(NEAREST, align_corners=False): src = min(floor(dst * in / out), in - 1) -- the rule
    torch.nn.functional.interpolate(mode='nearest') implements."""
    torch = pytest.importorskip("torch")
    x = np.random.RandomState(oh + ow).randn(2, h, w, 3).astype(np.float32)
    T = _shim()
    got = T.resize_images(x, [oh, ow], method=T._ResizeMethod.NEAREST_NEIGHBOR).numpy()
    want = torch.nn.functional.interpolate(torch.from_numpy(x).permute(0, 3, 1, 2), size=(oh, ow), mode="nearest")
    assert np.array_equal(got, want.permute(0, 2, 3, 1).numpy())


def test_scipy_pyramid_is_the_literal_pyramid():
    """bench.py's "reference-python" CPU arm calls scipy.ndimage.zoom exactly like from_image.py:55-59; the literal
    oracle's own spline code reproduces it to float32 rounding."""
    pytest.importorskip("scipy")
    frame = synthetic_frame(1, 0, 240, 320)
    a = lit.from_image_scipy(frame, 3, (96, 64), 1.5)
    b = lit.from_image(frame, 3, (96, 64), 1.5)
    assert a.shape == b.shape and np.abs(a - b).max() <= 1e-5 * 255
