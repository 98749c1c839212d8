"""Host weight generators: the reference's own known-answer tests (ported) + bit-exact comparison with outputs of the
reference's generators executed in the build container (tests/golden/generators.npz, made by make_golden.py)."""
import importlib
import time

import numpy as np
import numpy.testing as npt
import pytest

import pysilent_b200 as sb
import pysilent_b200.constant_convolutions as cc
from pysilent_b200.util.normalize import normalize_tensor_positive_negative
from pysilent_b200.util.orientation import simplex_coordinates, above_axis_simplex_coordinates, axis_coordinates
from pysilent_b200.util import attractor


# ---- reference tests/test_center_surround_tensors.py:8-74 -------------------------------------------------------------------
def test_surround_center_1d_exact():
    t = sb.center_surround_tensor(1, [0, 1, 0], [1, 0, 0], [0, 0, 1], [1, 0, 0])
    npt.assert_array_equal(t, [[[0, 0, 0], [0, 0, 0], [1, 0, 0]], [[0, 0, 0], [2, 0, 0], [0, 0, 0]],
                               [[0, 0, 0], [0, 0, 0], [1, 0, 0]]])


def test_surround_center_2d_values_and_layout():
    t = sb.center_surround_tensor(2, [0, 1, 0], [1, 0, 0], [0, 0, 1], [1, 0, 0])
    assert t.shape == (3, 3, 3, 3)
    want = np.zeros((3, 3, 3, 3))
    for ky in range(3):
        for kx in range(3):
            d = abs(ky - 1) + abs(kx - 1)
            if d:
                want[ky, kx, 2, 0] = {1: 1.0, 2: 0.70710678}[d]     # surround_in channel 2 -> surround_out channel 0
    want[1, 1, 1, 0] = 6.82842712                                    # center_in channel 1 -> center_out channel 0
    npt.assert_array_almost_equal(t, want)


def test_surround_center_time_guard():
    for i in range(1, 11):
        t1 = time.time()
        sb.center_surround_tensor(i, [0, 1, 0], [1, 0, 0], [0, 0, 1], [1, 0, 0])
        assert time.time() - t1 <= 1.0, "%d-dimensional center-surround too slow" % i


# ---- reference tests/test_normalize_center_surround.py:9-26 ------------------------------------------------------------------
def test_normalize_basic_and_in_place():
    t1 = np.squeeze(sb.center_surround_tensor(1, [1], [1], [1], [-1]))
    npt.assert_array_almost_equal(t1, [-1, 2, -1])
    npt.assert_array_almost_equal(normalize_tensor_positive_negative(t1), [-.5, 1, -.5])
    t2 = np.squeeze(sb.center_surround_tensor(2, [1], [1], [1], [-1]))
    npt.assert_array_almost_equal(t2, [[-0.70710678, -1., -0.70710678], [-1., 6.82842712, -1.],
                                       [-0.70710678, -1., -0.70710678]])
    want = [[-0.10355339, -0.14644661, -0.10355339], [-0.14644661, 1., -0.14644661], [-0.10355339, -0.14644661, -0.10355339]]
    out = normalize_tensor_positive_negative(t2)
    npt.assert_array_almost_equal(out, want)
    npt.assert_array_almost_equal(t2, want)      # mutated in place (reference test line 24)
    assert out is t2


# ---- reference tests/test_simplex_coordinates.py:9-22 ------------------------------------------------------------------------
def test_simplex_tables():
    npt.assert_array_almost_equal(simplex_coordinates(2), [[1., 0.], [-0.5, 0.8660254], [-0.5, -0.8660254]])
    npt.assert_array_almost_equal(simplex_coordinates(3), [[1., 0., 0.], [-0.33333333, 0.94280904, 0.],
                                                           [-0.33333333, -0.47140452, 0.81649658],
                                                           [-0.33333333, -0.47140452, -0.81649658]])
    npt.assert_array_equal(axis_coordinates(3), np.eye(3))


# ---- bit-exact against the reference's generators ------------------------------------------------------------------------------
def _ours():
    rgc = importlib.import_module("pysilent_b200.constant_convolutions.center_surround.rgc")
    et = importlib.import_module("pysilent_b200.constant_convolutions.edge_orientation_detector.edge_tensor")
    vecs = [np.array([np.cos(k * np.pi / 4), np.sin(k * np.pi / 4)]) for k in range(8)]
    eye, spread = np.eye(8), [1, 1, 1, 0, 0, 0, 0, 0]
    return {
        "cs_1d_test": lambda: cc.center_surround_tensor(1, [0, 1, 0], [1, 0, 0], [0, 0, 1], [1, 0, 0]),
        "cs_2d_test": lambda: cc.center_surround_tensor(2, [0, 1, 0], [1, 0, 0], [0, 0, 1], [1, 0, 0]),
        "cs_3d": lambda: cc.center_surround_tensor(3, [1, .5], [1, -2], [.25, 1], [-1, 3]),
        "midget_rgc_2": lambda: cc.midget_rgc(2), "midget_rgc_1": lambda: cc.midget_rgc(1),
        "midget_rgc_full_2": lambda: rgc.midget_rgc_full(2),
        "rgby_2": lambda: cc.rgby(2), "rgby_3_2": lambda: cc.rgby_3(2), "rgby_3_3": lambda: cc.rgby_3(3),
        "rgb_2d_stripe": cc.rgb_2d_stripe_tensors,
        "rgb_2d_stripe_in": lambda: cc.rgb_2d_stripe_tensors(in_channel=(1, .5, 0)),
        "stripe_3d": lambda: cc.stripe_tensor([0.0, 0.6, 0.8], [1, 0], [2, 1], [1, 1], [-1, .5]),
        "rgb_2d_edge": cc.rgb_2d_edge_tensors, "rgb_2d_edge_time_diff": cc.rgb_2d_edge_tensors_time_diff,
        "rgb_2d_end_7x7": et.rgb_2d_end_tensors, "rgb_2d_end": cc.rgb_2d_end_tensors,
        "blur_2_7": lambda: cc.blur_tensor(2, lengths=7), "blur_2_default": lambda: cc.blur_tensor(2),
        "blur_3_list": lambda: cc.blur_tensor(3, lengths=[3, 5, 3], channels_in=2, channels_out=1),
        "stripe_8": lambda: sum(cc.stripe_tensor(v, spread, list(eye[k] * 4), spread, list(-eye[k] * 4))
                                for k, v in enumerate(vecs)),
        "end_8": lambda: sum(cc.end_tensor(3 * v, list(eye[k]), list(.25 * eye[k]), list(eye[k]), list(.5 * eye[k]))
                             for k, v in enumerate(vecs)),
        "blur_8": lambda: cc.blur_tensor(2, 7, channels_in=8, channels_out=8),
        "simplex_2": lambda: simplex_coordinates(2), "simplex_3": lambda: simplex_coordinates(3),
        "simplex_5": lambda: simplex_coordinates(5), "above_axis_simplex_3": lambda: above_axis_simplex_coordinates(3),
    }


@pytest.mark.parametrize("name", sorted(_ours()))
def test_generator_bit_exact_vs_reference(name, goldens):
    got = _ours()[name]()
    want = goldens["generators"][name]
    assert got.dtype == np.float64 and got.shape == want.shape
    assert np.array_equal(got, want), "max |d| = %g" % np.abs(got - want).max()


def test_normalize_bit_exact_vs_reference(goldens):
    G = goldens["generators"]
    t = G["norm_in"].copy()
    assert np.array_equal(normalize_tensor_positive_negative(t), G["norm_out"])
    t2 = G["norm_rand_in"].copy()
    assert np.array_equal(normalize_tensor_positive_negative(t2, 3.0, 0.5), G["norm_rand_out"])


def test_package_exports_3x3_end_tensors():
    # constant_convolutions/__init__.py:5 overrides the 7x7 rgb_2d_end_tensors with the 3x3 one
    assert cc.rgb_2d_end_tensors().shape == (3, 3, 3, 3)
    assert cc.contrast_adjust() == [[[1, -0.5, -0.5], [-0.5, 1, -0.5], [-0.5, -0.5, 1]]]


def test_structure_the_fused_kernel_relies_on():
    rgc, stripe, blur, end = cc.midget_rgc(2), cc.rgb_2d_stripe_tensors(), cc.blur_tensor(2, lengths=7), cc.rgb_2d_end_tensors()
    off_diag = ~np.eye(3, dtype=bool)
    assert (rgc[:, :, off_diag] == 0).all()                                  # depthwise
    assert (stripe == stripe[:, :, :1, :]).all()                             # identical input slices
    assert (blur == blur[:, :, :1, :1]).all() and np.array_equal(blur[..., 0, 0], blur[..., 0, 0].T)
    assert (end != 0).all()                                                  # dense: NaN spreads like a dense conv
    assert abs(rgc[rgc > 0].sum() - 4) < 1e-12 and abs(rgc[rgc < 0].sum() + 2) < 1e-12


def test_error_behaviour_matches_reference():
    with pytest.raises(AssertionError):
        cc.center_surround_tensor(0, [1], [1], [1], [1])
    with pytest.raises(ValueError):      # non-square channel maps: numpy broadcast error in the reference
        cc.stripe_tensor([1.0, 0.0], [1, 1, 1], [1, 0], [1, 1, 1], [1, 0])
    with pytest.raises(ValueError):
        cc.end_tensor([3.0, 0.0], [1, 1], [1, 0, 0], [1, 1], [1, 0, 0])
    with pytest.raises(AssertionError):
        cc.blur_tensor(0)


def test_attractors():
    f = attractor.euclidian_attractor_function_generator(2)
    assert f(0) == 1.0 and f(1) == pytest.approx(2 / 3 - 1) and f(-1) == -f(1)
    g = attractor.linear_attractor_function_generator()
    assert g(0) == 1.0 and g(1) == -1.0 and g(-.5) == 0.0
    assert attractor.piecewise_attractor_function(.4) == 1.0 and attractor.piecewise_attractor_function(.5) == -.5
    assert np.isfinite(attractor.log_attractor_function(.3))
