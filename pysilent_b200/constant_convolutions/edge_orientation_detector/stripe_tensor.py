"""3^n thin-edge ("stripe") oriented filters, one per simplex direction.

Reference: ``slam_recognition/constant_convolutions/edge_orientation_detector/stripe_tensor.py:21-108``.
"""
import math

import numpy as np

from ...util.attractor import euclidian_attractor_function_generator
from ...util.normalize import normalize_tensor_positive_negative
from ...util.orientation import above_axis_simplex_coordinates
from .._fill import fill_center_surround


def _projected_profile(normal_vector, width, origin, profile_of, signed):
    """Profile value of every tap's (signed) distance from the facet through ``origin`` with the given normal."""
    ndim = len(normal_vector)
    taps = np.indices((width,) * ndim).reshape(ndim, -1).T.tolist()
    profile = np.ndarray(shape=[width] * ndim)
    flat = profile.reshape(-1)
    for j, tap in enumerate(taps):
        along = 0
        for t, nrm in zip(tap, normal_vector):
            along = along + (t - origin) * nrm
        projection = normal_vector * along
        if signed:
            acc = 0
            for p in projection:
                acc = acc + p * abs(p)
            dist = math.sqrt(abs(acc)) * (1 if acc >= 0 else -1)
        else:
            acc = 0
            for p in projection:
                acc = acc + p ** 2
            dist = math.sqrt(acc)
        flat[j] = profile_of(dist)
    return profile


def stripe_tensor(normal_vector, center_in, center_out, surround_in, surround_out,
                  attractor_function=euclidian_attractor_function_generator):
    """One oriented stripe filter, float64 ``[3]*ndim + [C, C]``.

    The tap profile is the attractor of the unsigned distance from the centre facet, normalised so positives and
    negatives each sum to 1; taps whose profile is exactly 0 stay 0 (strict ``> 0`` centre test, ``:66``).
    """
    assert len(normal_vector) >= 1
    ndim = len(normal_vector)
    profile_of = attractor_function(ndim)
    if isinstance(normal_vector, list):
        normal_vector = np.asarray(normal_vector)
    profile = _projected_profile(normal_vector, 3, 1, profile_of, signed=False)
    normalize_tensor_positive_negative(profile)
    return fill_center_surround(profile, center_in, center_out, surround_in, surround_out, zero_is_center=False)


def simplex_stripe_tensors(dimensions, centers_in, centers_out, surrounds_in, surrounds_out,
                           attractor_function=euclidian_attractor_function_generator):
    """One stripe filter per above-axis simplex direction (the minimum set covering all thin-edge orientations)."""
    return [stripe_tensor(v, ci, co, si, so, attractor_function)
            for v, ci, co, si, so in zip(above_axis_simplex_coordinates(dimensions), centers_in, centers_out,
                                         surrounds_in, surrounds_out)]


def rgb_2d_stripe_tensors(in_channel=(1, 1, 1)):
    """The 2-D stripe bank summed into one ``[3, 3, 3, 3]`` filter; orientation is coded in the output colour.

    All three input-channel slices are identical, so the convolution only depends on the channel sum.
    """
    x = 2
    lit = [[2 * x, -.5 * x, -.5 * x], [-.5 * x, 2 * x, -.5 * x], [-.5 * x, -.5 * x, 2 * x]]
    dark = [[-2 * x, .5 * x, .5 * x], [.25 * x, -2 * x, .5 * x], [.5 * x, .5 * x, -2 * x]]
    return sum(simplex_stripe_tensors(2, [in_channel] * 3, lit, [in_channel] * 3, dark))
