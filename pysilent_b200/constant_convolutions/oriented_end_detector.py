"""Line-end (end-stopped) filters: taps are weighted by their angle to an end vector.

Reference: ``slam_recognition/constant_convolutions/oriented_end_detector.py:13-99``. Tap offsets are measured from
``size / 2`` (true division: 1.5 on a 3-grid), so the filter is deliberately off-centre: part of results parity.
"""
import math

import numpy as np

from ..util.attractor import linear_attractor_function_generator
from ..util.normalize import normalize_tensor_positive_negative
from ..util.orientation import simplex_coordinates
from ._fill import fill_center_surround


def end_tensor(end_vector, center_in, center_out, surround_in, surround_out,
               attractor_function=linear_attractor_function_generator, size=3):
    """One line-end filter, float64 ``[size]*ndim + [C, C]``.

    Profile: ``attractor(acos(cos(angle(tap - size/2, end_vector))) - pi/2)``, then normalised so positives and
    negatives each sum to 1; ``z >= 0`` taps take the centre colours (``:51``).
    """
    ndim = len(end_vector)
    assert ndim >= 1
    if not isinstance(end_vector, np.ndarray):
        end_vector = np.asarray(end_vector)
    profile_of = attractor_function()
    origin = np.asarray([size / 2 for _ in range(ndim)])
    end_norm = np.linalg.norm(end_vector)
    profile = np.ndarray(shape=[size] * ndim)
    flat = profile.reshape(-1)
    for j, tap in enumerate(np.indices((size,) * ndim).reshape(ndim, -1).T):
        offset = np.asarray(tuple(tap.tolist())) - origin
        cosine = np.dot(offset, end_vector) / (np.linalg.norm(offset) * end_norm)
        angle_dist = (math.acos(cosine) - math.pi / 2.0) / math.pi
        flat[j] = profile_of(angle_dist * math.pi)
    normalize_tensor_positive_negative(profile)
    return fill_center_surround(profile, center_in, center_out, surround_in, surround_out, zero_is_center=True)


def simplex_end_tensors(dimension, centers_in, centers_out, surrounds_in, surrounds_out,
                        attractor_function=linear_attractor_function_generator, flip=True):
    """One end filter per simplex vertex (scaled by 3). ``flip=True`` reverses axis 1 of the vertex table, i.e. swaps
    (x, y) in 2-D (reference ``:64,:78``)."""
    simplex = simplex_coordinates(dimension)
    simplex *= 3
    if flip is not None:
        simplex = np.flip(simplex, flip)
    return [end_tensor(v, ci, co, si, so, attractor_function)
            for v, ci, co, si, so in zip(simplex, centers_in, centers_out, surrounds_in, surrounds_out)]


def rgb_2d_end_tensors(north_input_channel=(1, 0, 0), southwest_input_channel=(0, 1, 0),
                       southeast_input_channel=(0, 0, 1)):
    """The 2-D line-end bank summed into one dense ``[3, 3, 3, 3]`` filter."""
    x, xx, y, yy = 0.5 / 2, -0.25 / 2, 1.0 / 2, 1.0 / 2
    inputs = [north_input_channel, southwest_input_channel, southeast_input_channel]
    lit = [[x if i == j else -xx for j in range(3)] for i in range(3)]
    dark = [[y if i == j else -yy for j in range(3)] for i in range(3)]
    return sum(simplex_end_tensors(2, inputs, lit, inputs, dark))
