"""7^n thick-edge oriented filters (signed-distance version of the stripe filters).

Reference: ``slam_recognition/constant_convolutions/edge_orientation_detector/edge_tensor.py:21-158``. The tap offset is
``t - 1`` on the 7-wide grid (``:56``), i.e. the facet passes through index 1, not the middle: part of results parity.
"""
import numpy as np

from ...util.attractor import euclidian_attractor_function_generator
from ...util.normalize import normalize_tensor_positive_negative
from ...util.orientation import simplex_coordinates
from .._fill import fill_center_surround
from .stripe_tensor import _projected_profile


def edge_tensor(normal_vector, center_in, center_out, surround_in, surround_out,
                attractor_function=euclidian_attractor_function_generator):
    """One oriented thick-edge filter, float64 ``[7]*ndim + [C, C]``."""
    assert len(normal_vector) >= 1
    ndim = len(normal_vector)
    profile_of = attractor_function(ndim, max_positive=0.0, max_negative=-1.0)
    if isinstance(normal_vector, list):
        normal_vector = np.asarray(normal_vector)
    profile = _projected_profile(normal_vector, 7, 1, profile_of, signed=True)
    normalize_tensor_positive_negative(profile)
    return fill_center_surround(profile, center_in, center_out, surround_in, surround_out, zero_is_center=True)


def simplex_edge_tensors(dimensions, centers_in, centers_out, surrounds_in, surrounds_out,
                         attractor_function=euclidian_attractor_function_generator, flip=None):
    """One edge filter per simplex vertex direction; ``flip`` reverses the given axis of the vertex table first."""
    simplex = simplex_coordinates(dimensions)
    if flip is not None:
        simplex = np.flip(simplex, flip)
    return [edge_tensor(v, ci, co, si, so, attractor_function)
            for v, ci, co, si, so in zip(simplex, centers_in, centers_out, surrounds_in, surrounds_out)]


def _opponent(x, diag, off):
    return [[diag * x if i == j else off * x for j in range(3)] for i in range(3)]


def rgb_2d_edge_tensors(in_channel=(1, 1, 1)):
    """2-D thick-edge bank summed into one ``[7, 7, 3, 3]`` filter."""
    x = 2
    return sum(simplex_edge_tensors(2, [in_channel] * 3, _opponent(x, 2, -.5), [in_channel] * 3,
                                    _opponent(x, -2, .5)))


def rgb_2d_edge_tensors_time_diff(in_channel=(1, 1, 1), surround_in_channel=(-1, -1, -1)):
    """Variant whose surround reads a negated (time-differenced) input."""
    x = 2
    return sum(simplex_edge_tensors(2, [in_channel] * 3, _opponent(x, 2, -1), [surround_in_channel] * 3,
                                    _opponent(x, -2, 1)))


def rgb_2d_end_tensors(north_input_channel=(1, -.5, -.5), southwest_input_channel=(-.5, 1, -.5),
                       southeast_input_channel=(-.5, -.5, 1)):
    """7x7 line-end variant. The package-level name ``rgb_2d_end_tensors`` is the 3x3 one from
    ``oriented_end_detector`` (reference ``constant_convolutions/__init__.py:5``); the arguments are unused, as in the
    reference (``edge_tensor.py:139-158``)."""
    x = 2
    inputs = [(0, 1, 1), (1, 0, 1), (1, 1, 0)]  # south, north-east, north-west
    colours = _opponent(x, 2, -1)
    return sum(simplex_edge_tensors(2, inputs, colours, inputs, colours, flip=1))
