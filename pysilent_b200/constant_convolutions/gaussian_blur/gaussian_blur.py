"""Radial blur filter ``1 / (2 r + 1)`` with every Cin x Cout slice identical.

Reference: ``slam_recognition/constant_convolutions/gaussian_blur/gaussian_blur.py:13-54``. Because all slices are
equal the regulator's 7x7x3x3 convolution collapses to one single-channel 7x7 pass over the channel sum.
"""
import math

import numpy as np

from ...util.attractor import euclidian_attractor_function_generator


def blur_tensor(n, lengths=3, channels_in=3, channels_out=3,
                attractor_function=euclidian_attractor_function_generator):
    """Return float64 ``lengths + [channels_in, channels_out]`` (``lengths`` an int or one int per dimension)."""
    assert n >= 1
    profile_of = attractor_function(n, max_negative=0)
    extent = [lengths] * n if isinstance(lengths, int) else [lengths[i] for i in range(n)]
    offsets = np.indices(extent).reshape(n, -1).T - np.asarray([int(e / 2) for e in extent])
    profile = np.empty(len(offsets))
    for j, off in enumerate(offsets.tolist()):
        sq = 0
        for d in off:
            sq = sq + d ** 2
        profile[j] = profile_of(math.sqrt(sq))
    gauss = np.ndarray(shape=extent + [channels_in, channels_out])
    gauss[...] = profile.reshape(extent)[..., np.newaxis, np.newaxis]
    return gauss
