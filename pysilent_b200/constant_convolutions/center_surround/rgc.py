"""Midget retinal-ganglion-cell filters: per-channel on-centre / off-surround.

Reference: ``slam_recognition/constant_convolutions/center_surround/rgc.py:14-54``. ``midget_rgc(2)`` is depthwise
(only ``in == out`` slices are non-zero), which the fused CUDA stack exploits by skipping zero taps.
"""
from ...util.normalize import normalize_tensor_positive_negative
from .center_surround_tensor import center_surround_tensor


def _one_hot(k, value, n=3):
    return [value if j == k else 0 for j in range(n)]


def _per_channel_sum(n, d, center_sign, surround_sign):
    out = None
    for k in range(3):
        term = center_surround_tensor(n, center_in=_one_hot(k, d), center_out=_one_hot(k, center_sign * d),
                                      surround_in=_one_hot(k, d), surround_out=_one_hot(k, surround_sign * d))
        out = term if out is None else out + term
    return out


def midget_rgc(n):
    """Each channel: its own centre minus its own surround; positives sum to 4, negatives to -2."""
    return normalize_tensor_positive_negative(_per_channel_sum(n, 1., +1, -1), 4.0, 2.0)


def midget_rgc_full(n):
    """Sign-flipped variant that expects negative input for the surround."""
    return normalize_tensor_positive_negative(_per_channel_sum(n, 1. / 2, -1, +1), 4.0, 2.0)
