"""Colour-opponent center-surround filters (red/green, blue/yellow).

Reference: ``slam_recognition/constant_convolutions/center_surround/rgby.py:14-56``. Channel 0 is blue because frames
arrive from OpenCV in BGR order.
"""
from ...util.normalize import normalize_tensor_positive_negative
from .center_surround_tensor import center_surround_tensor


def _opponent_sum(n, terms):
    out = None
    for center_in, center_out, surround_in, surround_out in terms:
        t = center_surround_tensor(n, center_in=center_in, center_out=center_out,
                                   surround_in=surround_in, surround_out=surround_out)
        out = t if out is None else out + t
    return out


def rgby(n):
    """3 -> 4 channel opponent filter (un-normalised), ``[3]*n + [3, 4]``."""
    d = 1.
    return _opponent_sum(n, [
        ([0, 0, d], [0, 0, d, 0], [0, d, 0], [0, 0, -d, 0]),              # red centre, green surround
        ([d, 0, 0], [d, 0, 0, 0], [0, d / 2, d / 2], [-d, 0, 0, 0]),      # blue centre, yellow surround
        ([0, d, 0], [0, d, 0, 0], [0, 0, d], [0, -d, 0, 0]),              # green centre, red surround
        ([0, d / 2, d / 2], [0, 0, 0, d], [d, 0, 0], [0, 0, 0, -d]),      # yellow centre, blue surround
    ])


def rgby_3(n):
    """3 -> 3 channel opponent filter (yellow folded back onto green+red); positives sum to 4, negatives to -2."""
    d = 1. / 3
    out = _opponent_sum(n, [
        ([0, 0, d], [0, 0, d], [0, d, 0], [0, 0, -d]),
        ([d, 0, 0], [d, 0, 0], [0, d / 2, d / 2], [-d, 0, 0]),
        ([0, d, 0], [0, d, 0], [0, 0, d], [0, -d, 0]),
        ([0, d / 2, d / 2], [0, d / 2, d / 2], [d, 0, 0], [0, -d / 2, -d / 2]),
    ])
    return normalize_tensor_positive_negative(out, 4.0, 2.0)
