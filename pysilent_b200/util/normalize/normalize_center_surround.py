"""Weight-time normaliser (host, numpy float64). Not an image op.

Reference: ``slam_recognition/util/normalize/normalize_center_surround.py:5-24``; its in-place behaviour is pinned by
``tests/test_normalize_center_surround.py:24`` and relied on by the stripe/edge/end generators, which ignore the
return value.
"""
import numpy as np


def _running_total(values):
    # Plain left-to-right IEEE adds in memory order, as the reference's ``sum`` over ``np.nditer`` performs them.
    total = 0.0
    for v in values:
        total = total + v
    return total


def normalize_tensor_positive_negative(tensor, positive_value=1.0, negative_value=1.0, epsilon=1e-12):
    """Scale ``tensor`` IN PLACE so positives sum to ``positive_value`` and negatives to ``-negative_value``.

    Returns the same array object that was passed in.
    """
    flat = tensor.ravel(order="K").tolist()
    sum_pos = max(_running_total([v for v in flat if v > 0]), epsilon)
    sum_neg = max(_running_total([-v for v in flat if v < 0]), epsilon)
    # Positives are rescaled first; the negative test then sees the updated values (reference :19-23).
    np.multiply(tensor, positive_value / sum_pos, out=tensor, where=tensor > 0)
    np.multiply(tensor, negative_value / sum_neg, out=tensor, where=tensor < 0)
    return tensor
