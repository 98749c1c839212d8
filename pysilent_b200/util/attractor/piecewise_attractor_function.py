"""Step attractor profile (no caller on the hot path).

Reference: ``slam_recognition/util/attractor/piecewise_attractor_function.py:1-10``.
"""


def piecewise_attractor_function(x, max_positive=1.0, max_negative=0.5):
    """``max_positive`` below 0.5, ``-max_negative`` from 0.5 on."""
    return max_positive if x < 0.5 else -max_negative
