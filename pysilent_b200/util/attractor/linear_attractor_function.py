"""Tent ("linear") attractor profile.

Reference: ``slam_recognition/util/attractor/linear_attractor_function.py:8-28``: ``f(x) = p - (p + n) |x|``.
"""


def linear_attractor_function_generator(max_positive=1.0, max_negative=1.0):
    """Return a tent function that is ``max_positive`` at 0 and ``-max_negative`` at ``|x| = 1``."""
    slope = max_negative + max_positive

    def linear_attractor_function(x):
        if x >= 0:
            return max_positive - slope * x
        return max_positive + slope * x

    return linear_attractor_function
