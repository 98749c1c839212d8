"""Logarithmic attractor profile (no caller on the hot path; kept for surface completeness).

Reference: ``slam_recognition/util/attractor/log_attractor_function.py:4-26``.
"""
import math


def log_attractor_function(x, max_positive=1.0, max_negative=0.5):
    """``-log2(x^2 + (x-1)^2 / 2^(p+x)) + log2((x-1)^2 + x^2 / 2^(n-1+x))``."""
    near = x ** 2 + ((x - 1) ** 2) / (2 ** (max_positive + x))
    far = (x - 1) ** 2 + (x ** 2) / (2 ** (max_negative - 1 + x))
    return -math.log(near, 2) + math.log(far, 2)
