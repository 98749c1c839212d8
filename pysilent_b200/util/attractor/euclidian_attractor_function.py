"""Inverse-power ("euclidian") attractor profile.

Reference: ``slam_recognition/util/attractor/euclidian_attractor_function.py:8-30``.
``f(x) = (p + n) / (2 x^(d-1) + 1)^(d-1) - n`` for ``x >= 0`` and ``f(-x) = -f(x)``, where ``d`` is the number of
dimensions, ``p`` the value at distance zero and ``-n`` the value at infinity.
"""


def euclidian_attractor_function_generator(n, max_positive=1.0, max_negative=1.0):
    """Return the attractor ``f`` for ``n``-dimensional space (odd-symmetric about 0)."""
    span = max_positive + max_negative
    power = n - 1

    def n_dimensional_euclid_function(x):
        sign = 1.0
        if not x >= 0:
            sign, x = -1.0, -x
        falloff = span / (((2 * x ** power) + 1) ** power) - max_negative
        return falloff if sign > 0 else -falloff

    return n_dimensional_euclid_function
