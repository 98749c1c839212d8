"""Simplex vertex coordinates: the ``n + 1`` unit vectors used as filter orientations.

Reference: ``slam_recognition/util/orientation/simplex_coordinates.py:4-34``; the 2-D and 3-D tables are pinned by
``tests/test_simplex_coordinates.py:9-22``.
"""
import numpy as np


def simplex_coordinates(n):
    """Rows are the ``n + 1`` vertices of a regular simplex inscribed in the unit sphere of ``R^n``.

    Built column by column. The correction term for column ``k`` sums the SQUARES of the pivot row's earlier entries
    (the reference multiplies ``c1 * c1``, ``simplex_coordinates.py:21``), which is part of results parity.
    """
    vertices = np.zeros([n + 1, n])
    for k in range(n):
        head = vertices[k, :k]
        squares = 0
        for c in head:
            squares = squares + c ** 2
        vertices[k, k] = np.sqrt(1.0 - squares)
        for row in range(k + 1, n + 1):
            carried = 0
            for c in head:
                carried = carried + c * c
            vertices[row, k] = (-1.0 / float(n) - carried) / vertices[k, k]
    return vertices


def axis_coordinates(n):
    return np.eye(n, n)


def above_axis_simplex_coordinates(n, axis=0):
    """Simplex vertices folded onto the non-negative side of ``axis``."""
    s = simplex_coordinates(n)
    s[:, axis] = np.absolute(s[:, axis])
    return s
