"""Scalar distance -> weight profiles used by the weight generators.

Mirrors ``slam_recognition/util/attractor/__init__.py:1-4``.
"""
from .log_attractor_function import log_attractor_function
from .piecewise_attractor_function import piecewise_attractor_function
from .euclidian_attractor_function import euclidian_attractor_function_generator
from .linear_attractor_function import linear_attractor_function_generator

__all__ = ["log_attractor_function", "piecewise_attractor_function",
           "euclidian_attractor_function_generator", "linear_attractor_function_generator"]
