"""Mirrors ``slam_recognition/util/orientation/__init__.py:1``."""
from .simplex_coordinates import simplex_coordinates, axis_coordinates, above_axis_simplex_coordinates

__all__ = ["simplex_coordinates", "axis_coordinates", "above_axis_simplex_coordinates"]
