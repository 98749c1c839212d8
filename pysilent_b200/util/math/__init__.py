"""Mirrors ``slam_recognition/util/math/__init__.py``."""
from .almost_equal import almost_equal, equality_distance

__all__ = ["almost_equal", "equality_distance"]
