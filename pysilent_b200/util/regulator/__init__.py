"""Mirrors ``slam_recognition/util/regulator/__init__.py``."""
from .gaussian_regulator_tensor import regulate_tensor

__all__ = ["regulate_tensor"]
