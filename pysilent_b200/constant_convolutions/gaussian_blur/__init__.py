"""Mirrors ``slam_recognition/constant_convolutions/gaussian_blur/__init__.py:1``."""
from .gaussian_blur import blur_tensor

__all__ = ["blur_tensor"]
