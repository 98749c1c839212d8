"""Mirrors ``slam_recognition/util/normalize/__init__.py:1``."""
from .normalize_center_surround import normalize_tensor_positive_negative

__all__ = ["normalize_tensor_positive_negative"]
