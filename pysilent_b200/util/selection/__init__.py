"""Mirrors ``slam_recognition/util/selection/__init__.py:1-2`` (hot-path subset)."""
from .isolate_rectangle import pad_inwards
from .top_value_points import top_value_points, max_value_indices_region

__all__ = ["pad_inwards", "top_value_points", "max_value_indices_region"]
