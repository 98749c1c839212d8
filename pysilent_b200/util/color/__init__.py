"""Mirrors ``slam_recognition/util/color/__init__.py`` (hot-path subset)."""
from .get_value import get_value_from_color

__all__ = ["get_value_from_color"]
