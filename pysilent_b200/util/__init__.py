"""Mirrors ``slam_recognition/util/__init__.py:1-3`` for the hot-path subset."""
from . import attractor, color, normalize, orientation, regulator, selection, zoom  # noqa: F401
from . import apply_filter, get_dimensions  # noqa: F401
