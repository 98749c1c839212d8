"""Mirrors ``slam_recognition/util/__init__.py:1-3`` (``relativity`` is out of scope)."""
from . import attractor, color, energy, math, normalize, orientation, regulator, selection, zoom  # noqa: F401
from . import apply_filter, get_dimensions, index_tensor  # noqa: F401
from .centroids import get_centroids  # noqa: F401
