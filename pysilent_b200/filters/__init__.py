"""Filter callables. Mirrors ``slam_recognition/filters/__init__.py:1-3``."""
from .orientation import orientation_filter
from .rgby import rgby_filter
from .rgc import rgc_filter

__all__ = ["orientation_filter", "rgby_filter", "rgc_filter"]
