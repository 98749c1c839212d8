"""Feature-point emit and top-percent mask. Reference: ``slam_recognition/util/selection/top_value_points.py:8-45``."""
from ... import _ops
from ..color.get_value import get_value_from_color


def _region_extent(v):
    if int(v) != v:
        raise ValueError("Ambiguous dimension: %s" % (v,))   # TensorShape rejects non-integral dimensions
    return int(v)


def top_value_points(color_tensor, top_percent=0.1, value_tensor=None):
    """Zero everything whose value is below ``(1 - top_percent) * max + top_percent * min`` of its level."""
    if value_tensor is None:
        value_tensor = get_value_from_color(color_tensor)
    return _ops.top_value_points(color_tensor, value_tensor, top_percent)


def max_value_indices_region(color_tensor, region_shape, value_tensor=None):
    """int64 ``[K, 4]`` rows ``(level, y, x, 0)``, row-major, of the pixels that equal the maximum of their region window.

    The pool window is the whole level and the stride is ``region_shape[1:3]`` ('SAME'), exactly as the reference builds
    it (``top_value_points.py:39-41``); with the default ``(h/2, w/2)`` region that is four overlapping windows.
    """
    if value_tensor is None:
        value_tensor = get_value_from_color(color_tensor)
    return _ops.max_value_indices_region(value_tensor, _region_extent(region_shape[1]), _region_extent(region_shape[2]))
