"""Reference: ``slam_recognition/util/color/get_value.py:6-12``."""
from ... import _ops


def get_value_from_color(color_tensor):
    """Channel mean with keepdims: ``reduce_sum(-1) * float32(1 / C)`` -> ``[N, h, w, 1]``."""
    return _ops.value_from_color(color_tensor)
