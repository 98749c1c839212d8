"""Reference: ``slam_recognition/util/apply_filter.py:4-7``."""
from .. import _ops


def apply_filter(tensor, filter):  # noqa: A002  (reference argument name)
    """``conv2d(tensor, float32(filter), strides 1, 'SAME')`` with the filter's native HWIO shape (no activation)."""
    return _ops.conv2d(tensor, filter)
