"""Mirrors ``slam_recognition/constant_convolutions/center_surround/__init__.py:1-3``."""
from .center_surround_tensor import center_surround_tensor
from .rgby import rgby, rgby_3
from .rgc import midget_rgc, midget_rgc_full

__all__ = ["center_surround_tensor", "rgby", "rgby_3", "midget_rgc"]
