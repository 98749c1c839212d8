"""Constant-convolution weight generators (host, numpy float64, HWIO layout ``[k, k, Cin, Cout]``).

Mirrors the export order of ``slam_recognition/constant_convolutions/__init__.py:1-5``: the 3x3
``oriented_end_detector.rgb_2d_end_tensors`` deliberately overrides the 7x7 one from ``edge_orientation_detector``.
The weights are generated once per plan and uploaded to ``__constant__`` memory by the CUDA layer.
"""
from .center_surround import *  # noqa: F401,F403
from .edge_orientation_detector import *  # noqa: F401,F403
from .gaussian_blur import *  # noqa: F401,F403
from .contrast_adjust import contrast_adjust  # noqa: F401
from .oriented_end_detector import end_tensor, simplex_end_tensors, rgb_2d_end_tensors  # noqa: F401
