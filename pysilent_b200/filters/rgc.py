"""Reference: ``slam_recognition/filters/rgc.py:6-18``."""
import numpy as np

from .. import _lib, _ops
from ..constant_convolutions.center_surround import midget_rgc
from ..util.get_dimensions import get_dimensions


def rgc_filter(tensor):
    """``relu(conv3x3(tensor, midget_rgc(rank - 2)))``: per-channel on-centre / off-surround response."""
    n_dimensions = get_dimensions(tensor)
    rgc = np.reshape(midget_rgc(n_dimensions), (3, 3, 3, 3))
    return _ops.conv2d(tensor, rgc, post=_lib.POST_RELU)
