"""Reference: ``slam_recognition/filters/rgby.py:6-14``."""
import numpy as np

from .. import _lib, _ops
from ..constant_convolutions.center_surround import rgby_3
from ..util.get_dimensions import get_dimensions


def rgby_filter(tensor):
    """``relu(conv3x3(tensor, rgby_3(rank - 2)))``: colour-opponent center-surround response."""
    n_dimensions = get_dimensions(tensor)
    rgby = np.reshape(rgby_3(n_dimensions), (3, 3, 3, 3))
    return _ops.conv2d(tensor, rgby, post=_lib.POST_RELU)
