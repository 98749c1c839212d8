"""Reference: ``slam_recognition/filters/orientation.py:12-35``."""
import numpy as np

from .. import _lib, _ops
from ..constant_convolutions.edge_orientation_detector import rgb_2d_stripe_tensors
from ..constant_convolutions.gaussian_blur.gaussian_blur import blur_tensor
from ..util.get_dimensions import get_dimensions
from ..util.regulator import regulate_tensor


def orientation_filter(tensor, blur_size=7):
    """Oriented stripe response, then the blur regulator with value 1.0 and root 0.1."""
    dimensions = get_dimensions(tensor)
    stripes = np.reshape(rgb_2d_stripe_tensors(), (3, 3, 3, 3))
    blur = np.reshape(blur_tensor(dimensions, lengths=blur_size), (blur_size, blur_size, 3, 3))
    compiled_orient = _ops.conv2d(tensor, stripes, post=_lib.POST_RELU)
    return regulate_tensor(compiled_orient, blur, 1.0, .1)
