"""pysilent_b200 -- B200-native (sm_100a) implementation of pySILEnT's ``slam_recognition`` filter pipeline.

Same operator surface as the reference package for the hot path (``slam_recognition/__init__.py:3-5`` exports
``center_surround_tensor``, ``stripe_tensor``, ``simplex_stripe_tensors`` and ``zoom``); the image-side operators run in
hand-written CUDA behind ``libsilent_b200.so`` and return ``torch`` CUDA tensors instead of lazy ``tf.Tensor``s.
"""
__version__ = "0.1.0"

from .constant_convolutions.center_surround import center_surround_tensor  # noqa: F401
from .constant_convolutions.edge_orientation_detector import stripe_tensor, simplex_stripe_tensors  # noqa: F401
from .util import zoom  # noqa: F401
from .pipeline import LineEndPipeline  # noqa: F401
from .recognition_testing import LineEndDisplayer  # noqa: F401
