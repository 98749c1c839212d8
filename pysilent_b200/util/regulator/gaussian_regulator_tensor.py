"""Adaptive gain. Reference: ``slam_recognition/util/regulator/gaussian_regulator_tensor.py:10-36``."""
from ... import _ops


def regulate_tensor(input_tensor, blur_tensor, regulation_value, regulation_root=1.0 / 2.0, strides=(1, 1, 1, 1),
                    padding='SAME'):
    """``input * (regulation_value / pow(min(conv2d(input, blur), 1), regulation_root))`` in one kernel.

    Where the blurred value is >= 1 the gain is exactly ``regulation_value``; where it is 0 the gain is ``inf`` and a
    zero input yields NaN, as in the reference graph.
    """
    if tuple(strides) != (1, 1, 1, 1) or padding != 'SAME':
        raise ValueError("regulate_tensor supports strides (1,1,1,1) and 'SAME' padding only")
    return _ops.regulate(input_tensor, blur_tensor, float(regulation_value), float(regulation_root))
