"""Reference: ``slam_recognition/util/math/almost_equal.py:4-9`` (elementwise helpers; the centroid kernel fuses
``equality_distance``)."""
import torch

from ... import _ops


def almost_equal(tensor1, tensor2, diff=0.51):
    a, b = _ops.as_device_tensor(tensor1), _ops.as_device_tensor(tensor2)
    return (a - b + diff) <= diff * 2


def equality_distance(tensor1, tensor2):
    return torch.abs(_ops.as_device_tensor(tensor1) - _ops.as_device_tensor(tensor2))
