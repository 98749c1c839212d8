"""Recovery of exhausted cells. Reference: ``slam_recognition/util/energy/recovery.py:4-22``."""
import torch

from ... import _ops


def generate_constant_recovery(tensor_in, recovery_amount=10):
    t = _ops.as_device_tensor(tensor_in)
    return torch.full_like(t, float(recovery_amount))


def generate_input_based_recovery(tensor_in, recovery_percentage=0.8):
    return _ops.as_device_tensor(tensor_in) * float(recovery_percentage)


def recovery_mode(is_input_based=False, is_constant=True):
    """The selection of ``generate_recovery`` as the code the fused boosting kernel takes: 1 constant, 2 input, 3 both."""
    if is_input_based and not is_constant:
        return 2
    if is_constant and not is_input_based:
        return 1
    if is_input_based and is_constant:
        return 3
    raise ValueError("You must choose a type of recovery")


def generate_recovery(tensor_in, is_input_based=False, is_constant=True):
    """selects which type of recovery to be used for neurons."""
    mode = recovery_mode(is_input_based, is_constant)
    if mode == 2:
        return generate_input_based_recovery(tensor_in)
    if mode == 1:
        return generate_constant_recovery(tensor_in)
    a, b = generate_input_based_recovery(tensor_in), generate_constant_recovery(tensor_in)
    return torch.where(torch.isnan(a), a, torch.maximum(a, b))
