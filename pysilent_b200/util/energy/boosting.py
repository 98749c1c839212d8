"""Stateful boosting non-max. Reference: ``slam_recognition/util/energy/boosting.py:6-42``."""
import torch

from ... import _lib, _ops
from .recovery import recovery_mode


def initialize_boosting(input_tensor, initial_multiplier=8):
    """The ``tf.Variable`` of the reference: a CUDA float32 tensor shaped like ``input_tensor`` filled with
    ``initial_multiplier``; ``get_boosting`` updates it in place."""
    t = _ops.as_device_tensor(input_tensor)
    return torch.full_like(t, float(initial_multiplier))


def get_boosting(input_tensor, exhaustion_tensor, exhaustion_max=1, excitation_max=1, input_based_recovery=False,
                 constant_recovery=True, for_visualizing=False):
    """Cells whose ``input ** energy`` equals the 3x3 maximum fire; firing exhausts the cell's energy, which recovers by a
    constant (or input-based) amount per call. Returns ``(has_fired, update_energy)``; ``exhaustion_tensor`` holds the
    new state afterwards (the reference's ``assign``). With ``for_visualizing`` both are the 3-channel display versions
    (``boosting.py:35-40``)."""
    mode = recovery_mode(input_based_recovery, constant_recovery)
    x = _ops.as_device_tensor(input_tensor)
    fired = _ops.boosting(x, exhaustion_tensor, exhaustion_max, excitation_max, mode)
    if not for_visualizing:
        return fired, exhaustion_tensor
    fired_input = _ops.pointwise(fired, _lib.PW_PRODUCT, x)                  # grayscale_to_rgb(has_fired) * input
    if exhaustion_max == 1 and excitation_max == 1:
        update = _ops.pointwise(exhaustion_tensor, _lib.PW_ENERGY_DISPLAY)
    else:                                                                     # general maxima: same two roundings
        normer = 255.0 / (exhaustion_max + excitation_max)
        centerer = (excitation_max / (exhaustion_max + excitation_max)) * 255.0
        update = exhaustion_tensor * normer + centerer
    return fired_input.expand(-1, -1, -1, 3).contiguous(), update.expand(-1, -1, -1, 3).contiguous()
