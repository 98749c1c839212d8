"""Reference: ``slam_recognition/util/get_dimensions.py:7-15``."""
import numpy as np
import torch


def get_dimensions(tensor):
    """Spatial rank of an ``[N, spatial..., C]`` tensor: ``rank - 2``. Raises ``TypeError`` for anything else."""
    if isinstance(tensor, torch.Tensor):
        return tensor.dim() - 2
    if isinstance(tensor, np.ndarray):
        return len(tensor.shape) - 2
    raise TypeError("Input to orientation filter must either be tensor or numpy array.")
