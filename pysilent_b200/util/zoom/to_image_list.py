"""Reference: ``slam_recognition/util/zoom/to_image_list.py:7-15``."""
import numpy as np
import torch


def zoom_tensor_to_image_list(zoom, axis=2):
    """Split a pyramid tensor into one uint8 image per level (host arrays, for display)."""
    if isinstance(zoom, torch.Tensor):
        zoom = zoom.detach().cpu().numpy()
    return [np.squeeze(zoom[p:p + 1]).astype(dtype=np.uint8) for p in range(zoom.shape[0])]
