"""Index tensors. Reference: ``slam_recognition/util/index_tensor.py:1-23`` (pinned by the reference's
``tests/test_index_tensor.py``: element ``[y, x]`` holds ``(x, y)`` while ``are_dimensions_reversed`` is set)."""
import itertools

import numpy as np

are_dimensions_reversed = True


def from_shape(shape):
    """``shape``: the NHWC-like shape of a tensor -> int32 ``dims + [len(dims)]`` array of coordinates."""
    dimension_list = [int(d) for d in list(shape)[1:-1]]
    index_tensor = np.zeros(dimension_list + [len(dimension_list)], dtype=np.int32)
    for xyz in itertools.product(*[range(d) for d in dimension_list]):
        index_tensor[xyz] = list(reversed(xyz)) if are_dimensions_reversed else xyz
    return index_tensor


def from_tensor(tensor):
    return from_shape(tuple(tensor.shape))
