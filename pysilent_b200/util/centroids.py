"""Block centroids of a value tensor. Reference: ``slam_recognition/util/centroids.py:21-71``."""
from .. import _ops


def _region(region_shape):
    dims = [float(d) for d in region_shape[1:3]]
    if any(int(d) != d or d <= 0 for d in dims):
        raise ValueError("Ambiguous dimension: %s" % (list(region_shape),))
    return int(dims[0]), int(dims[1])


def get_centroids(value_tensor, region_shape, debug=False):
    """``(value_centroids, total_pool)``: the L1 distance of every pixel to the value-weighted centroid of its
    ``region_shape[1:]`` block (``[N,h,w,1]``; NaN where a block is empty, like the reference's 0 / 0), and the block
    totals (``[N,ceil(h/rh),ceil(w/rw),1]``). ``region_shape`` is ``[1, rh, rw]`` as in ``recognition_testing.py:39``."""
    rh, rw = _region(region_shape)
    centroids, total, _ = _ops.get_centroids(value_tensor, rh, rw)
    return centroids, total


def get_centroids_array(value_tensor, region_shape, debug=False):
    """``corrected_centroid_pool`` ``[N,oh,ow,2]``: per block the value-weighted mean index ``(x, y)``."""
    rh, rw = _region(region_shape)
    return _ops.get_centroids(value_tensor, rh, rw, want_distance=False)[2]
