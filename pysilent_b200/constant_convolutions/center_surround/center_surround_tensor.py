"""3^n center-surround (on/off, DoG-like) filter generator.

Reference: ``slam_recognition/constant_convolutions/center_surround/center_surround_tensor.py:17-48``; pinned by
``tests/test_center_surround_tensors.py:8-63`` (exact 1-D table, 2-D values .70710678 / 1. / 6.82842712) and by the
<= 1 s per call guard for 1..10 dimensions (``:66-74``), which a vectorised build meets easily.
"""
import numpy as np


def center_surround_tensor(ndim, center_in, center_out, surround_in, surround_out):
    """Return float64 ``[3]*ndim + [len(center_in), len(center_out)]``.

    A surround tap at Manhattan distance ``m`` from the centre weighs ``1/sqrt(m)`` times
    ``surround_in (x) surround_out``; the centre tap weighs the sum of all surround weights times
    ``center_in (x) center_out``.
    """
    assert ndim >= 1
    manhattan = np.abs(np.indices((3,) * ndim) - 1).sum(axis=0)
    ring = np.zeros(manhattan.shape)
    np.divide(1.0, np.sqrt(manhattan, where=manhattan > 0, out=np.ones(manhattan.shape)), out=ring,
              where=manhattan > 0)
    total = 0
    for w in ring.ravel().tolist():  # left-to-right, C order, like the reference's running total
        if w != 0.0:
            total += w
    surround = np.asarray([[o * i for o in surround_out] for i in surround_in], dtype=np.float64)
    center = np.asarray([[o * i for o in center_out] for i in center_in], dtype=np.float64)
    out = np.ndarray(shape=[3] * ndim + [len(center_in), len(center_out)])
    out[...] = surround * ring[..., np.newaxis, np.newaxis]
    out[(1,) * ndim] = center * total
    return out
