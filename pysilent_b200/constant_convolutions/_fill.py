"""Shared outer-product fill used by the stripe / edge / end generators.

The three reference generators (``stripe_tensor.py:61-68``, ``edge_tensor.py:65-72``, ``oriented_end_detector.py:47-53``)
declare ``[..., len(center_out), len(center_in)]`` but fill ``[in][out]``, so only square channel maps work and a
non-square map raises numpy's broadcast ``ValueError``. The net effect per tap with profile value ``z`` is
``W[tap, i, o] = |z| * surround_out[o] * surround_in[i]`` where ``z < 0``, ``|z| * center_out[o] * center_in[i]`` where
``z`` counts as centre, and 0 otherwise.
"""
import numpy as np


def fill_center_surround(profile, center_in, center_out, surround_in, surround_out, zero_is_center):
    """Expand a scalar tap profile into an HWIO filter.

    :param profile: float64 array of tap profile values (already normalised), shape ``[k]*ndim``.
    :param zero_is_center: whether ``z == 0`` taps take the centre colours (edge/end) or stay zero (stripe).
    """
    declared = (len(center_out), len(center_in))
    for name, ins, outs in (("surround", surround_in, surround_out), ("center", center_in, center_out)):
        produced = (len(ins), len(outs))
        if produced != declared:
            raise ValueError("could not broadcast input array from shape ({},{}) into shape ({},{}) [{} colours]".format(
                produced[0], produced[1], declared[0], declared[1], name))
    surround = np.asarray([[surround_out[o] * surround_in[i] for o in range(len(surround_out))]
                           for i in range(len(surround_in))], dtype=np.float64)
    center = np.asarray([[center_out[o] * center_in[i] for o in range(len(center_out))]
                         for i in range(len(center_in))], dtype=np.float64)
    magnitude = np.abs(profile)[..., np.newaxis, np.newaxis]
    is_surround = (profile < 0)[..., np.newaxis, np.newaxis]
    is_center = ((profile >= 0) if zero_is_center else (profile > 0))[..., np.newaxis, np.newaxis]
    out = np.zeros(profile.shape + declared)
    out = np.where(is_surround, surround * magnitude, out)
    out = np.where(is_center, center * magnitude, out)
    return out
