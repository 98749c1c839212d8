"""Border mask. Reference: ``slam_recognition/util/selection/isolate_rectangle.py:19-23``."""
from ... import _ops


def pad_inwards(tensor, paddings):
    """Multiply by a box that is 0 on the given border widths and 1 inside (so NaN * 0 stays NaN).

    ``paddings`` is ``[[0, 0], [top, bottom], [left, right], [0, 0]]`` as passed to ``tf.pad``.
    """
    paddings = [[int(a), int(b)] for a, b in paddings]
    if len(paddings) != 4 or paddings[0] != [0, 0] or paddings[3] != [0, 0]:
        raise ValueError("pad_inwards supports spatial paddings [[0,0],[t,b],[l,r],[0,0]] only")
    return _ops.pad_inwards(tensor, paddings[1][0], paddings[1][1], paddings[2][0], paddings[2][1])
