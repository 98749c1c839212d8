"""Reference: ``slam_recognition/constant_convolutions/contrast_adjust.py:1-4`` (no caller)."""


def contrast_adjust():
    """A fixed 1x3x3 colour-mixing list: each channel minus half of the other two."""
    return [[[1.0 if i == o else -0.5 for o in range(3)] for i in range(3)]]
