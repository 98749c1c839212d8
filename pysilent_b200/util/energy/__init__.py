"""Mirrors ``slam_recognition/util/energy/__init__.py``."""
from .boosting import initialize_boosting, get_boosting
from .recovery import generate_constant_recovery, generate_input_based_recovery, generate_recovery, recovery_mode

__all__ = ["initialize_boosting", "get_boosting", "generate_constant_recovery", "generate_input_based_recovery",
           "generate_recovery", "recovery_mode"]
