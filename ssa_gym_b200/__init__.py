"""ssa_gym_b200 — B200-native implementation of ssa-gym's per-step UKF hot path."""
__version__ = "0.1.0"
