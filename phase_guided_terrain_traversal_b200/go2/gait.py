"""Gait phase constants and the foot-height reference (go2/gait.py:8-49), host numpy version.

The device copy used by the reward lives in csrc/pgtt_env.cuh:gait_get_z; this one exists for the
public API (`gait.PHASES`, `gait.p_stance`, `gait.get_z`) and for tests.
"""
import numpy as np

PHASES = np.array([0.0, np.pi, np.pi, 0.0])   # FR, FL, RR, RL : trot
p_stance = 0.5


def cubic_hermite(t, p0, p1):
    """Zero end tangents: h00(t) p0 + h01(t) p1."""
    t2, t3 = t * t, t * t * t
    return (2 * t3 - 3 * t2 + 1) * p0 + (-2 * t3 + 3 * t2) * p1


def get_z(phi, swing_height=0.08, swing_min=0.0):
    """Foot height target over one gait cycle phi in [0, 2 pi): stance, then up to the peak, then down."""
    phi = np.asarray(phi, dtype=np.float64)
    swing_height = np.broadcast_to(np.asarray(swing_height, dtype=np.float64), phi.shape)
    t_stance = 2 * np.pi * p_stance
    t_swing = 2 * np.pi * (1 - p_stance) / 2
    t_peak = 2 * np.pi * (1 + p_stance) / 2
    up = cubic_hermite((phi - t_stance) / t_swing, swing_min, swing_height)
    down = cubic_hermite((phi - t_peak) / t_swing, swing_height, swing_min)
    return np.where(phi <= t_stance, swing_min, np.where(phi <= t_peak, up, down))
