"""go2/utility.py:4-8: yaw of a (w, x, y, z) quaternion = scipy `as_euler('xyz')[2]`."""


def quat_to_yaw(quat):
    """Works on torch tensors or numpy arrays with a trailing dim of 4."""
    w, x, y, z = quat[..., 0], quat[..., 1], quat[..., 2], quat[..., 3]
    num, den = 2.0 * (w * z + x * y), 1.0 - 2.0 * (y * y + z * z)
    if hasattr(quat, "atan2") or type(quat).__module__.startswith("torch"):
        import torch
        return torch.atan2(num, den)
    import numpy as np
    return np.arctan2(num, den)
