"""Small tensor helpers (reference utils/math.py).  Only wrap_to_pi is on the hot path, where it is
evaluated inside the fused env kernel; these torch versions exist for API parity."""
import numpy as np
import torch


def quat_apply(q, v):
    shape = v.shape
    q = q.reshape(-1, 4)
    v = v.reshape(-1, 3)
    u = q[:, :3]
    t = u.cross(v, dim=-1) * 2
    return (v + q[:, 3:] * t + u.cross(t, dim=-1)).view(shape)


def quat_apply_yaw(quat, vec):
    qy = quat.clone().view(-1, 4)
    qy[:, :2] = 0.0
    qy = qy / qy.norm(p=2, dim=-1).clamp(min=1e-9).unsqueeze(-1)
    return quat_apply(qy, vec)


def wrap_to_pi(angles):
    angles %= 2 * np.pi
    angles -= 2 * np.pi * (angles > np.pi)
    return angles


def torch_rand_sqrt_float(lower, upper, shape, device):
    r = 2 * torch.rand(*shape, device=device) - 1
    r = torch.where(r < 0.0, -torch.sqrt(-r), torch.sqrt(r))
    return (upper - lower) * (r + 1.0) / 2.0 + lower
