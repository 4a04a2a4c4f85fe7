"""Seeded synthetic physics frames: the stand-in for what gym.simulate() + refresh_*
leave in Isaac Gym's four state tensors.  A ring of K pre-generated frames, produced outside any timed region.

This module is deliberately STANDALONE (torch + math only, no intra-package import): the product's
`SyntheticPhysics` uses it, and the bench's reference arm loads the same file by path into the test-only fake
`isaacgym` so that the unmodified reference consumes frames of the same distribution at the same (copy-only)
per-step cost as the product -- without importing anything of the product package."""
import math

import torch


def generate_ring(num_envs, device, cmd_ranges, env_origins, dof_lower, dof_upper, num_bodies, decimation=10, seed=5,
                  ring=6, p_base_contact=0.002, feet=(6, 12), knees=(4, 10)):
    """Returns dict(root (K,N,13), dof (K*decimation,N,nd,2), contact (K,N,nb,3), rigid (K,N,nb,13)) on `device`."""
    N, nb, nd = num_envs, num_bodies, len(dof_lower)
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)

    def randn(*s):
        return torch.randn(*s, generator=g, device=dev)

    def rand(*s):
        return torch.rand(*s, generator=g, device=dev)

    origins = torch.zeros(N, 3, device=dev) if env_origins is None else env_origins.to(dev)
    lo = torch.tensor(dof_lower, device=dev)
    hi = torch.tensor(dof_upper, device=dev)
    K = ring
    root = torch.zeros(K, N, 13, device=dev)
    dof = torch.zeros(K * decimation, N, nd, 2, device=dev)
    contact = torch.zeros(K, N, nb, 3, device=dev)
    rigid = torch.zeros(K, N, nb, 13, device=dev)
    for k in range(K):
        r = root[k]
        r[:, 0:2] = origins[:, 0:2] + (2 * rand(N, 2) - 1)
        r[:, 2] = 0.95 + 0.02 * randn(N)
        rpy = 0.1 * randn(N, 3)
        rpy[:, 2] = (2 * rand(N) - 1) * math.pi
        cr, sr = torch.cos(rpy[:, 0] / 2), torch.sin(rpy[:, 0] / 2)
        cp, sp = torch.cos(rpy[:, 1] / 2), torch.sin(rpy[:, 1] / 2)
        cy, sy = torch.cos(rpy[:, 2] / 2), torch.sin(rpy[:, 2] / 2)
        r[:, 3] = sr * cp * cy - cr * sp * sy
        r[:, 4] = cr * sp * cy + sr * cp * sy
        r[:, 5] = cr * cp * sy - sr * sp * cy
        r[:, 6] = cr * cp * cy + sr * sp * sy
        cx = cmd_ranges["lin_vel_x"][0] + (cmd_ranges["lin_vel_x"][1] - cmd_ranges["lin_vel_x"][0]) * rand(N)
        cyv = cmd_ranges["lin_vel_y"][0] + (cmd_ranges["lin_vel_y"][1] - cmd_ranges["lin_vel_y"][0]) * rand(N)
        yaw = rpy[:, 2]
        r[:, 7] = torch.cos(yaw) * cx - torch.sin(yaw) * cyv + 0.2 * randn(N)
        r[:, 8] = torch.sin(yaw) * cx + torch.cos(yaw) * cyv + 0.2 * randn(N)
        r[:, 9] = 0.2 * randn(N)
        r[:, 10:13] = 0.3 * randn(N, 3)

        clock = math.sin(2 * math.pi * (k + 0.25) / K)
        stance = torch.tensor([clock >= 0, clock < 0], device=dev).repeat(N, 1)
        in_contact = stance ^ (rand(N, 2) < 0.10)
        c = contact[k]
        for j, b in enumerate(feet):
            c[:, b, 2] = (200 + 400 * rand(N)) * in_contact[:, j]
            c[:, b, 0:2] = 20 * randn(N, 2) * in_contact[:, j:j + 1]
        hit = rand(N) < p_base_contact
        c[:, 0, :] = hit.unsqueeze(1) * (2.0 + 5 * rand(N, 3))

        rg = rigid[k]
        swing = (~in_contact).float()
        for j, b in enumerate(feet):
            side = 0.15 if j == 0 else -0.15
            rg[:, b, 0] = r[:, 0] + 0.05 * randn(N)
            rg[:, b, 1] = r[:, 1] + side + 0.03 * randn(N)
            rg[:, b, 2] = 0.05 + 0.06 * swing[:, j] * abs(clock)
            rg[:, b, 7:9] = 0.3 * randn(N, 2) * swing[:, j:j + 1]
        for j, b in enumerate(knees):
            side = 0.12 if j == 0 else -0.12
            rg[:, b, 0] = r[:, 0] + 0.02 * randn(N)
            rg[:, b, 1] = r[:, 1] + side + 0.02 * randn(N)
            rg[:, b, 2] = 0.45
    for s in range(K * decimation):
        q = 0.2 * randn(N, nd)
        dof[s, :, :, 0] = torch.max(torch.min(q, hi), lo)
        dof[s, :, :, 1] = randn(N, nd)
    return dict(root=root, dof=dof, contact=contact, rigid=rigid)
