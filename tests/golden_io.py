"""Helpers shared by the parity tests: load golden fixtures into oracle-style state dicts."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)

    def __getitem__(self, k):
        return self.z[k]

    def t(self, k):
        return torch.from_numpy(np.array(self.z[k]))

    def group(self, prefix):
        return {k[len(prefix):]: torch.from_numpy(np.array(self.z[k])) for k in self.z.files if k.startswith(prefix)}

    def has(self, k):
        return k in self.z.files


def oracle_state_from_golden(g, prefix="init."):
    """Build an oracle.env_oracle state dict from a golden snapshot."""
    from oracle import env_oracle as eo
    d = g.group(prefix)
    n = d["root_states"].shape[0]
    S = eo.new_state(n)
    for k, v in d.items():
        if k == "common_step_counter":
            S[k] = int(v)
        elif k in S:
            S[k] = v.clone()
    S["extras_time_outs"] = torch.zeros(n, dtype=torch.bool)
    S["obs_buf"] = S["obs_hist"].reshape(n, -1).clone()
    S["privileged_obs_buf"] = S["critic_hist"].reshape(n, -1).clone()
    return S
