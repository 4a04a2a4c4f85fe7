"""CPU tests of the host-side drop-in API: config trees, registry, CLI, helpers (no GPU, no compute calls)."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _norm(x):
    if isinstance(x, dict):
        return {k: _norm(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_norm(v) for v in x]
    if isinstance(x, float) and x == int(x):
        return int(x)
    return x


def test_config_trees_equal_the_reference():
    from humanoid.envs import XBotLCfg, XBotLCfgPPO
    from humanoid.utils.helpers import class_to_dict
    ref = json.load(open(os.path.join(HERE, "golden", "cfg_dump.json")))
    mine = {"XBotLCfg": class_to_dict(XBotLCfg()), "XBotLCfgPPO": class_to_dict(XBotLCfgPPO())}
    for name in ref:
        assert _norm(json.loads(json.dumps(mine[name]))) == _norm(ref[name]), name
    # reward order is alphabetical (dir()), which fixes the fp32 summation order of the env kernel
    assert list(mine["XBotLCfg"]["rewards"]["scales"].keys()) == sorted(mine["XBotLCfg"]["rewards"]["scales"].keys())


def test_config_instances_are_independent():
    from humanoid.envs import XBotLCfg
    a, b = XBotLCfg(), XBotLCfg()
    a.env.num_envs = 7
    assert b.env.num_envs == 4096 and XBotLCfg.env.num_envs == 4096


def test_registry_and_cli():
    from humanoid.envs import XBotLFreeEnv
    from humanoid.utils import task_registry, get_args
    assert task_registry.get_task_class("humanoid_ppo") is XBotLFreeEnv
    env_cfg, train_cfg = task_registry.get_cfgs("humanoid_ppo")
    assert env_cfg.seed == train_cfg.seed == 5
    args = get_args(["--task=humanoid_ppo", "--headless", "--num_envs=128", "--seed=9", "--max_iterations=3", "--rl_device=cuda:0"])
    assert args.task == "humanoid_ppo" and args.headless and args.num_envs == 128 and args.sim_device == "cuda:0"
    assert args.physics_engine == 1 and args.use_gpu_pipeline and args.horovod is False
    with pytest.raises(ValueError):
        task_registry.make_env("no_such_task", args=args)


def test_update_cfg_from_args_and_sim_params():
    import numpy as np
    from humanoid.envs import XBotLCfg, XBotLCfgPPO
    from humanoid.utils.helpers import update_cfg_from_args, get_args, parse_sim_params, class_to_dict
    args = get_args(["--num_envs=64", "--seed=3", "--max_iterations=11", "--experiment_name=x", "--resume"])
    e, t = update_cfg_from_args(XBotLCfg(), XBotLCfgPPO(), args)
    assert e.env.num_envs == 64 and t.seed == 3 and t.runner.max_iterations == 11 and t.runner.experiment_name == "x" and t.runner.resume
    sp = parse_sim_params(args, {"sim": class_to_dict(e.sim)})
    assert sp.dt == float(np.float32(0.001)) and sp.physx.num_position_iterations == 4
    assert int(8.0 / (10 * sp.dt)) == 799          # the float32 dt of gymapi.SimParams (SURVEY hazard 3)


def test_get_load_path(tmp_path):
    from humanoid.utils.helpers import get_load_path
    for run in ("Jan02_10-00-00_a", "Mar05_09-00-00_b", "exported"):
        (tmp_path / run).mkdir()
    for m in ("model_0.pt", "model_100.pt", "model_20.pt"):
        (tmp_path / "Mar05_09-00-00_b" / m).write_text("")
    assert get_load_path(str(tmp_path)).endswith("Mar05_09-00-00_b/model_100.pt")
    assert get_load_path(str(tmp_path), load_run="Jan02_10-00-00_a", checkpoint=7).endswith("Jan02_10-00-00_a/model_7.pt")
    with pytest.raises(ValueError):
        get_load_path(str(tmp_path / "missing"))


def test_wrap_to_pi_and_quat_helpers():
    import math
    import torch
    from humanoid.utils.math import wrap_to_pi, quat_apply_yaw
    a = wrap_to_pi(torch.tensor([0.0, 3.5, -3.5, 7.0]))
    assert torch.allclose(a, torch.tensor([0.0, 3.5 - 2 * math.pi, 2 * math.pi - 3.5, 7.0 - 2 * math.pi]), atol=1e-6)
    q = torch.tensor([[0.0, 0.0, math.sin(math.pi / 4), math.cos(math.pi / 4)]])
    v = quat_apply_yaw(q, torch.tensor([[1.0, 0.0, 0.0]]))
    assert torch.allclose(v, torch.tensor([[0.0, 1.0, 0.0]]), atol=1e-6)
