"""CPU tests of the host-side drop-in API: config trees, registry, CLI, helpers (no GPU, no compute calls)."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


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


def test_sim2sim_plumbing_matches_a_direct_restatement():
    """scripts/sim2sim.py (reference scripts/sim2sim.py:113-160): the policy input at every 100 Hz tick is the 15-frame
    stack of clipped 47-wide frames, oldest first; actions are clipped and scaled into PD targets.  Runs on CPU with the
    reference's shipped actor (weights from the committed KAT fixture)."""
    import importlib.util
    import numpy as np
    import torch
    from golden_io import Golden
    spec = importlib.util.spec_from_file_location("hg_sim2sim", os.path.join(ROOT, "humanoid-gym_b200", "humanoid", "scripts", "sim2sim.py"))
    s2s = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(s2s)
    k = Golden("policy_example_kat.npz")
    w = k.group("w.")
    layers = []
    for i in (0, 2, 4, 6):
        lin = torch.nn.Linear(w[f"{i}.weight"].shape[1], w[f"{i}.weight"].shape[0])
        lin.weight.data.copy_(w[f"{i}.weight"]), lin.bias.data.copy_(w[f"{i}.bias"])
        layers += [lin] + ([torch.nn.ELU()] if i < 6 else [])
    policy = torch.jit.script(torch.nn.Sequential(*layers))
    rec = []
    calls, sec = s2s.run(policy, low_level_steps=400, record=rec)
    assert calls == 40 and len(rec) == 40 and sec > 0
    x, a = rec[-1]
    assert x.shape == (1, 705) and np.isfinite(x).all() and np.abs(x).max() <= 18 and np.abs(a).max() <= 18
    # frame i of call n is frame i+1 of call n-1 (history shift), the newest frame carries the gait clock of its tick
    assert np.array_equal(rec[-1][0][0, :658], rec[-2][0][0, 47:])
    t = 390 * 0.001
    assert abs(x[0, 658] - np.sin(2 * np.pi * t / 0.64)) < 1e-6 and abs(x[0, 659] - np.cos(2 * np.pi * t / 0.64)) < 1e-6
    assert x[0, 660] == np.float32(0.4 * 2.0)                       # cmd.vx * obs_scales.lin_vel
    assert np.allclose(x[0, 658 + 29:658 + 41], rec[-2][1])          # last action
    # the first call sees an all-zero history except its own frame, and the actor's known answer on zeros is reproduced
    y0 = policy(torch.zeros(1, 705))[0].detach().numpy()
    np.testing.assert_allclose(y0, [0.0847, -0.0234, 0.0057, 0.2348, 0.6382, -0.2275, -0.1129, -0.1501, 0.2042, 0.3535, 0.0077, -0.4530], atol=5e-5)
