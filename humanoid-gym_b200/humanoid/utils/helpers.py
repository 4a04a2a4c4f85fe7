"""Setup / IO helpers with the reference's names and behaviour (reference utils/helpers.py),
without the isaacgym dependency: argument parsing and SimParams are provided here."""
import argparse
import copy
import datetime
import os
import random

import numpy as np
import torch

from humanoid import LEGGED_GYM_ROOT_DIR, LEGGED_GYM_ENVS_DIR  # noqa: F401

SIM_PHYSX = 1
SIM_FLEX = 0


def class_to_dict(obj) -> dict:
    """Recursive dir()-ordered (hence ALPHABETICAL) dump of a config object -- the reward order of the
    env kernel depends on this ordering (reference helpers.py:44-59)."""
    if not hasattr(obj, "__dict__"):
        return obj
    out = {}
    for key in dir(obj):
        if key.startswith("_"):
            continue
        val = getattr(obj, key)
        out[key] = [class_to_dict(v) for v in val] if isinstance(val, list) else class_to_dict(val)
    return out


def update_class_from_dict(obj, d):
    for key, val in d.items():
        attr = getattr(obj, key, None)
        if isinstance(attr, type):
            update_class_from_dict(attr, val)
        else:
            setattr(obj, key, val)


def set_seed(seed):
    if seed == -1:
        seed = np.random.randint(0, 10000)
    print("Setting seed: {}".format(seed))
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)


class _Namespace:
    pass


class SimParams:
    """Stand-in for gymapi.SimParams.  `dt` is a C float in the real struct, so a value written
    there reads back float32-rounded -- which makes decimation*dt = 0.010000000475 and the command
    resampling period int(8/dt) = 799 (SURVEY.md section 8c hazard 3)."""

    def __init__(self):
        self._dt = float(np.float32(1.0 / 60.0))
        self.substeps = 2
        self.up_axis = 1
        self.use_gpu_pipeline = True
        self.gravity = [0.0, 0.0, -9.81]
        self.physx = _Namespace()
        self.physx.use_gpu = True
        self.physx.num_subscenes = 0
        self.physx.num_threads = 0

    @property
    def dt(self):
        return self._dt

    @dt.setter
    def dt(self, v):
        self._dt = float(np.float32(v))


def parse_sim_params(args, cfg):
    sim_params = SimParams()
    if args.physics_engine == SIM_PHYSX:
        sim_params.physx.use_gpu = args.use_gpu
        sim_params.physx.num_subscenes = args.subscenes
    sim_params.use_gpu_pipeline = args.use_gpu_pipeline
    if "sim" in cfg:
        for k, v in cfg["sim"].items():
            if k == "physx":
                for kk, vv in v.items():
                    setattr(sim_params.physx, kk, vv)
            else:
                setattr(sim_params, k, v)
    if args.physics_engine == SIM_PHYSX and args.num_threads > 0:
        sim_params.physx.num_threads = args.num_threads
    return sim_params


def get_load_path(root, load_run=-1, checkpoint=-1):
    def month_key(name):
        return (datetime.datetime.strptime(name[:3], "%b").month, int(name[3:5]), name[6:])

    try:
        runs = os.listdir(root)
        try:
            runs.sort(key=month_key)
        except ValueError as e:
            print("WARNING - Could not sort runs by month: " + str(e))
            runs.sort()
        if "exported" in runs:
            runs.remove("exported")
        last_run = os.path.join(root, runs[-1])
    except Exception:
        raise ValueError("No runs in this directory: " + root)
    load_run = last_run if load_run == -1 else os.path.join(root, load_run)
    if checkpoint == -1:
        models = [f for f in os.listdir(load_run) if "model" in f]
        models.sort(key=lambda m: "{0:0>15}".format(m))
        model = models[-1]
    else:
        model = "model_{}.pt".format(checkpoint)
    return os.path.join(load_run, model)


def update_cfg_from_args(env_cfg, cfg_train, args):
    if env_cfg is not None and args.num_envs is not None:
        env_cfg.env.num_envs = args.num_envs
    if cfg_train is not None:
        if args.seed is not None:
            cfg_train.seed = args.seed
        r = cfg_train.runner
        if args.max_iterations is not None:
            r.max_iterations = args.max_iterations
        if args.resume:
            r.resume = args.resume
        for name in ("experiment_name", "run_name", "load_run", "checkpoint"):
            v = getattr(args, name)
            if v is not None:
                setattr(r, name, v)
    return env_cfg, cfg_train


def get_args(argv=None):
    """Same flags as the reference CLI (helpers.py:167-245 + the gymutil.parse_arguments base set)."""
    p = argparse.ArgumentParser(description="RL Policy")
    p.add_argument("--sim_device", type=str, default="cuda:0", help="Physics device: cpu | cuda:N")
    p.add_argument("--pipeline", type=str, default="gpu", help="Tensor API pipeline (cpu/gpu)")
    p.add_argument("--graphics_device_id", type=int, default=0)
    p.add_argument("--physx", action="store_true")
    p.add_argument("--flex", action="store_true")
    p.add_argument("--num_threads", type=int, default=0)
    p.add_argument("--subscenes", type=int, default=0)
    p.add_argument("--slices", type=int, default=None)
    p.add_argument("--task", type=str, default="XBotL_free")
    p.add_argument("--resume", action="store_true", default=False)
    p.add_argument("--experiment_name", type=str)
    p.add_argument("--run_name", type=str)
    p.add_argument("--load_run", type=str)
    p.add_argument("--checkpoint", type=int)
    p.add_argument("--headless", action="store_true", default=False)
    p.add_argument("--horovod", action="store_true", default=False, help="parsed and ignored, as in the reference")
    p.add_argument("--rl_device", type=str, default="cuda:0")
    p.add_argument("--num_envs", type=int)
    p.add_argument("--seed", type=int)
    p.add_argument("--max_iterations", type=int)
    args = p.parse_args(argv)
    dev = args.sim_device
    if dev in ("cpu", "cuda"):
        args.sim_device_type, args.compute_device_id = dev, 0
    else:
        kind, idx = dev.split(":")
        args.sim_device_type, args.compute_device_id = kind, int(idx)
    args.use_gpu_pipeline = args.pipeline.lower() in ("gpu", "cuda") and args.sim_device_type == "cuda"
    args.physics_engine = SIM_FLEX if args.flex else SIM_PHYSX
    args.use_gpu = args.sim_device_type == "cuda"
    if args.slices is None:
        args.slices = args.subscenes
    args.sim_device_id = args.compute_device_id
    args.sim_device = args.sim_device_type
    if args.sim_device == "cuda":
        args.sim_device += f":{args.sim_device_id}"
    # one process per GPU under torchrun: every rank drives its own device
    if "LOCAL_RANK" in os.environ and args.sim_device_type == "cuda":
        lr = int(os.environ["LOCAL_RANK"])
        args.sim_device = args.rl_device = f"cuda:{lr}"
        args.sim_device_id = args.compute_device_id = lr
    return args


def export_policy_as_jit(actor_critic, path):
    os.makedirs(path, exist_ok=True)
    path = os.path.join(path, "policy_1.pt")
    model = copy.deepcopy(actor_critic.actor).to("cpu")
    torch.jit.script(model).save(path)
