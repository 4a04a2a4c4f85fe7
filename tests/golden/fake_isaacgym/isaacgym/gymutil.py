"""gymutil subset: device-string parsing, argument parsing, sim-config merge."""
import argparse
from . import gymapi


def parse_device_str(device_str):
    device_str = str(device_str)
    if device_str in ("cpu", "cuda"):
        return device_str, 0
    parts = device_str.split(":")
    assert len(parts) == 2 and parts[0] == "cuda", f"bad device string {device_str}"
    return "cuda", int(parts[1])


def parse_arguments(description="fake", headless=False, no_graphics=False, custom_parameters=()):
    p = argparse.ArgumentParser(description=description)
    p.add_argument("--sim_device", type=str, default="cuda:0")
    p.add_argument("--pipeline", type=str, default="gpu")
    p.add_argument("--graphics_device_id", type=int, default=0)
    p.add_argument("--physx", action="store_true")
    p.add_argument("--flex", action="store_true")
    p.add_argument("--num_threads", type=int, default=0)
    p.add_argument("--subscenes", type=int, default=0)
    p.add_argument("--slices", type=int, default=0)
    for a in custom_parameters:
        kw = {k: v for k, v in a.items() if k != "name"}
        p.add_argument(a["name"], **kw)
    args, _ = p.parse_known_args()
    args.sim_device_type, args.compute_device_id = parse_device_str(args.sim_device)
    pipeline = args.pipeline.lower()
    args.use_gpu_pipeline = pipeline in ("gpu", "cuda") and args.sim_device_type == "cuda"
    args.physics_engine = gymapi.SIM_PHYSX
    args.use_gpu = args.sim_device_type == "cuda"
    if args.slices is None or args.slices == 0:
        args.slices = args.subscenes
    return args


def parse_sim_config(sim_cfg, sim_params):
    """Copy a nested dict into SimParams.  `dt` is stored as a C float in the
    real SimParams struct, so it comes back as float32-rounded."""
    import numpy as np
    for k, v in sim_cfg.items():
        if k == "physx":
            for kk, vv in v.items():
                setattr(sim_params.physx, kk, vv)
        elif k == "dt":
            sim_params.dt = float(np.float32(v))
        elif k == "gravity":
            sim_params.gravity = gymapi.Vec3(*v)
        else:
            setattr(sim_params, k, v)
