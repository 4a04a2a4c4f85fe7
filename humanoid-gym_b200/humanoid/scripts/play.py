"""python humanoid/scripts/play.py --task=humanoid_ppo --load_run <run> [--checkpoint N]
(reference scripts/play.py:48-150 without the viewer / video / plotting): load a checkpoint through the task
registry, roll the policy out on a few envs with domain randomisation and noise off, and export the actor as
TorchScript (`exported/policies/policy_1.pt`), the file scripts/sim2sim.py and the robot consume."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from humanoid import LEGGED_GYM_ROOT_DIR  # noqa: E402
from humanoid.envs import *  # noqa: F401,F403,E402
from humanoid.utils import get_args, export_policy_as_jit, task_registry  # noqa: E402

EXPORT_POLICY = True


def play(args, steps=None):
    import torch
    env_cfg, train_cfg = task_registry.get_cfgs(name=args.task)
    env_cfg.env.num_envs = min(env_cfg.env.num_envs, 1 if args.num_envs is None else args.num_envs)   # reference play.py:51-63
    env_cfg.sim.max_gpu_contact_pairs = 2 ** 10
    env_cfg.terrain.curriculum = False
    env_cfg.noise.add_noise = False
    env_cfg.domain_rand.push_robots = False
    env_cfg.domain_rand.joint_angle_noise = 0.0
    env_cfg.noise.curriculum = False
    env_cfg.noise.noise_level = 0.5
    train_cfg.seed = 123145
    env, _ = task_registry.make_env(name=args.task, args=args, env_cfg=env_cfg)
    env.set_camera(env_cfg.viewer.pos, env_cfg.viewer.lookat) if hasattr(env, "set_camera") else None
    obs = env.get_observations()
    train_cfg.runner.resume = bool(getattr(args, "load_run", None) or getattr(args, "resume", False))
    ppo_runner, train_cfg = task_registry.make_alg_runner(env=env, name=args.task, args=args, train_cfg=train_cfg,
                                                          log_root="default" if train_cfg.runner.resume else None)
    policy = ppo_runner.get_inference_policy(device=env.device)
    path = None
    if EXPORT_POLICY:
        path = os.path.join(LEGGED_GYM_ROOT_DIR, "logs", train_cfg.runner.experiment_name, "exported", "policies")
        export_policy_as_jit(ppo_runner.alg.actor_critic, path)
        print("Exported policy as jit script to: ", path)
    n = int(steps if steps is not None else 10 * env.max_episode_length)
    with torch.inference_mode():
        for _ in range(n):
            actions = policy(obs.detach())
            obs, _, rews, dones, infos = env.step(actions.detach())
    return path


if __name__ == "__main__":
    play(get_args())
