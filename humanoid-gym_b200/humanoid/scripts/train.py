"""Entry point: python humanoid/scripts/train.py --task=humanoid_ppo --headless
(reference scripts/train.py:36-43).  Under torchrun each rank trains its own env shard and the policy
gradients are all-reduced once per optimizer step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from humanoid.envs import *  # noqa: F401,F403,E402
from humanoid.utils import get_args, task_registry  # noqa: E402


def train(args):
    import torch
    import torch.distributed as dist
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl")
    env, env_cfg = task_registry.make_env(name=args.task, args=args)
    log_root = "default" if int(os.environ.get("RANK", "0")) == 0 else None
    ppo_runner, train_cfg = task_registry.make_alg_runner(env=env, name=args.task, args=args, log_root=log_root)
    ppo_runner.learn(num_learning_iterations=train_cfg.runner.max_iterations, init_at_random_ep_len=True)


if __name__ == "__main__":
    train(get_args())
