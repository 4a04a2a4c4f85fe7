#!/usr/bin/env python
"""Reference arm of bench.py: the UNMODIFIED roboterax/humanoid-gym PyTorch path, timed.

    python tools/reference_arm.py --device cpu|cuda:0 --num-envs 4096 --steps K --warmup W [--threads T]

Runs the reference's own stock code path -- `task_registry.make_env` -> `XBotLFreeEnv`, `make_alg_runner` ->
`OnPolicyRunner.learn(W + K, init_at_random_ep_len=True)` (what scripts/train.py does, reference
scripts/train.py:41-46) -- at the full configuration (T = num_steps_per_env = 60, 2 epochs x 4 minibatches).
Nothing of the reference is patched: the only instrumentation is a wrapper around `runner.log`, which the
reference calls once at the end of every iteration, to timestamp the iteration boundaries.

Where the reference comes from: `baseline/_ref` (pip install --no-deps --target of /root/reference, git-ignored,
travels to the GPU box) or /root/reference itself when present.  Isaac Gym is the test-only fake
(tests/golden/fake_isaacgym) in "ring" mode: pre-generated synthetic frames, copied per refresh -- the same
frames and per-step physics cost as the product arm's SyntheticPhysics.  This process never imports the product
package (the name `humanoid` resolves to the reference) and never maps libhg_b200.so.

Prints ONE JSON line: {"env_steps_per_sec", "ms_per_iteration", "collection_s", "learn_s", "device", "threads", ...}.
"""
import argparse
import json
import os
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference():
    for cand, kind in ((os.path.join(ROOT, "baseline", "_ref"), "baseline/_ref (pip install --no-deps --target of the reference)"),
                       ("/root/reference", "/root/reference")):
        if os.path.isdir(os.path.join(cand, "humanoid", "algo")):
            return cand, kind
    return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()

    ref, ref_kind = find_reference()
    if ref is None:
        print(json.dumps({"unavailable": "no reference tree (baseline/_ref or /root/reference) on this box"}))
        return 0
    # the product package dir must NOT be importable here: `humanoid` has to be the reference
    sys.path[:] = [p for p in sys.path if os.path.basename(p.rstrip("/")) != "humanoid-gym_b200"]
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden", "fake_isaacgym"))
    sys.path.insert(0, ref)
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    os.environ["WANDB_MODE"] = "disabled"
    os.environ["HG_FAKE_GYM"] = "ring"
    sys.argv = sys.argv[:1]                       # the reference's get_args() parses sys.argv

    import torch
    is_cuda = a.device.startswith("cuda")
    if a.threads > 0:
        torch.set_num_threads(a.threads)
    if is_cuda:
        # the reference's fp32 path: no TF32 (torch defaults already keep matmul TF32 off; make it explicit)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        torch.cuda.set_device(torch.device(a.device))

    import humanoid                                # noqa: F401  (the reference)
    assert os.path.realpath(humanoid.__file__).startswith(os.path.realpath(ref)), humanoid.__file__
    from humanoid.envs import XBotLFreeEnv         # noqa: F401  registers humanoid_ppo
    from humanoid.utils import task_registry

    dev_id = int(a.device.split(":")[1]) if ":" in a.device else 0
    args = argparse.Namespace(
        task="humanoid_ppo", resume=False, experiment_name=None, run_name=None, load_run=None, checkpoint=None,
        headless=True, horovod=False, rl_device=a.device, num_envs=a.num_envs, seed=5, max_iterations=None,
        physics_engine=1, use_gpu=is_cuda, use_gpu_pipeline=is_cuda, subscenes=0, num_threads=0,
        sim_device=a.device, sim_device_type="cuda" if is_cuda else "cpu", compute_device_id=dev_id,
        sim_device_id=dev_id, device=a.device)
    env, _ = task_registry.make_env(name="humanoid_ppo", args=args)
    assert str(env.device) == a.device, (env.device, a.device)
    log_root = tempfile.mkdtemp(prefix="hg_ref_arm_")
    runner, train_cfg = task_registry.make_alg_runner(env=env, name="humanoid_ppo", args=args, log_root=log_root)
    T = runner.num_steps_per_env

    marks, parts = [], []
    stock_log = runner.log

    def log_and_mark(locs, *x, **k):               # iteration boundary (the reference calls log() once per iteration)
        if is_cuda:
            torch.cuda.synchronize()
        marks.append(time.time())
        parts.append((locs["collection_time"], locs["learn_time"]))
        if os.environ.get("HG_REF_VERBOSE") == "1":
            stock_log(locs, *x, **k)
        else:                                      # keep the reference's own bookkeeping, drop the console table
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                stock_log(locs, *x, **k)

    runner.log = log_and_mark
    t_start = time.time()
    runner.learn(num_learning_iterations=a.warmup + a.steps, init_at_random_ep_len=True)
    W, K = a.warmup, a.steps
    t0 = marks[W - 1] if W > 0 else t_start
    dt = marks[W + K - 1] - t0
    coll = sum(p[0] for p in parts[W:W + K]) / K
    learn = sum(p[1] for p in parts[W:W + K]) / K
    out = {
        "env_steps_per_sec": a.num_envs * T * K / dt,
        "env_steps_per_sec_reference_metric": a.num_envs * T / (coll + learn),   # on_policy_runner.py:199-203
        "ms_per_iteration": dt / K * 1e3, "collection_s": coll, "learn_s": learn,
        "num_envs": a.num_envs, "num_steps_per_env": T, "steps": K, "warmup": W, "device": a.device,
        "threads": torch.get_num_threads(), "reference_tree": ref_kind, "torch": torch.__version__,
        "loaded_product_so": any("libhg_b200" in line for line in open("/proc/self/maps")),
    }
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
