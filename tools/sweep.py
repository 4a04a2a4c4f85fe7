#!/usr/bin/env python
"""BASELINE.json configs[4]: env-steps/s at num_envs in {1024, 4096, 16384, 65536} on this GPU, product vs the
unmodified reference's PyTorch path on the same GPU (R-GPU) -- writes a markdown table (stdout).

    python tools/sweep.py [--gpus-note "1 x B200"] > gpurun_out/sweep.md
(the multi-GPU columns come from `bench.py --gpus G --num-envs N` under torchrun: weak scaling, see SCALE_rNN.json)"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def product(N, steps, warmup):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--quick", "--num-envs", str(N), "--steps", str(steps), "--warmup", str(warmup)],
                       capture_output=True, text=True, timeout=1200, cwd=ROOT)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    if not lines:
        return {"error": p.stderr[-300:]}
    return json.loads(lines[-1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1024,4096,8192,16384,65536")
    ap.add_argument("--no-reference", action="store_true")
    a = ap.parse_args()
    print("| num_envs | product env-steps/s | ms / iteration | R-GPU env-steps/s (reference on the same B200) | ratio |")
    print("|---|---|---|---|---|")
    for N in [int(x) for x in a.sizes.split(",")]:
        steps, warm = (5, 3) if N <= 16384 else (3, 2)
        pr = product(N, steps, warm)
        if "error" in pr:
            print(f"| {N} | error: {pr['error'][-120:]} | | | |")
            continue
        ref = {"unavailable": "skipped"} if a.no_reference else bench._reference_subprocess("cuda:0", N, 2, 1, timeout=1500)
        if "unavailable" in ref:
            print(f"| {N} | {pr['value']:,.0f} | {pr['ms_per_step']:.2f} | n/a ({ref['unavailable'][:60]}) | |")
        else:
            print(f"| {N} | {pr['value']:,.0f} | {pr['ms_per_step']:.2f} | {ref['env_steps_per_sec']:,.0f} | {pr['value'] / ref['env_steps_per_sec']:.1f}x |")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
