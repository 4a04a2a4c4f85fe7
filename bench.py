#!/usr/bin/env python
"""bench.py -- env-steps/sec of the humanoid_ppo hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W             # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...    # the reference's PyTorch path on the host cores

A "step" is one learning iteration of `scripts/train.py --task=humanoid_ppo`: T=60 env steps of
num_envs=4096 envs per GPU (act -> env.step -> process_env_step), compute_returns, and PPO.update
(2 epochs x 4 minibatches), i.e. 245,760 env-steps per GPU.  metric = N*T*K*world / time, exactly
on_policy_runner.py:199-203 aggregated over ranks.  Physics is the seeded synthetic tensor source of
SURVEY.md section 8d (neither Isaac Gym nor MuJoCo exists in this image); its frames are pre-generated
outside the timed region.

Printed JSON (rank 0, one line): see README / DESIGN.md section 6.  `value` is timed with CUDA events around
exactly K steps with inputs resident in HBM; `e2e` is the same loop driven through the public
task_registry/OnPolicyRunner API with the physics frames in PINNED HOST memory (host->device copy of every
frame inside the timed region) and a device->host read of the iteration's losses and mean reward.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "humanoid-gym_b200")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _use_product():
    """Put the product package on sys.path (product arm only: the reference arm must resolve `humanoid` to the
    reference and must never map libhg_b200.so)."""
    if PKG not in sys.path:
        sys.path.insert(0, PKG)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

T_STEPS = 60
ENV_BYTES_PER_STEP = 7985            # algorithmic bytes of the fused post-physics kernel / env-step
FLOPS_FWD = 1052672 + 795392         # per sample, actor + critic forward
FLOPS_BWD = 1383424 + 1254400        # per sample, actor + critic backward


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, index=0):
        self.samples, self.reasons, self._stop, self.index = [], set(), threading.Event(), index
        self.max_mhz = None

    def _run_nvml(self):
        """NVML in-process: ~10 us per query, so even a 150 ms timed region gets dozens of samples."""
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))      # proves the path works
        while not self._stop.is_set():
            self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            r = int(get_reasons(h))
            for b, n in bits.items():
                if r & b:
                    self.reasons.add(n)
            self._stop.wait(0.005)

    def _run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.th.join(timeout=3)

    def summary(self):
        s = sorted(self.samples)
        return dict(sm_mhz=s[len(s) // 2] if s else None, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(s))


# ----------------------------------------------------------------------------------------------
# product arm
# ----------------------------------------------------------------------------------------------
def _make_runner(num_envs, device, physics, seed=5):
    os.environ["HG_PHYSICS"] = physics
    os.environ.setdefault("WANDB_MODE", "disabled")
    from humanoid.envs import XBotLCfg  # noqa: F401  (registers humanoid_ppo)
    from humanoid.utils import task_registry
    from humanoid.utils.helpers import get_args
    args = get_args(["--task=humanoid_ppo", "--headless", f"--num_envs={num_envs}", f"--sim_device={device}",
                     f"--rl_device={device}", f"--seed={seed}"])
    env, _ = task_registry.make_env("humanoid_ppo", args=args)
    runner, _ = task_registry.make_alg_runner(env, name="humanoid_ppo", args=args, log_root=None)
    env.episode_length_buf = torch.randint_like(env.episode_length_buf, high=int(env.max_episode_length))
    return env, runner


def _iterate(runner, state, book=None):
    obs, cobs = state
    with torch.inference_mode():
        obs, cobs = runner.collect(obs, cobs, book)  # CUDA-graph replay of the 60-step rollout when the physics allows
        vl, sl = runner.alg.update()
    return (obs, cobs), (vl, sl)


def _time_iterations(runner, state, steps, device, world, e2e=False, book=None):
    """CUDA-event time of exactly `steps` iterations, barrier + synchronize on both sides, max over ranks."""
    from humanoid import _native as nat
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = nat.launch_count() + getattr(runner, "replayed_launches", 0)
    w0 = time.time()
    ev0.record()
    results = []
    dbg = os.environ.get("HG_BENCH_DEBUG") == "1"
    for _ in range(steps):
        if dbg:
            torch.cuda.synchronize(device)
            t_a = time.time()
            with torch.inference_mode():
                o, c = runner.collect(*state)
                torch.cuda.synchronize(device)
                t_b = time.time()
                losses = runner.alg.update()
                torch.cuda.synchronize(device)
            state = (o, c)
            print(f"[bench dbg] rank {os.environ.get('RANK', '0')} e2e={e2e}: collect {1e3 * (t_b - t_a):.1f} ms, update {1e3 * (time.time() - t_b):.1f} ms, "
                  f"graph={'yes' if getattr(runner, '_graph', None) is not None else 'no'}", file=sys.stderr, flush=True)
        else:
            state, losses = _iterate(runner, state, book)
        if e2e:   # device->host read of the step's result: losses, mean reward, and what train.py's logger reads every iteration
            ep = book.drain_infos() if book is not None else None
            results.append((losses, float(runner.env.rew_buf.mean().item()), ep))
    ev1.record()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    wall = (time.time() - w0) * 1e3
    launches = nat.launch_count() + getattr(runner, "replayed_launches", 0) - n0
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), wall, launches, state


# dram__bytes_read.sum + dram__bytes_write.sum of post_physics_kernel from the committed `ncu --set full` captures
# (profiles/r02c_env_n*_ncu_full.md, the round-2 kernel).  At N=4096 the 16.8 MB the kernel writes are still in the 126 MB L2 when it
# ends, so only the reads reach DRAM inside the launch.
NCU_TRAFFIC_SOURCE = "profiles/r02c_env_n4096_ncu_full.md, profiles/r02c_env_n65536_ncu_full.md (ncu --set full, one launch)"
_NCU_ENV_TRAFFIC = {4096: 19809792 + 1536, 65536: 319186688 + 215517696}


def _ncu_traffic(num_envs):
    return _NCU_ENV_TRAFFIC.get(int(num_envs))


def _kernel_rooflines(runner, device, pk):
    """Live CUDA-event timing of the two kernels that bound the path: the MLP GEMM chain of one PPO
    minibatch (tensor roofline) and the fused post-physics env kernel (HBM roofline)."""
    from humanoid import _native as nat
    env, alg = runner.env, runner.alg
    N = env.num_envs
    out = {}
    # -- fused env kernel: time K back-to-back launches on the launching stream -------------------------
    st = torch.cuda.current_stream(device)
    reps = 20
    for _ in range(3):
        env._launch_post_physics(nat.PHASE_STEP_ALL)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(192 * 1024 * 1024 // 4, device=device)     # > 126 MB L2
    tot = 0.0
    for _ in range(reps):
        flush.zero_()                                              # L2 flush between timed launches
        e0.record(st)
        env._launch_post_physics(nat.PHASE_STEP_ALL)
        e1.record(st)
        torch.cuda.synchronize(device)
        tot += e0.elapsed_time(e1)
    t_env = tot / reps * 1e-3
    gbs = ENV_BYTES_PER_STEP * N / t_env / 1e9
    out["roofline_env"] = dict(num_envs=N, kernel="post_physics_kernel", bound="hbm", achieved=round(gbs, 1), peak=pk["hbm"], unit="GB/s",
                               frac=round(gbs / pk["hbm"], 4), traffic=_ncu_traffic(N), us_per_launch=round(t_env * 1e6, 2),
                               bytes_per_launch=ENV_BYTES_PER_STEP * N, peak_source=pk["src"], l2="flushed between launches",
                               traffic_source=NCU_TRAFFIC_SOURCE if _ncu_traffic(N) else None)
    # -- MLP fwd+bwd chain of one minibatch (inputs 237 MB > L2) -------------------------------------------
    s = alg.storage
    B = (s.num_envs * s.num_transitions_per_env) // alg.num_mini_batches
    idx = torch.randperm(s.num_envs * s.num_transitions_per_env, device=device)[:B]
    split = alg.use_split_path()
    mb = s.gather(idx, split=split)
    ac = alg.actor_critic
    flat = ac.flat_params()
    sp = nat.stream_ptr(device.index)

    if split:
        w = alg._scratch_split(B)
        ac.refresh_split()
        xs_a, xs_c = nat.Split.of(mb["obs_split"]), nat.Split.of(mb["priv_split"])

        def chain():
            ac.native_forward_split("actor", xs_a, w["mean"], w["hid_a"])
            ac.native_forward_split("critic", xs_c, w["value"], w["hid_c"])
            ac.native_backward_split("actor", xs_a, w["hid_a"], w["d_mean"], w["dhid_a"], alg._grad)
            ac.native_backward_split("critic", xs_c, w["hid_c"], w["d_value"], w["dhid_c"], alg._grad)
    else:
        w = alg._scratch(B)

        def chain():
            ac.native_forward("actor", mb["obs"], w["mean"], hidden=w["hid_a"])
            ac.native_forward("critic", mb["priv_obs"], w["value"], hidden=w["hid_c"])
            g = alg._grad.data_ptr()
            nat.check(nat.lib.hg_mlp_backward(ac._desc["actor"], flat.data_ptr(), mb["obs"].data_ptr(), mb["obs"].stride(0), w["hid_a"].data_ptr(),
                                              w["d_mean"].data_ptr(), w["dhid_a"].data_ptr(), g, B, sp))
            nat.check(nat.lib.hg_mlp_backward(ac._desc["critic"], flat.data_ptr(), mb["priv_obs"].data_ptr(), mb["priv_obs"].stride(0), w["hid_c"].data_ptr(),
                                              w["d_value"].data_ptr(), w["dhid_c"].data_ptr(), g, B, sp))
    w["d_mean"].normal_()
    w["d_value"].normal_()
    for _ in range(3):
        chain()
    torch.cuda.synchronize(device)
    e0.record(st)
    reps = 5
    for _ in range(reps):
        chain()
    e1.record(st)
    torch.cuda.synchronize(device)
    t = e0.elapsed_time(e1) / reps * 1e-3
    flops = (FLOPS_FWD + FLOPS_BWD) * B
    tf = flops / t / 1e12
    mode = nat.lib.hg_set_gemm_mode(-1)
    # the chain is timed in isolation (5 repetitions, a few ms) -> MEASURED_PEAKS' BURST cuBLAS bf16 figure is the denominator;
    # kind::tf32 runs at half the bf16 rate
    if split:
        peak, passes, kind = pk["bf16"], 3, "bf16"
        engine = "gemm_bf3_kernel (tcgen05 kind::f16 on pre-split bf16 hi/lo operands, 3 MMAs per product) + head kernels"
    else:
        peak = pk["bf16"] / 2.0
        passes, kind = {0: 0, 1: 3, 2: 1, 4: 3}[mode], "tf32"
        engine = {0: "gemm_kernel (exact-fp32 CUDA-core path)", 1: "gemm_tc_kernel (tcgen05 kind::tf32, 3xTF32)",
                  2: "gemm_tc_kernel (tcgen05 kind::tf32, 1 pass)", 4: "gemm_tc_kernel (tcgen05 kind::tf32, 3xTF32)"}[mode]
    out["roofline"] = dict(kernel="ActorCritic fwd+bwd GEMM chain of one 61,440-sample minibatch: " + engine, bound="tensor",
                           achieved=round(tf, 2), peak=round(peak, 1), unit="TFLOP/s", frac=round(tf / peak, 4),
                           traffic=None, ms_per_minibatch=round(t * 1e3, 3), flops_per_launch_group=flops,
                           mma_passes=passes, mma_kind=kind, tensor_pipe_tflops=round(tf * max(passes, 1), 2),
                           tensor_pipe_frac=round(tf * max(passes, 1) / peak, 4),
                           note="achieved = ALGORITHMIC fp32 FLOPs (4.486 MFLOP/sample, SURVEY 8d) / time; the split-precision scheme "
                                "issues `mma_passes` tensor MMAs of kind `mma_kind` per algorithmic product, so the tensor pipe itself "
                                "runs at tensor_pipe_tflops; frac is the algorithmic rate over the dense peak of that kind",
                           peak_source=pk["src"] + f"; burst cuBLAS bf16 figure (chain timed in isolation){'' if split else ' / 2 for kind::tf32'}")
    return out


def env_roofline_large(device, pk, N=65536):
    """The fused env kernel at the size where it is bandwidth- rather than latency-bound (BASELINE.json sweep config)."""
    from humanoid import _native as nat
    env, _ = _make_env_only(N, str(device))
    st = torch.cuda.current_stream(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(192 * 1024 * 1024 // 4, device=device)
    for _ in range(3):
        env.step(torch.randn(N, 12, device=device))
    tot, reps = 0.0, 10
    for _ in range(reps):
        flush.zero_()
        e0.record(st)
        env._launch_post_physics(nat.PHASE_STEP_ALL)
        e1.record(st)
        torch.cuda.synchronize(device)
        tot += e0.elapsed_time(e1)
    t = tot / reps * 1e-3
    gbs = ENV_BYTES_PER_STEP * N / t / 1e9
    return dict(kernel="post_physics_kernel", num_envs=N, bound="hbm", achieved=round(gbs, 1), peak=pk["hbm"], unit="GB/s",
                frac=round(gbs / pk["hbm"], 4), traffic=_ncu_traffic(N), us_per_launch=round(t * 1e6, 2),
                bytes_per_launch=ENV_BYTES_PER_STEP * N, peak_source=pk["src"], l2="flushed between launches",
                traffic_source=NCU_TRAFFIC_SOURCE if _ncu_traffic(N) else None)


def _make_env_only(num_envs, device, seed=5):
    os.environ["HG_PHYSICS"] = "synthetic"
    from humanoid.envs import XBotLCfg  # noqa: F401
    from humanoid.utils import task_registry
    from humanoid.utils.helpers import get_args
    args = get_args(["--task=humanoid_ppo", "--headless", f"--num_envs={num_envs}", f"--sim_device={device}",
                     f"--rl_device={device}", f"--seed={seed}"])
    return task_registry.make_env("humanoid_ppo", args=args)


def run_product(args):
    _use_product()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    device = torch.device("cuda", local)
    pk = peaks()
    N = args.num_envs

    # ---- device-resident arm (value) -----------------------------------------------------------------
    env, runner = _make_runner(N, str(device), "synthetic")
    state = (env.get_observations(), env.get_privileged_observations())
    for _ in range(args.warmup):
        state, _ = _iterate(runner, state)
    with ClockSampler(local) as clocks:
        ms, wall_ms, launches, state = _time_iterations(runner, state, args.steps, device, world)
    env_steps = N * T_STEPS * args.steps * world
    value = env_steps / (ms * 1e-3)
    print(f"[bench] rank {rank}: {value:.0f} env-steps/s, {ms / args.steps:.2f} ms/step, host wall {wall_ms / args.steps:.2f} ms/step",
          file=sys.stderr, flush=True)
    if args.quick:
        if rank == 0:
            print(json.dumps({"metric": "env_steps_per_sec", "value": round(value, 1), "unit": "env-steps/s", "n_gpus": world,
                              "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "quick": True,
                              "gpu_launches": int(launches // args.steps)}))
        return
    _log("value arm done; e2e arm (host-resident physics frames)")
    del runner, env
    torch.cuda.empty_cache()

    # ---- end-to-end arm (host-resident physics frames, D2H of results) ---------------------------------
    # (runs before the rank-0-only kernel rooflines, so that every rank enters its warm-up and timed region together)
    env, runner = _make_runner(N, str(device), "synthetic_host")
    state = (env.get_observations(), env.get_privileged_observations())
    # the e2e arm runs what `scripts/train.py` runs with a log dir: per-step episode bookkeeping (one native launch per env
    # step, inside the rollout graph) and the per-iteration read-back of finished episodes + the 22 reward-term means
    from humanoid.algo.ppo.on_policy_runner import _EpisodeBook
    book = _EpisodeBook(env.num_envs, T_STEPS, len(env.extras.get("episode", {})), str(device))
    for _ in range(max(1, args.warmup)):
        state, _ = _iterate(runner, state, book)
        book.drain_infos()
    n_done0 = len(book.rewbuffer)
    ms_e, _, _, state = _time_iterations(runner, state, args.steps, device, world, e2e=True, book=book)
    e2e_value = env_steps / (ms_e * 1e-3)
    _log("e2e arm done; kernel rooflines")
    h2d = env.gym.h2d_bytes_per_step() * T_STEPS
    finished = float(torch.isfinite(book.done_rew).sum().item())          # finished episodes of the last iteration
    d2h = 8 * 4 + 4 + int(2 * 4 * finished) + 4 * len(env.extras.get("episode", {}))
    extra = {}
    if rank == 0:
        try:
            with torch.inference_mode():
                extra = _kernel_rooflines(runner, device, pk)
        except Exception as e:               # never leave the other ranks waiting at the barrier below
            extra = {"roofline_error": f"{type(e).__name__}: {e}"}
    _log("rooflines done")
    del runner, env
    if world > 1:
        dist.barrier()                       # the other ranks wait here for rank 0's roofline timings: orderly teardown

    if rank != 0:
        return
    line = {
        "metric": "env_steps_per_sec", "value": round(value, 1), "unit": "env-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "humanoid_ppo XBot-L, %d envs/GPU x T=60 steps + PPO update (2 epochs x 4 minibatches) per step"
                               % N, "num_envs_per_gpu": N, "num_steps_per_env": T_STEPS, "global_envs": N * world,
                   "parallelism": f"env-sharded dp{world}, one NCCL all-reduce of the flat gradient per optimizer step",
                   "physics": "synthetic tensor source (SURVEY.md 8d), frames pre-generated outside the timed region",
                   "l2": "per-step working set (rollout storage 947 MB + minibatch 237 MB) exceeds the 126 MB L2"},
        "e2e": {"value": round(e2e_value, 1), "unit": "env-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e / args.steps, 3),
                "note": "physics frames staged from pinned host memory every env step through a double buffer (the H2D of step s+1 overlaps step s; memcpy nodes of the rollout graph); per-step episode bookkeeping on (the log_dir path of train.py); losses, mean reward, finished-episode rewards / lengths and the 22 reward-term means read back every iteration"},
        "gpu_launches": int(launches // args.steps),
        "clocks": clocks.summary(),
        "host_wall_ms_per_step": round(wall_ms / args.steps, 3),
    }
    line.update(extra)
    if world == 1 and not args.no_env_sweep:
        _log("env kernel at N=65536 (its bandwidth regime)")
        line["roofline_env_65536"] = env_roofline_large(device, pk)
    if world == 1 and not args.no_cpu_baseline:
        torch.cuda.empty_cache()
        if not args.no_ref_gpu:
            _log("R-GPU: the unmodified reference on cuda:0 (subprocess)")
            line["reference_gpu"] = reference_gpu(N)
            if "value" in line["reference_gpu"]:
                line["vs_reference_gpu"] = {"value_ratio": round(value / line["reference_gpu"]["value"], 2),
                                            "e2e_ratio": round(e2e_value / line["reference_gpu"]["value"], 2),
                                            "target": ">= 10 (north_star)"}
        _log("cpu_baseline: the unmodified reference on the host cores (subprocess)")
        line["cpu_baseline"] = cpu_baseline(N, sample_T=args.cpu_T)
        try:
            line["cpu_baseline_sim2sim"] = sim2sim_cpu()
        except Exception as e:
            line["cpu_baseline_sim2sim"] = {"unavailable": f"{type(e).__name__}: {e}"}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference's PyTorch path (tools/reference_arm.py, a clean subprocess
# that never imports the product), on the host cores and -- as the ">= 10x" denominator -- on the same B200
# ----------------------------------------------------------------------------------------------
def _log(msg):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def _cpu_threads():
    """Host threads for the reference PyTorch path.  Its tensors are small (N x 12 .. N x 705 fp32), so intra-op
    parallelism stops paying long before a 100+-core host is full; more threads only add fork/join overhead."""
    return int(os.environ.get("HG_REF_THREADS", min(os.cpu_count() or 1, 32)))


def _reference_subprocess(device, num_envs, steps, warmup, timeout=1500):
    """Run tools/reference_arm.py (the unmodified reference over the ring-mode fake gym).  Returns its dict, or
    {"unavailable": why}."""
    if os.environ.get("HG_REF_FORCE_PORT") == "1":
        return {"unavailable": "HG_REF_FORCE_PORT=1"}
    cmd = [sys.executable, os.path.join(ROOT, "tools", "reference_arm.py"), "--device", device, "--num-envs", str(num_envs),
           "--steps", str(steps), "--warmup", str(warmup), "--threads", str(_cpu_threads())]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    env["PYTHONPATH"] = ""
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    except subprocess.TimeoutExpired:
        return {"unavailable": f"reference arm on {device} exceeded {timeout}s"}
    lines = [ln for ln in p.stdout.strip().splitlines() if ln.startswith("{")]
    if p.returncode != 0 or not lines:
        return {"unavailable": f"reference arm on {device} failed (rc {p.returncode}): {p.stderr.strip()[-300:]}"}
    return json.loads(lines[-1])


def _oracle_port(num_envs, T):
    """Fallback when no reference tree is on the box: the oracle restatement (kind "port")."""
    _use_product()
    from humanoid.physics import SyntheticPhysics
    from oracle.runner_oracle import OracleTrainer
    from oracle import env_oracle as eo
    torch.set_num_threads(_cpu_threads())
    ranges = {"lin_vel_x": [-0.3, 0.6], "lin_vel_y": [-0.3, 0.3]}
    ph = SyntheticPhysics(num_envs, "cpu", ranges, eo.grid_origins(num_envs), decimation=10, seed=5)
    return OracleTrainer(num_envs, ph, T=T)


def _ref_sample_text(r, what):
    return (f"{what}: {r['steps']} learning iteration(s) after {r['warmup']} warm-up of the UNMODIFIED reference "
            f"(OnPolicyRunner.learn via task_registry, {r['reference_tree']}) over the ring-mode fake isaacgym, "
            f"N={r['num_envs']}, T={r['num_steps_per_env']}, 2 epochs x 4 minibatches, torch {r['torch']} fp32 on {r['device']}, "
            f"{r['threads']} host threads; collection {r['collection_s']:.2f}s + learn {r['learn_s']:.2f}s per iteration")


def reference_gpu(num_envs, steps=3, warmup=2):
    """R-GPU: the reference's PyTorch path on the same B200 -- the denominator of north_star's '>= 10x'."""
    r = _reference_subprocess("cuda:0", num_envs, steps, warmup)
    if "unavailable" in r:
        return r
    return {"value": round(r["env_steps_per_sec"], 1), "unit": "env-steps/s", "device": "cuda:0", "kind": "reference",
            "ms_per_step": round(r["ms_per_iteration"], 2), "collection_s": round(r["collection_s"], 4),
            "learn_s": round(r["learn_s"], 4), "loaded_product_so": r["loaded_product_so"],
            "sample": _ref_sample_text(r, "R-GPU"),
            "note": "Isaac Gym cannot run on sm_100, so the reference's env+PPO torch code runs over the same synthetic frames "
                    "as the product (TF32 off); torch_utils helpers are eager here (Isaac Gym ships them torch.jit.script-ed)"}


def cpu_baseline(num_envs, sample_T):
    r = _reference_subprocess("cpu", num_envs, steps=1, warmup=1)
    if "unavailable" not in r:
        return {"value": round(r["env_steps_per_sec"], 1), "unit": "env-steps/s", "cores": r["threads"], "kind": "reference",
                "sample": _ref_sample_text(r, "cpu_baseline"), "collection_s": round(r["collection_s"], 3),
                "learn_s": round(r["learn_s"], 3), "loaded_product_so": r["loaded_product_so"]}
    _log(f"cpu_baseline: reference tree unavailable ({r['unavailable']}); oracle port, N={num_envs}, T={sample_T}")
    tr = _oracle_port(num_envs, sample_T)
    c, l = tr.iteration()
    v = num_envs * sample_T / (c + l)
    return {"value": round(v, 1), "unit": "env-steps/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"1 learning iteration of the oracle port (no reference tree on this box: {r['unavailable']}), "
                      f"N={num_envs}, T={sample_T}, collection {c:.2f}s + learn {l:.2f}s, torch CPU fp32",
            "collection_s": round(c, 3), "learn_s": round(l, 3)}


def sim2sim_cpu(calls=6000):
    """BASELINE.json configs[0] (C1): the reference's sim2sim deployment loop on the host -- 100 Hz observation assembly,
    15-frame stack, TorchScript actor, PD law at 1 kHz (reference scripts/sim2sim.py:113-160) -- with the reference's
    shipped policy (weights from the committed KAT fixture).  MuJoCo is not installable here: mj_step is not run."""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("hg_sim2sim", os.path.join(PKG, "humanoid", "scripts", "sim2sim.py"))
    s2s = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(s2s)
    k = np.load(os.path.join(ROOT, "tests", "golden", "policy_example_kat.npz"))
    layers = []
    for i in (0, 2, 4, 6):
        w, b = torch.from_numpy(k[f"w.{i}.weight"]), torch.from_numpy(k[f"w.{i}.bias"])
        lin = torch.nn.Linear(w.shape[1], w.shape[0])
        lin.weight.data.copy_(w), lin.bias.data.copy_(b)
        layers += [lin] + ([torch.nn.ELU()] if i < 6 else [])
    policy = torch.jit.script(torch.nn.Sequential(*layers))
    threads = torch.get_num_threads()
    torch.set_num_threads(1)                             # a batch-1 MLP: one core, like the robot's control thread
    try:
        s2s.run(policy, low_level_steps=500)             # warm-up
        n, sec = s2s.run(policy, low_level_steps=calls * 10)
    finally:
        torch.set_num_threads(threads)
    return {"value": round(n / sec, 1), "unit": "policy calls/s (1 env)", "cores": 1, "kind": "reference-shaped loop",
            "ms_per_call": round(1e3 * sec / n, 4), "calls": n,
            "sample": f"{n} policy calls (60 s of 100 Hz control): obs assembly + 15-frame stack + TorchScript actor + PD law "
                      f"at 1 kHz; MuJoCo unavailable (mujoco==2.3.6 not installable): mj_step was not run"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N = args.num_envs
    r = _reference_subprocess("cpu", N, args.steps, args.warmup)
    if "unavailable" in r:
        # no reference tree travelled to this box: time the oracle restatement instead and SAY so
        T = args.ref_T
        tr = _oracle_port(N, T)
        for _ in range(args.warmup):
            tr.iteration()
        t0 = time.time()
        c = l = 0.0
        for _ in range(args.steps):
            a, b = tr.iteration()
            c, l = c + a, l + b
        dt = time.time() - t0
        v, ms, cores, kind = N * T * args.steps / dt, dt / args.steps * 1e3, torch.get_num_threads(), "port"
        c, l = c / args.steps, l / args.steps
        sample = (f"FALLBACK ({r['unavailable']}): oracle port of the reference PyTorch path, N={N}, T={T} of 60 env steps "
                  f"per iteration, 2 epochs x 4 minibatches, {cores} host threads")
        so = None
    else:
        v, ms, cores, kind, T = r["env_steps_per_sec"], r["ms_per_iteration"], r["threads"], "reference", r["num_steps_per_env"]
        c, l, so = r["collection_s"], r["learn_s"], r["loaded_product_so"]
        sample = _ref_sample_text(r, "each step = one full learning iteration")
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": round(v, 1), "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "humanoid_ppo XBot-L, %d envs x T=%d steps + PPO update (2 epochs x 4 minibatches) per step; "
                               "reference PyTorch path on the host CPU" % (N, T), "num_envs_per_gpu": N, "num_steps_per_env": T,
                   "physics": "synthetic tensor source (SURVEY.md 8d), same generator and ring as the product arm"},
        "cpu_baseline": {"value": round(v, 1), "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(v, 1), "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "collection_s": round(c, 3), "learn_s": round(l, 3), "loaded_product_so": so}
    if args.ref_gpu and torch.cuda.is_available():
        line["reference_gpu"] = reference_gpu(N)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--num-envs", type=int, default=4096, help="envs per GPU (BASELINE.json configs[1])")
    ap.add_argument("--ref-T", type=int, default=12, help="env steps per iteration of the oracle-port FALLBACK (no reference tree on the box)")
    ap.add_argument("--ref-gpu", action="store_true", help="reference arm: also time the reference on cuda:0 (R-GPU)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="product arm: skip the R-GPU measurement")
    ap.add_argument("--cpu-T", type=int, default=12, help="env steps of the cpu_baseline sample (one iteration)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-env-sweep", action="store_true")
    ap.add_argument("--quick", action="store_true", help="value arm only (for ncu launch lists): no e2e, rooflines, cpu baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
