"""Under torchrun: per-phase wall time (rollout / update) of the host-resident-frames arm on every rank."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench

local = int(os.environ.get("LOCAL_RANK", "0"))
rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = f"cuda:{local}"
for kind, graph in (("synthetic_host", "1"), ("synthetic_host", "0"), ("synthetic", "1")):
    os.environ["HG_CUDA_GRAPH"] = graph
    env, runner = bench._make_runner(4096, dev, kind)
    obs, cobs = env.get_observations(), env.get_privileged_observations()
    for i in range(5):
        torch.cuda.synchronize()
        t0 = time.time()
        with torch.inference_mode():
            obs, cobs = runner.collect(obs, cobs)
            torch.cuda.synchronize()
            t1 = time.time()
            runner.alg.update()
            torch.cuda.synchronize()
        t2 = time.time()
        if i >= 2:
            print(f"rank {rank} {kind} graph={graph} iter {i}: rollout {1e3 * (t1 - t0):.1f} ms  update {1e3 * (t2 - t1):.1f} ms", flush=True)
    del env, runner
    torch.cuda.empty_cache()
dist.destroy_process_group()
