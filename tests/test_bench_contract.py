"""bench.py contract, the half that runs without a GPU: the reference arm (`--impl reference`) times the UNMODIFIED
reference (baseline/_ref or /root/reference over the ring-mode fake isaacgym; oracle port only as the stated fallback)
on the host cores and prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *flags):
    env = dict(os.environ)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-T", "2", "--num-envs", "64", *flags], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout.strip().splitlines()


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    lines = _run()
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "humanoid")) or os.path.isdir("/root/reference/humanoid")
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    if have_ref:
        assert d["config"]["num_steps_per_env"] == 60            # the full configuration, not a shortened sample
        assert d["loaded_product_so"] is False                    # the reference arm never maps libhg_b200.so
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_port_fallback_says_so():
    d = json.loads(_run({"HG_REF_FORCE_PORT": "1"})[0])
    assert d["cpu_baseline"]["kind"] == "port" and "FALLBACK" in d["cpu_baseline"]["sample"]
    assert d["config"]["num_steps_per_env"] == 2


def test_reference_arm_under_a_multi_rank_launch_runs_on_rank_zero_only():
    # ranks other than 0 exit 0 without output (the driver launches the arm through torchrun for N > 1)
    lines = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert lines == []
