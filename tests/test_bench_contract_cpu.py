"""CPU: the driver-facing contract of bench.py that can be checked without a GPU -- the reference arm (`--impl reference` times the
oracle's fp32 restatement on the host cores) prints ONE JSON line with the agreed keys, and the product arm refuses to run without CUDA
instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "3"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("GGNN node-state-updates/sec") and d["unit"] == "node-updates/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["value"] - d["e2e"]["value"]) < 1e-6 * d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "cfg2" in cb["sample"]
    assert d["config"]["workload"].startswith("cfg2") and d["config"]["hidden"] == 100 and d["config"]["layer_timesteps"] == [4]
    # both arms print the SAME config dict (bench.config_of), so the driver can see they ran the same workload
    sys.path.insert(0, ROOT)
    import bench
    from gated_graph_neural_network_samples_b200 import workloads
    assert d["config"] == json.loads(json.dumps(bench.config_of(workloads.build("cfg2", seed=0), 1)))


def test_reference_arm_under_a_multi_rank_launch_runs_on_rank_zero_only():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    r = _run(["--steps", "1", "--warmup", "3"])
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
