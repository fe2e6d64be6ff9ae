"""CPU: bench.py's reference arm (the CPU port of the reference, the one place outside tests/
that may execute oracle/) prints exactly one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line(stories15m):
    # the default headline workload is llama2-7B (27 GB of host weights for the CPU arm): the contract is
    # checked on stories15M through the documented override
    env = dict(os.environ, L2B_BENCH_CPU_BUDGET_S="3", L2B_BENCH_WORKLOAD="stories15M")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["config"]["workload"] == "stories15M" and d["n_gpus"] == 1
    # same keys / values as the GPU arm's config, and the CLI's steps / warm-up echoed
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.bench_config("stories15M", 256, 1)
    assert d["steps"] == 2 and d["warmup"] == 1
    assert 2 <= d["sampled_positions_per_step"] <= 256
    assert abs(d["ms_per_step"] - 1e3 * 256 / d["value"]) < 1e-6 * d["ms_per_step"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 10 and d["gpu_launches"] == 0


def test_gpu_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True,
                       text=True, cwd=ROOT, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)


def test_default_headline_workload_is_the_same_at_every_gpu_count():
    """SCALE divides the N-GPU value by the 1-GPU value: both must be the same workload."""
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    os.environ.pop("L2B_BENCH_WORKLOAD", None)
    for n in (1, 2, 4, 8):
        assert bench.pick_workload(argparse.Namespace(workload="auto", gpus=n)) == "llama2-7B"
