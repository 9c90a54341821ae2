"""bench.py's reference arm runs without a GPU (the C oracle port on the host cores): its JSON line must carry what the driver
reads — impl, metric / unit of the GPU arm, cpu_baseline describing the run, an e2e object without device traffic."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_reference(*extra, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3", *extra],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-500:]
    return [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line():
    lines = run_reference("--gpus", "1")
    assert len(lines) == 1
    b = lines[0]
    assert b["impl"] == "reference" and b["metric"] == "env-steps/sec (flow-field+dynamics)" and b["unit"] == "env-steps/s"
    assert b["higher_is_better"] is True and b["value"] > 0 and b["steps"] == 3 and b["gpu_launches"] == 0
    assert b["config"]["envs_per_gpu"] == 4096 and b["config"]["grid"] == 128 and b["config"]["window"] == 100
    cb = b["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == b["value"] and "median of" in cb["sample"]
    assert b["e2e"] == {"value": b["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert b["sample_envs"] == cb["cores"] * 32


def test_reference_arm_only_rank0_prints():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    assert run_reference("--gpus", "2", env=env) == []
