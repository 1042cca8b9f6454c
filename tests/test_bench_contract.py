"""bench.py's reference arm on the CPU (the b200 arm needs a device): one JSON line on stdout with the keys the driver reads, the same metric,
unit and config as the b200 arm, and the bounded-sample cpu_baseline; ranks other than 0 print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env)


def test_reference_arm_prints_one_json_line():
    p = run()
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.split("\n") if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gbp*guides/s" and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and abs(d["e2e"]["value"] - d["value"]) < 1e-12 and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "sample" in cb and cb["value"] == d["value"]
    assert d["config"]["workload"].startswith("SearchReference, 100 guides") and d["config"]["genome_bp"] > 3_000_000_000
    assert "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    p = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and "no CPU path" in (p.stderr + p.stdout)
