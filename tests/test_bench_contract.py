"""bench.py contract (CPU part): the reference arm prints ONE JSON line with the keys the driver reads, on the same
metric / unit as the GPU arm, and non-zero ranks of a torchrun launch exit without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "64", "--steps", "2",
                           "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env)


def test_reference_arm_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("3D r2c+c2r GFLOP/s") and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({"RANK": "3", "WORLD_SIZE": "8", "LOCAL_RANK": "3"})
    assert r.returncode == 0 and r.stdout.strip() == ""
