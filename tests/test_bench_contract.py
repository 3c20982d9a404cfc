"""bench.py's reference arm (`--impl reference`) honours the driver's contract without a GPU: one JSON line with
the agreed keys on rank 0, nothing (exit 0) on the other ranks.  (The GPU arm needs a B200: `-m gpu` / the driver.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *args):
    env = dict(os.environ, **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], env=env, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run({"RANK": "0", "WORLD_SIZE": "1"}, "--impl", "reference", "--workload", "det5", "--steps", "1",
             "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "DeformConv2d fwd+bwd images/sec" and d["unit"] == "images/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("det5")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"}, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0 and r.stdout.strip() == ""
