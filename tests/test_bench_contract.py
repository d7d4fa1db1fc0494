"""bench.py's reference arm and its line format, checked without a GPU.

The reference arm (`--impl reference`) times the C float64 port of the reference step (oracle/mobi_oracle.c) on the host
cores -- the one place outside tests/ and smoke() where oracle/ is executed; it must run on a box with no GPU and print
one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _json_line(out: str) -> dict:
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1"],
                         cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    d = _json_line(res.stdout)
    assert d["impl"] == "reference" and d["metric"] == "env-steps/sec" and d["unit"] == "env-steps/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("config[1]")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                         cwd=ROOT, capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    assert not [l for l in res.stdout.splitlines() if l.startswith("{")]


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    res = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True,
                         text=True, timeout=300)
    assert res.returncode != 0
    assert not [l for l in res.stdout.splitlines() if l.startswith("{")]      # no number from a CPU fallback
