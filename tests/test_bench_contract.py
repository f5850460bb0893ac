"""The reference arm of bench.py runs without a GPU: its JSON line must keep the contract the
driver parses (metric/unit/impl/e2e/cpu_baseline keys; one line on stdout)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference",
                          "--gpus", "1", "--steps", "1", "--warmup", "1", "--cpu-batch", "32"],
                         capture_output=True, text=True, timeout=600, cwd=REPO,
                         env={**os.environ, "OMP_NUM_THREADS": "1"})       # as under torchrun
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("samples/sec")
    assert d["unit"] == "samples/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and "batch 32" in cb["sample"]
    assert d["gpu_launches"] == 0 and d["dtype"] == "f32" and d["vs_baseline"] is None


def test_other_ranks_of_the_reference_arm_exit_quietly():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=REPO,
                         env={**os.environ, "RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""
