"""bench.py's contract where it can be checked without a GPU: the reference arm (the reference's own CPU code through
oracle/_ref, or the restatement) prints ONE JSON line with the agreed keys and the same `config` the GPU arm names, and the GPU arm
refuses to run without a device instead of falling back to anything."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py")] + args, cwd=REPO, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                          text=True, timeout=timeout, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})


def test_reference_arm_prints_the_contract_line():
    r = run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [line for line in r.stdout.splitlines() if line.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "columns/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"].startswith("range-image columns/s") and d["value"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("synthetic 64-ring") and d["config"]["batch_firings"] == 4096
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gpu_arm_refuses_to_run_without_a_device():
    r = run(["--steps", "1", "--warmup", "3"])
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout)
    assert not [line for line in r.stdout.splitlines() if line.startswith("{")], "no bench line may be printed without a GPU"
