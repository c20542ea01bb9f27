"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm prints ONE
JSON line with the required keys; the own arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def _run(*args):
    env = dict(os.environ, OMP_NUM_THREADS="4")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT)


def test_reference_arm_json_line():
    res = _run("--impl", "reference", "--scale", "0.01", "--steps", "2", "--warmup", "1")
    assert res.returncode == 0, res.stderr
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "spmm_gflops" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True and d["dtype"] == "f32"
    assert d["value"] > 0 and d["steps"] == 2 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_only_rank0_prints():
    env_rank1 = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--scale", "0.01"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env_rank1, cwd=ROOT)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_own_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return
    res = _run("--steps", "3")
    assert res.returncode != 0 and "no CPU path" in (res.stderr + res.stdout)
