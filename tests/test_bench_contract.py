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


def test_committed_bench_line_carries_the_whole_contract():
    """profiles/r01_bench_citpatents_n1.json (round 1's line; round 2's: test_round2_bench_lines below) is the own arm's line as printed on a B200 by the final kernels of round 1:
    every key of the bench contract is there and the derived figures are consistent with each other."""
    with open(os.path.join(ROOT, "profiles", "r01_bench_citpatents_n1.json")) as f:
        d = json.loads(f.read())
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["metric"] == "spmm_gflops" and d["unit"] == "GFLOP/s" and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] >= d["steps"]
    cfg = d["config"]
    assert "cit-Patents" in cfg["workload"] and cfg["K"] == 128 and cfg["nnz"] == 16_518_948 and "larger than L2" in cfg["l2"]
    flops = 2.0 * cfg["nnz"] * cfg["K"]
    assert abs(d["value"] - flops / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["bytes_min"] / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert r["traffic"] is None or r["traffic"] > r["bytes_min"]   # ncu DRAM bytes per launch vs algorithmic bytes
    e = d["e2e"]
    assert e["unit"] == "GFLOP/s" and 0 < e["value"] < d["value"]
    assert e["h2d_bytes_per_step"] >= 4 * cfg["N"] * cfg["K"] and e["d2h_bytes_per_step"] == 4 * cfg["M"] * cfg["K"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    assert d["vs_baseline"] == d["value"] / 140.4
    ref = d["reference_kernel_same_gpu"]
    assert ref["bitwise_equal_to_ours"] is True and ref["ms"] > d["ms_per_step"]


def test_in_run_parity_check_accepts_the_oracle_and_catches_a_wrong_row(pkg, oracle):
    """bench.check_parity (what every rank runs on its block of C before the line is printed): a C that IS the oracle's
    result passes with every sampled row compared bit for bit; one flipped bit in one sampled row is reported."""
    import numpy as np
    import bench
    from gespmm_b200 import capi, graphs
    rp, ci = graphs.rmat(N=4000, nnz=60000, seed=3)
    val = torch.rand(ci.numel(), generator=torch.Generator().manual_seed(1)) - 0.5
    K = 128
    B = graphs.cli_dense(4000, K, seed=2)
    C = torch.from_numpy(oracle.spmm(rp.numpy(), ci.numpy(), val.numpy(), B.numpy(), fma=True))
    p = bench.check_parity(oracle, capi, rp, ci, val, B, C, K, sample=500, seed=5)
    assert p["rows_bad"] == 0 and p["rows_checked"] >= 400 and p["rows_bitwise"] == p["rows_checked"] and p["max_rel"] == 0.0
    deg = (rp[1:] - rp[:-1])
    hub = int(torch.argmax(deg))           # the longest row is always in the sample
    C2 = C.clone()
    C2[hub, 7] = float(np.nextafter(np.float32(C2[hub, 7]), np.float32(np.inf)))
    assert bench.check_parity(oracle, capi, rp, ci, val, B, C2, K, sample=500, seed=5)["rows_bad"] == 1
    # unvalued, and a width whose default walker re-associates (only rows of <= 1 nonzero are compared bit for bit)
    Cu = torch.from_numpy(oracle.spmm(rp.numpy(), ci.numpy(), None, B[:, :32].contiguous().numpy()))
    p = bench.check_parity(oracle, capi, rp, ci, None, B[:, :32].contiguous(), Cu, 32, sample=500, seed=5)
    assert p["rows_bad"] == 0 and 0 < p["rows_bitwise"] < p["rows_checked"]


def test_roofline_bytes_models():
    import bench
    assert bench.bytes_min(10, 20, 4, 30, valued=True) == 4 * 11 + 8 * 30 + 4 * 20 * 4 + 4 * 10 * 4
    assert bench.bytes_min(10, 20, 4, 30, valued=False) == 4 * 11 + 4 * 30 + 4 * 20 * 4 + 4 * 10 * 4

    class Shard:
        row_lo, row_hi, nnz_local = 5, 15, 30
    # a rank is charged for the B rows its block references, not for all of B
    assert bench.rank_bytes(Shard, 4, True, distinct=7) == 4 * 11 + 8 * 30 + 4 * 4 * 7 + 4 * 10 * 4
    assert bench.host_threads() >= 1


def test_round2_bench_lines():
    """profiles/r02_bench_n{1,2,4,8}_builder*.json: the own arm's lines as printed on B200s by the final code of round 2
    (builder runs; the driver's own runs are BENCH_r02 / SCALE_r02).  What round 1's verdict asked of them: parity checked
    inside the run at every N, the R-MAT 10M/200M record with its speed-up on the N > 1 lines, a roofline fraction that stays
    below 1, host-to-device bytes that do not grow with N, the uniformly random variant beside the N = 1 headline."""
    lines = {}
    for n, name in ((1, "r02_bench_n1_builder.json"), (2, "r02_bench_n2_builder.json"), (4, "r02_bench_n4_builder.json"),
                    (8, "r02_bench_n8_builder.json")):
        with open(os.path.join(ROOT, "profiles", name)) as f:
            lines[n] = json.loads(f.read())
    for n, d in lines.items():
        assert BASE_KEYS <= set(d) and d["n_gpus"] == n and d["metric"] == "spmm_gflops"
        assert d["parity"]["ok"] is True and d["parity"]["rows_bad"] == 0 and d["parity"]["rows_checked"] >= 1500 * n
        assert d["parity"]["rows_bitwise_vs_oracle"] == d["parity"]["rows_checked"]          # K = 128: every sampled row bit for bit
        assert len(d["batches_ms_per_step"]) == 5 and min(d["batches_ms_per_step"]) <= d["ms_per_step"] <= max(d["batches_ms_per_step"])
        r = d["roofline"]
        assert r["bound"] == "hbm" and 0.3 < r["frac"] < 1.0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert (r["traffic"] is None) == (n > 1)                                              # ncu bytes only where they were measured
        assert d["gpu_launches_per_step"] == 1                                                # no long rows: the long-row kernel is skipped
        e = d["e2e"]
        assert 2.0e9 < e["h2d_bytes_per_step"] < 2.2e9 and e["d2h_bytes_per_step"] == 4 * d["config"]["M"] * d["config"]["K"]
        assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    assert lines[1]["e2e"]["value"] < lines[2]["e2e"]["value"] and lines[1]["e2e"]["value"] < lines[8]["e2e"]["value"]
    assert lines[1]["value"] < lines[2]["value"] < lines[4]["value"] < lines[8]["value"]
    assert "uniform_variant" in lines[1] and lines[1]["uniform_variant"]["rows_bad"] == 0
    assert lines[1]["uniform_variant"]["roofline_frac"] < lines[1]["roofline"]["frac"]
    for n in (2, 4, 8):
        m = lines[n]["rmat"]
        assert m["n_gpus"] == n and m["parity"]["ok"] and m["parity"]["rows_bad"] == 0 and m["b_replicas_identical"]
        assert abs(m["speedup"] - m["ms_1gpu"] / m["ms_per_step"]) < 1e-9 and m["roofline"]["frac"] < 1.0
        assert m["b_allgather_ms"] < m["b_broadcast_ms"]
    assert lines[4]["rmat"]["speedup"] >= 3.5                                                 # BASELINE.json north_star
    assert lines[2]["rmat"]["speedup"] > 1.8 and lines[8]["rmat"]["speedup"] > 6.5
