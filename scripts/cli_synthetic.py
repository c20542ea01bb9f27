#!/usr/bin/env python
"""BASELINE.json configs[1] through the CLI: write a synthetic shape-alike as a MatrixMarket file (the library's
parallel writer: 0.4 s for the 16.5 M entries of the cit-Patents shape), run bin/spmm_test on it exactly as
run_test.sh runs the reference's driver (`spmm_test <mtx> [dev]`, K = 128, 256, 512, 200 iterations, CSV cells
appended to spmm_test_out.out) with the reference's own kernel as the baseline cell, and print the CLI's JSON lines.
    python scripts/cli_synthetic.py [--workload citpatents|products|reddit|rmat] [--scale 1.0] [--dir DIR] [CLI flags ...]
GPU box.  The second run on the same DIR loads the parsed CSR image instead of re-parsing (--cache).
"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="citpatents")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--dir", default=os.path.join(ROOT, "gpurun_out"))
    args, cli_flags = ap.parse_known_args()
    entry.load_package()
    from gespmm_b200 import build, graphs
    oracle = entry.load_oracle()
    os.makedirs(args.dir, exist_ok=True)
    path = os.path.join(args.dir, "%s_s%g.mtx" % (args.workload, args.scale))
    if not os.path.exists(path):
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        t0 = time.time()
        rowptr, colind = bench.make_graph(args.workload, args.scale, dev)
        t1 = time.time()
        graphs.write_mtx(path, rowptr.cpu(), colind.cpu())
        print("# generated %s in %.1f s, wrote %s (%.0f MB) in %.1f s" % (args.workload, t1 - t0, path, os.path.getsize(path) / 1e6,
                                                                       time.time() - t1), file=sys.stderr)
    cmd = [build.CLI, path, "0", "--json", "--cache", "--out", os.path.join(args.dir, "spmm_test_out.out")]
    if oracle.have_ref(oracle.REF_CLI_KERNELS):
        cmd += ["--baseline-lib", oracle.REF_CLI_KERNELS]
    res = subprocess.run(cmd + cli_flags, stdout=subprocess.PIPE, text=True)
    print(res.stdout, end="")
    return res.returncode


if __name__ == "__main__":
    sys.exit(main())
