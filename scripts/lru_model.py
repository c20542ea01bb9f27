#!/usr/bin/env python
"""How much DRAM traffic would an ideal LRU cache of the L2's size leave on the bench's cit-Patents shape?
Walks the graph's B-row access stream (a) in plain row order and (b) interleaved the way the kernel's 3552 resident
128-key tasks interleave it, through an exact LRU of C rows of 512 bytes (scripts/lru_model.c), and prints the misses.
CPU only:  python scripts/lru_model.py [--workload citpatents] [--task 128] [--resident 3552]
"""
import argparse
import os
import struct
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="citpatents")
    ap.add_argument("--task", type=int, default=128)
    ap.add_argument("--resident", type=int, default=148 * 24)
    ap.add_argument("--capacities", default="50000,100000,150000,200000,250000")
    args = ap.parse_args()
    entry.load_package()
    rp, ci = bench.make_graph(args.workload, 1.0, "cpu")
    rp, ci = rp.numpy().astype(np.int64), ci.numpy()
    M, nnz = len(rp) - 1, len(ci)
    tmp = tempfile.mkdtemp()
    exe = os.path.join(tmp, "lru")
    subprocess.check_call(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "scripts", "lru_model.c")])

    def run(cols, label):
        path = os.path.join(tmp, "stream.bin")
        with open(path, "wb") as f:
            f.write(struct.pack("<q", nnz)); f.write(struct.pack("<i", M)); f.write(np.ascontiguousarray(cols).tobytes())
        print("# " + label)
        sys.stdout.flush()
        subprocess.check_call([exe, path] + args.capacities.split(","))

    run(ci, "B rows in CSR (row) order, one row at a time")
    key = rp[:-1] + np.arange(M)
    task_of_nz = np.repeat(key // args.task, np.diff(rp))
    first = np.zeros(task_of_nz.max() + 2, dtype=np.int64)
    first[1:] = np.cumsum(np.bincount(task_of_nz, minlength=task_of_nz.max() + 1))
    pos = np.arange(nnz) - first[task_of_nz]
    order = np.lexsort((task_of_nz, pos, task_of_nz // args.resident))
    run(ci[order], "%d-key tasks, %d resident at a time, their nonzeros interleaved round-robin (the kernel's frontier)" % (args.task, args.resident))


if __name__ == "__main__":
    main()
