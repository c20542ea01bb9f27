#!/usr/bin/env python
"""Narrow-B sweep (GPU): ring walker (GESPMM_VARIANT=0, sequential order) against the sub-warp walker
(GESPMM_VARIANT=2) for K <= 64 over several graph shapes and task windows, one process, one JSON line per run.
    python scripts/sweep_narrow.py [--workloads products,reddit,citpatents,rmat] [--Ks 16,32,64] [--tasks 0,256,512,1024]
Every variant's result is compared with the ring walker's: max |diff| relative to max |C| (re-association only).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="products,reddit,citpatents,rmat")
    ap.add_argument("--Ks", default="16,32,64")
    ap.add_argument("--tasks", default="0,256,512,1024")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--rmat-scale", type=float, default=0.5)
    args = ap.parse_args()
    entry.load_package()
    from gespmm_b200 import graphs
    from gespmm_b200.op import spmm
    dev = torch.device("cuda:0")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for wl in args.workloads.split(","):
        rowptr, colind = bench.make_graph(wl, args.rmat_scale if wl == "rmat" else 1.0, dev)
        M, nnz = rowptr.numel() - 1, colind.numel()
        val = torch.ones(nnz, device=dev)
        for K in (int(k) for k in args.Ks.split(",")):
            B = graphs.cli_dense(M, K, seed=1, device=dev)
            flops = 2.0 * nnz * K
            first = None
            for valued in (True, False):
                for variant in (0, 2):
                    for task in args.tasks.split(","):
                        os.environ["GESPMM_VARIANT"], os.environ["GESPMM_TASK"] = str(variant), task
                        run = (lambda: spmm.csr_spmm(rowptr, colind, val, B)) if valued else (lambda: spmm.csr_spmm_no_edge_value(rowptr, colind, B))
                        for _ in range(3):
                            C = run()
                        torch.cuda.synchronize()
                        e0.record()
                        for _ in range(args.iters):
                            C = run()
                        e1.record()
                        torch.cuda.synchronize()
                        ms = e0.elapsed_time(e1) / args.iters
                        if first is None:
                            first, rel = C.clone(), 0.0
                        else:
                            rel = float((C - first).abs().max() / first.abs().max().clamp_min(1e-30))
                        print(json.dumps({"workload": wl, "M": M, "nnz": nnz, "K": K, "valued": valued, "variant": variant, "task": int(task),
                                          "ms": round(ms, 4), "gflops": round(flops / ms / 1e6, 1), "gather_gbs": round(nnz * K * 4 / ms / 1e6, 1),
                                          "max_rel_diff_vs_ring": rel}), flush=True)
                        del C
            del B, first
        del rowptr, colind, val
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
