#!/usr/bin/env python
"""Panel-width sweep (GPU): GESPMM_PANEL_V = 1, 2, 4 (columns per pass = 128 V) for wide B on one shape.
    python scripts/sweep_panel.py --workload reddit --Ks 256,512
A narrower panel walks A once per 128 V columns (colind re-read K / (128 V) times) but gathers from an
N x 128 V slice of B, which may fit the 126 MB L2 when all of B does not."""
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
    ap.add_argument("--workload", default="reddit")
    ap.add_argument("--Ks", default="256,512")
    ap.add_argument("--Vs", default="0,1,2,4")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    entry.load_package()
    from gespmm_b200 import capi, graphs
    from gespmm_b200.op import spmm
    dev = torch.device("cuda:0")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rowptr, colind = bench.make_graph(args.workload, args.scale, dev)
    M, nnz = rowptr.numel() - 1, colind.numel()
    val = torch.ones(nnz, device=dev)
    for K in (int(k) for k in args.Ks.split(",")):
        B = graphs.cli_dense(M, K, seed=1, device=dev)
        first = None
        for v in args.Vs.split(","):
            os.environ["GESPMM_PANEL_V"] = v
            capi.reload_env()  # the library reads its environment once
            for _ in range(3):
                C = spmm.csr_spmm(rowptr, colind, val, B)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.iters):
                C = spmm.csr_spmm(rowptr, colind, val, B)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            if first is None:
                first = C.clone()
            print(json.dumps({"workload": args.workload, "M": M, "nnz": nnz, "K": K, "panel_v": int(v), "ms": round(ms, 4),
                              "gflops": round(2.0 * nnz * K / ms / 1e6, 1), "bitwise_equal_to_first": bool(torch.equal(C, first))}), flush=True)
            del C
        del B, first


if __name__ == "__main__":
    main()
