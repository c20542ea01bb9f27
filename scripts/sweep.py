#!/usr/bin/env python
"""Tuning sweep (GPU): times gespmm_csr_spmm_f32 under GESPMM_VARIANT / GESPMM_TASK / GESPMM_LONG
overrides on one workload and checks every variant's result against the first one.
    python scripts/sweep.py --workload citpatents --K 128 --variants 0,1,2 --tasks 0 --longs 0
"""
import argparse
import itertools
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
    ap.add_argument("--workload", default="citpatents")
    ap.add_argument("--K", type=int, default=128)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--variants", default="0,1")
    ap.add_argument("--tasks", default="0")
    ap.add_argument("--longs", default="0")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--unvalued", action="store_true")
    ap.add_argument("--ref", action="store_true", help="also time the reference kernel")
    args = ap.parse_args()
    entry.load_package()
    from gespmm_b200 import capi, graphs
    from gespmm_b200.op import spmm
    dev = torch.device("cuda:0")
    rowptr, colind = bench.make_graph(args.workload, args.scale, dev)
    M = rowptr.numel() - 1
    nnz = colind.numel()
    B = graphs.cli_dense(M, args.K, seed=1, device=dev)
    val = None if args.unvalued else torch.ones(nnz, device=dev)
    flops = 2.0 * nnz * args.K
    first = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    print("# %s K=%d M=%d nnz=%d" % (args.workload, args.K, M, nnz), flush=True)
    for v, t, l in itertools.product(args.variants.split(","), args.tasks.split(","), args.longs.split(",")):
        os.environ["GESPMM_VARIANT"], os.environ["GESPMM_TASK"], os.environ["GESPMM_LONG"] = v, t, l
        capi.reload_env()  # the library reads its environment once
        run = (lambda: spmm.csr_spmm_no_edge_value(rowptr, colind, B)) if val is None else (lambda: spmm.csr_spmm(rowptr, colind, val, B))
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
            first = C.clone()
            same = "ref"
        else:
            same = "bitwise" if torch.equal(C, first) else "maxdiff=%.3g" % float((C - first).abs().max())
        print(json.dumps({"variant": int(v), "task": int(t), "long": int(l), "ms": round(ms, 4), "gflops": round(flops / ms / 1e6, 1), "vs_first": same}), flush=True)
        del C
    if args.ref:
        oracle = entry.load_oracle()
        L = oracle.ref_cli_kernels()
        ones = torch.ones(nnz, device=dev)
        Cr = torch.empty(M, args.K, device=dev)
        rms = L.ref_spmm_time_ms(2, 8, M, args.K, rowptr.data_ptr(), colind.data_ptr(), ones.data_ptr(), B.data_ptr(), Cr.data_ptr(), 3, args.iters)
        print(json.dumps({"reference_kernel": "spmm_test2 tile_row 8", "ms": round(rms, 4), "gflops": round(flops / rms / 1e6, 1),
                          "vs_first": "bitwise" if torch.equal(Cr, first) else "maxdiff=%.3g" % float((Cr - first).abs().max())}))


if __name__ == "__main__":
    main()
