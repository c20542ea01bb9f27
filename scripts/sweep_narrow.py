#!/usr/bin/env python
"""Multi-shape sweep (GPU), one process, one JSON line per run.  Default: the ring walker (GESPMM_VARIANT=0,
sequential order) against the sub-warp walker (GESPMM_VARIANT=2) for K <= 64 over four graph shapes and task windows:
    python scripts/sweep_narrow.py [--workloads products,reddit,citpatents,rmat] [--Ks 16,32,64] [--tasks 0,256,512,1024]
K sweep of the shipped configuration next to the reference kernel (spmm_test2<float>, tile_row 8) on the same GPU:
    python scripts/sweep_narrow.py --workloads products --Ks 32,64,128,256,512 --variants -1 --tasks 0 --valued 1 --ref
Every run's result is compared with the first run's of its (workload, K): max |diff| relative to max |C|.
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
    ap.add_argument("--variants", default="0,2", help="GESPMM_VARIANT values; -1 = the library's automatic choice")
    ap.add_argument("--valued", default="1,0")
    ap.add_argument("--ref", action="store_true", help="also time the reference's spmm_test2<float> (oracle/_ref)")
    args = ap.parse_args()
    entry.load_package()
    from gespmm_b200 import capi, graphs
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
            for valued in (bool(int(x)) for x in args.valued.split(",")):
                for variant in (int(x) for x in args.variants.split(",")):
                    for task in args.tasks.split(","):
                        os.environ["GESPMM_VARIANT"], os.environ["GESPMM_TASK"] = str(variant), task
                        capi.reload_env()  # the library reads its environment once
                        torch.cuda.empty_cache()
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
            if args.ref and M * K < 2**31:
                L = entry.load_oracle().ref_cli_kernels()
                Cr = torch.empty(M, K, device=dev)
                # the kernel the reference itself would pick: the extension's K thresholds (spmm_kernel.cu:437-456) below 64,
                # the CLI's timed configuration (spmm_test.cu:724,756) from 64 on
                method, tile_row = (0, max(1, 128 // K)) if K < 32 else ((1, 4) if K < 64 else (2, 8))
                rms = L.ref_spmm_time_ms(method, tile_row, M, K, rowptr.data_ptr(), colind.data_ptr(), val.data_ptr(), B.data_ptr(), Cr.data_ptr(), 2, args.iters)
                rel = float((Cr - first).abs().max() / first.abs().max().clamp_min(1e-30))
                print(json.dumps({"workload": wl, "M": M, "nnz": nnz, "K": K, "reference_kernel": "spmm_test%d<float> tile_row %d" % (method, tile_row), "ms": round(rms, 4),
                                  "gflops": round(flops / rms / 1e6, 1), "max_rel_diff_vs_ring": rel}), flush=True)
                del Cr
            del B, first
        del rowptr, colind, val
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
