#!/usr/bin/env python
"""Hot-column L2 pinning sweep (GPU): the H most referenced rows of B are gathered evict_last, the others evict_first
(bulk walker, gespmm_opts.hot_columns), on a skewed graph (default: R-MAT 10M/200M, K = 128, valued A == 1).
    python scripts/sweep_hot.py [--workload rmat] [--scale 1.0] [--hots 0,50000,100000,150000,200000] [--policies 22,18,20]
    python scripts/sweep_hot.py --one 5,22,100000      # walker,policy,H: 5 launches, for ncu --metrics dram__bytes...
One JSON line per configuration (median / min ms of --batches batches of --iters launches).
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


def hot_bitmap(colind, N, H):
    """uint32 bitmap of the H columns with the most nonzeros, and the share of the gathers they receive."""
    deg = torch.bincount(colind.long(), minlength=N)
    if H <= 0:
        return None, 0.0
    top = torch.topk(deg, min(H, N)).indices
    bits = torch.zeros(((N + 31) // 32) * 32, dtype=torch.int64, device=colind.device)
    bits[top] = 1
    words = (bits.view(-1, 32) << torch.arange(32, device=colind.device)).sum(1)
    words = torch.where(words >= 2**31, words - 2**32, words).to(torch.int32)
    return words.contiguous(), float(deg[top].sum()) / float(colind.numel())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="rmat")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--K", type=int, default=128)
    ap.add_argument("--hots", default="50000,100000,150000,200000")
    ap.add_argument("--policies", default="22,18,20")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--batches", type=int, default=3)
    ap.add_argument("--one", default=None)
    args = ap.parse_args()
    entry.load_package()
    from gespmm_b200 import capi, graphs
    dev = torch.device("cuda:0")
    rowptr, colind = bench.make_graph(args.workload, args.scale, dev)
    M, nnz, K = rowptr.numel() - 1, colind.numel(), args.K
    val = torch.ones(nnz, device=dev)
    B = graphs.cli_dense(M, K, seed=1, device=dev)
    C = torch.empty(M, K, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    maps = {}

    def run(walker, policy, H):
        if H not in maps:
            maps[H] = hot_bitmap(colind, M, H)
        hot, _ = maps[H]
        o = capi.opts(walker=walker, l2_policy=policy, l2_window_rows=-1, hot_columns=None if hot is None else hot.data_ptr())
        capi.csr_spmm_f32_ex(M, M, K, nnz, rowptr.data_ptr(), colind.data_ptr(), val.data_ptr(), B.data_ptr(), K, C.data_ptr(), K, o, st)

    if args.one:
        w, p, H = (int(x) for x in args.one.split(","))
        for _ in range(5):
            run(w, p, H)
        torch.cuda.synchronize()
        return
    ref = None
    configs = [(0, 0, 0), (5, 0, 0)] + [(5, p, H) for H in (int(x) for x in args.hots.split(",")) for p in (int(x) for x in args.policies.split(","))]
    for walker, policy, H in configs:
        for _ in range(2):
            run(walker, policy, H)
        torch.cuda.synchronize()
        times = []
        for _ in range(args.batches):
            e0.record()
            for _ in range(args.iters):
                run(walker, policy, H)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / args.iters)
        if ref is None:
            ref = C.clone()
        times.sort()
        print(json.dumps({"workload": args.workload, "K": K, "walker": walker, "policy": policy, "hot_columns": H,
                          "hot_share_of_gathers": round(maps[H][1], 4), "ms_median": round(times[len(times) // 2], 4),
                          "ms_min": round(times[0], 4), "bitwise_equal_to_first": bool(torch.equal(C, ref))}), flush=True)


if __name__ == "__main__":
    main()
