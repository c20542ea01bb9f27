#!/usr/bin/env python
"""L2 eviction-priority steering sweep (GPU): gespmm_opts.l2_policy / l2_window_rows / task_keys and the resident-CTA cap
(GESPMM_SMEM_PAD) on the cit-Patents shape (clustered and uniformly random columns), K = 128, valued A == 1.
    python scripts/sweep_l2.py [--workloads citpatents,citpatents_uniform] [--policies 0,20,22,...] [--windows 131072,...]
    python scripts/sweep_l2.py --one 5,22,131072,0,0      # one configuration, 5 launches (for ncu --metrics dram__bytes...)
policy = near + 4 far + 16 store, each 0 normal / 1 evict_first / 2 evict_last / 3 unchanged; 0 = the plain walker.
One JSON line per configuration: median ms of `--batches` batches of `--iters` launches.
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
    ap.add_argument("--workloads", default="citpatents,citpatents_uniform")
    ap.add_argument("--K", type=int, default=128)
    ap.add_argument("--policies", default="0,16,4,20,22,18,54,21")
    ap.add_argument("--windows", default="32768,131072,524288")
    ap.add_argument("--tasks", default="0")
    ap.add_argument("--pads", default="0")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batches", type=int, default=5)
    ap.add_argument("--walkers", default="0,5", help="gespmm_opts.walker: 0 automatic (cp.async ring), 5 TMA bulk copies")
    ap.add_argument("--one", default=None, help="walker,policy,window,task,pad: run just this, 5 launches")
    args = ap.parse_args()
    entry.load_package()
    from gespmm_b200 import capi, graphs
    dev = torch.device("cuda:0")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream().cuda_stream
    for wl in args.workloads.split(","):
        rowptr, colind = bench.make_graph(wl, 1.0, dev)
        M, nnz, K = rowptr.numel() - 1, colind.numel(), args.K
        val = torch.ones(nnz, device=dev)
        B = graphs.cli_dense(M, K, seed=1, device=dev)
        C = torch.empty(M, K, device=dev)
        ref = None

        def run(walker, policy, window, task, pad):
            os.environ["GESPMM_SMEM_PAD"] = str(pad)
            capi.reload_env()
            o = capi.opts(walker=walker, l2_policy=policy, l2_window_rows=window, task_keys=task)
            capi.csr_spmm_f32_ex(M, M, K, nnz, rowptr.data_ptr(), colind.data_ptr(), val.data_ptr(), B.data_ptr(), K,
                                 C.data_ptr(), K, o, st)

        if args.one:
            cfg = [int(x) for x in args.one.split(",")]
            for _ in range(5):
                run(*cfg)
            torch.cuda.synchronize()
            continue
        for walker, pad in ((w, q) for w in (int(x) for x in args.walkers.split(",")) for q in (int(x) for x in args.pads.split(","))):
            for task in (int(x) for x in args.tasks.split(",")):
                for policy in (int(x) for x in args.policies.split(",")):
                    if walker != 5 and (policy & 15):
                        continue  # only the bulk walker's gathers can carry a priority
                    for window in ((int(x) for x in args.windows.split(",")) if (policy & 15) else (0,)):
                        for _ in range(3):
                            run(walker, policy, window, task, pad)
                        torch.cuda.synchronize()
                        times = []
                        for _ in range(args.batches):
                            e0.record()
                            for _ in range(args.iters):
                                run(walker, policy, window, task, pad)
                            e1.record()
                            torch.cuda.synchronize()
                            times.append(e0.elapsed_time(e1) / args.iters)
                        if ref is None:
                            ref = C.clone()
                        same = bool(torch.equal(C, ref))
                        times.sort()
                        print(json.dumps({"workload": wl, "K": K, "walker": walker, "policy": policy, "near": policy & 3, "far": (policy >> 2) & 3,
                                          "store": (policy >> 4) & 3, "window": window, "task": task, "smem_pad": pad,
                                          "ms_median": round(times[len(times) // 2], 4), "ms_min": round(times[0], 4),
                                          "bitwise_equal_to_first": same}), flush=True)
        del rowptr, colind, val, B, C, ref
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
