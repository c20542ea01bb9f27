#!/usr/bin/env python
"""BASELINE.json configs[2]: the Reddit shape (233 k nodes, 114.6 M edges, symmetric), K = 256, through the PyTorch
operator -- SPMMFunction.apply forward and backward (op.py:8-36; the graph is symmetric, so the CSC arrays are the CSR
arrays) -- next to the reference extension's csr_spmm_no_edge_value (oracle/_ref/ref_spmm, when built) on the same GPU.
    python scripts/op_reddit.py [--K 256] [--iters 10] [--scale 1.0]
Prints one JSON line.  GPU box.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402


def timed(fn, iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=256)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    entry.load_package()
    from gespmm_b200 import graphs
    from gespmm_b200.op import SPMMFunction
    oracle = entry.load_oracle()
    dev = torch.device("cuda:0")
    rowptr, colind = graphs.reddit_like(seed=2, device=dev, scale=args.scale)
    N, nnz, K = rowptr.numel() - 1, colind.numel(), args.K
    x = torch.randn(N, K, device=dev, requires_grad=True)
    g = torch.randn(N, K, device=dev)
    flops = 2.0 * nnz * K

    fwd_ms = timed(lambda: SPMMFunction.apply(rowptr, colind, rowptr, colind, x.detach()), args.iters)

    def fwd_bwd():
        x.grad = None
        SPMMFunction.apply(rowptr, colind, rowptr, colind, x).backward(g)
    both_ms = timed(fwd_bwd, args.iters)
    out = {"workload": "Reddit shape-alike N=%d nnz=%d symmetric seed 2" % (N, nnz), "K": K, "unvalued": True,
           "forward_ms": fwd_ms, "forward_gflops": flops / fwd_ms / 1e6, "forward_backward_ms": both_ms,
           "forward_backward_gflops": 2 * flops / both_ms / 1e6}
    y = SPMMFunction.apply(rowptr, colind, rowptr, colind, x.detach())
    if oracle.have_ref(oracle.REF_EXT) and N * K < 2**31:
        ref = oracle.ref_extension()
        ref_ms = timed(lambda: ref.csr_spmm_no_edge_value(rowptr, colind, x.detach()), args.iters)
        r = ref.csr_spmm_no_edge_value(rowptr, colind, x.detach())
        torch.cuda.synchronize()
        short = (rowptr[1:] - rowptr[:-1]) <= 4096
        out["reference_extension"] = {"kernel": "topoCacheCoarsenSPMMKernel via ref_spmm.csr_spmm_no_edge_value (spmm_kernel.cu:31-96, 200-205)",
                                      "forward_ms": ref_ms, "forward_gflops": flops / ref_ms / 1e6,
                                      "bitwise_equal_on_rows_up_to_4096": bool(torch.equal(y[short], r[short])),
                                      "max_abs_diff": float((y - r).abs().max()), "max_abs_ref": float(r.abs().max())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
