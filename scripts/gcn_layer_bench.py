#!/usr/bin/env python
"""What the fused normalisation / bias epilogue (GCNConv(fuse_norm=True), gespmm_opts.row_scale / col_scale / bias) buys:
the aggregation half of a GCN layer -- x * out_deg_norm -> SpMM -> * in_deg_norm -> + bias (pytorch-custom/op.py:142-147) --
forward and forward + backward, unfused (three element-wise kernels around the product) against fused (one kernel), on the
BASELINE shapes; and the PubMed-shaped training loop's time per epoch both ways.
    python scripts/gcn_layer_bench.py [--workloads products,citpatents,reddit] [--Ks 64,128,47] [--iters 10]
One JSON line per (workload, K).  GPU box.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402


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
    ap.add_argument("--workloads", default="products,citpatents,reddit")
    ap.add_argument("--Ks", default="64,128,47")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--epochs", type=int, default=100)
    args = ap.parse_args()
    entry.load_package()
    from gespmm_b200.op import FusedSPMMFunction, SPMMFunction
    dev = torch.device("cuda:0")
    for wl in args.workloads.split(","):
        rowptr, colind = bench.make_graph(wl, 1.0, dev)
        N = rowptr.numel() - 1
        symmetric = wl in ("products", "reddit", "pubmed")
        if symmetric:
            colptr, rowind = rowptr, colind
        else:
            from gespmm_b200.op import spmm
            colptr = torch.empty_like(rowptr); rowind = torch.empty_like(colind)
            spmm.csr2csc(rowptr, colind, colptr, rowind, torch.ones(colind.numel(), device=dev))
        rs = 1.0 / torch.sqrt((rowptr[1:] - rowptr[:-1]).float().clamp_(min=1)).unsqueeze(1)
        cs = 1.0 / torch.sqrt((colptr[1:] - colptr[:-1]).float().clamp_(min=1)).unsqueeze(1)
        for K in (int(k) for k in args.Ks.split(",")):
            x = torch.randn(N, K, device=dev, requires_grad=True)
            bias = torch.randn(K, device=dev, requires_grad=True)
            g = torch.randn(N, K, device=dev)

            def unfused():
                return SPMMFunction.apply(rowptr, colind, colptr, rowind, x * cs) * rs + bias

            def fused():
                return FusedSPMMFunction.apply(rowptr, colind, colptr, rowind, x, rs, cs, bias)

            def fb(f):
                def run():
                    x.grad = None; bias.grad = None
                    f().backward(g)
                return run
            with torch.no_grad():
                same = bool(torch.equal(unfused(), fused()))
                f_un, f_fu = timed(unfused, args.iters), timed(fused, args.iters)
            b_un, b_fu = timed(fb(unfused), args.iters), timed(fb(fused), args.iters)
            print(json.dumps({"workload": wl, "N": N, "nnz": int(colind.numel()), "K": K, "bitwise_equal": same,
                              "forward_ms": {"unfused": round(f_un, 4), "fused": round(f_fu, 4), "speedup": round(f_un / f_fu, 3)},
                              "forward_backward_ms": {"unfused": round(b_un, 4), "fused": round(b_fu, 4), "speedup": round(b_un / b_fu, 3)}}), flush=True)
            del x, bias, g
        del rowptr, colind
        torch.cuda.empty_cache()
    # the training loop of gcn_custom.py on the PubMed-shaped graph, seconds per epoch
    import importlib.util
    spec = importlib.util.spec_from_file_location("gcn_custom", os.path.join(ROOT, "ge-spmm_b200", "gcn_custom.py"))
    gcn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gcn)
    out = {}
    for fuse in (False, True):
        gcn.run(n_hidden=64, layers=2, epochs=5, fuse_norm=fuse, log=lambda *_: None)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = gcn.run(n_hidden=64, layers=2, epochs=args.epochs, fuse_norm=fuse, log=lambda *_: None)
        torch.cuda.synchronize()
        out["fused" if fuse else "unfused"] = {"ms_per_epoch": round((time.perf_counter() - t0) * 1e3 / args.epochs, 3), "last_loss": round(res["last_loss"], 4)}
    print(json.dumps({"workload": "gcn_custom.py, PubMed-shaped synthetic graph, 2 layers, hidden 64, train + eval per epoch", **out}), flush=True)


if __name__ == "__main__":
    main()
