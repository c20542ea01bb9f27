#!/usr/bin/env python
"""Randomised differential test of the C ABI against the oracle (GPU box): random small graphs (empty rows, rows ending on
every chunk boundary, an occasional long / huge row), random K, random gespmm_opts (walker, sequential flag, fused vectors,
padding workspace, longest-row hint, L2 policy), sum and max.  Every case: bit for bit where
gespmm_row_sum_is_sequential_ex says so, 1e-4 of max(|G|, sum|a||b|) elsewhere; strided operands; the untouched padding
columns of C stay untouched.
    python scripts/fuzz_gpu.py [--cases 600] [--seed 0]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

KS = [1, 2, 3, 4, 5, 7, 8, 12, 13, 16, 17, 20, 24, 31, 32, 33, 36, 41, 47, 48, 63, 64, 65, 68, 100, 127, 128, 129, 132, 200, 256, 260, 384, 513]


def run(cases, seed, log=print):
    """Returns the number of failing cases."""
    class args:  # noqa: N801
        pass
    args.cases, args.seed = cases, seed
    print = log  # noqa: A001
    entry.load_package()
    from gespmm_b200 import capi
    oracle = entry.load_oracle()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(args.seed)
    st = torch.cuda.current_stream().cuda_stream
    walkers = [capi.WALKER_AUTO, capi.WALKER_AUTO, capi.WALKER_RING, capi.WALKER_REGISTER, capi.WALKER_SUBWARP, capi.WALKER_ROWS, capi.WALKER_BULK]
    bad = 0
    for case in range(args.cases):
        M = int(rng.integers(1, 400)); N = int(rng.integers(1, 400))
        shape = rng.integers(0, 4)
        deg = rng.integers(0, [3, 12, 70, 40][shape], M)
        deg[rng.random(M) < rng.random() * 0.6] = 0
        if rng.random() < 0.5:   # rows that end exactly on chunk boundaries
            deg[rng.integers(0, M, 3)] = rng.choice([31, 32, 33, 63, 64, 65, 96, 128])
        if rng.random() < 0.12:
            deg[rng.integers(0, M)] = int(rng.choice([4096, 4097, 5000, 9000, 33000]))
        rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
        nnz = int(rowptr[-1])
        colind = rng.integers(0, N, nnz).astype(np.int32)
        K = int(rng.choice(KS))
        valued = rng.random() < 0.5
        val = rng.standard_normal(nnz).astype(np.float32) if valued else None
        B = rng.standard_normal((N, K)).astype(np.float32)
        ldb = K + int(rng.choice([0, 0, 1, 3, 4, 8])); ldc = K + int(rng.choice([0, 0, 2, 4, 5]))
        mx = rng.random() < 0.15
        fuse = (not mx) and rng.random() < 0.35
        rs = (rng.random(M).astype(np.float32) + 0.5) if fuse and rng.random() < 0.7 else None
        cs = (rng.random(N).astype(np.float32) + 0.5) if fuse and rng.random() < 0.7 else None
        bias = rng.standard_normal(K).astype(np.float32) if fuse and rng.random() < 0.7 else None
        walker = int(rng.choice(walkers)); seq = rng.random() < 0.4
        use_ws = rng.random() < 0.6
        hint = int(rng.choice([0, 0, 16, 22, 21])) if not fuse and not mx else 0
        t = lambda a, dt=None: None if a is None else torch.as_tensor(a, device=dev)
        rp, ci, vd, rsd, csd, bd = t(rowptr), t(colind), t(val), t(rs), t(cs), t(bias)
        Bd = torch.zeros(N, ldb, device=dev); Bd[:, :K] = t(B)
        Cd = torch.full((M, ldc), -7.0, device=dev)
        ws_bytes = capi.pad_workspace_bytes(M, N, K, max(nnz, 4 * (M + N))) if use_ws else 0
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        o = capi.opts(sequential=seq, walker=walker, max_row_nnz=int(deg.max()) if rng.random() < 0.5 else -1,
                      row_scale=None if rsd is None else rsd.data_ptr(), col_scale=None if csd is None else csd.data_ptr(),
                      bias=None if bd is None else bd.data_ptr(), l2_policy=hint, l2_window_rows=int(rng.choice([0, 5, -1])),
                      workspace=ws.data_ptr() if ws_bytes else None, workspace_bytes=ws_bytes)
        vp = None if vd is None else vd.data_ptr()
        try:
            if mx:
                capi.csr_spmm_max_f32(M, N, K, nnz, rp.data_ptr(), ci.data_ptr() if nnz else None, vp, Bd.data_ptr(), ldb, Cd.data_ptr(), ldc, -10000.0, st)
            else:
                capi.csr_spmm_f32_ex(M, N, K, nnz, rp.data_ptr(), ci.data_ptr() if nnz else None, vp, Bd.data_ptr(), ldb, Cd.data_ptr(), ldc, o, st)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print("case %d raised %r: M=%d N=%d K=%d nnz=%d walker=%d seq=%s fuse=%s ws=%s max=%s" % (case, e, M, N, K, nnz, walker, seq, fuse, use_ws, mx)); bad += 1
            continue
        got = Cd[:, :K].cpu().numpy()
        ok = bool((Cd[:, K:] == -7.0).all())
        if mx:
            want = oracle.spmm_max(rowptr, colind, val, B, init=-10000.0)
            ok = ok and np.array_equal(got, want)
        else:
            Bs = B if cs is None else (B * cs[:, None]).astype(np.float32)
            want = oracle.spmm(rowptr, colind, val, Bs, fma=True)
            G, mag = oracle.spmm_f64(rowptr, colind, val, Bs)
            if rs is not None:
                want = (want * rs[:, None]).astype(np.float32); G = G * rs[:, None]; mag = mag * rs[:, None]
            if bias is not None:
                want = want + bias; G = G + bias; mag = mag + np.abs(bias)
            seqrows = np.array([capi.row_sum_is_sequential(K, int(d), o) for d in deg], dtype=bool)
            ok = ok and np.array_equal(got[seqrows], want[seqrows])
            ok = ok and bool((np.abs(got.astype(np.float64) - G) <= 1e-4 * np.maximum(np.abs(G), mag) + 1e-30).all())
        if not ok:
            bad += 1
            print("case %d MISMATCH: M=%d N=%d K=%d nnz=%d maxdeg=%d walker=%d seq=%s valued=%s fuse=(%s,%s,%s) ws=%s hint=%d max=%s ldb=%d ldc=%d" %
                  (case, M, N, K, nnz, int(deg.max()), walker, seq, valued, rs is not None, cs is not None, bias is not None, bool(ws_bytes), hint, mx, ldb, ldc))
    print("fuzz: %d cases, %d bad" % (args.cases, bad))
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=600)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    sys.exit(1 if run(a.cases, a.seed) else 0)


if __name__ == "__main__":
    main()
