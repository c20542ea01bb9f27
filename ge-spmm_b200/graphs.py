"""Seeded synthetic CSR generators: shape-alikes of the graphs BASELINE.json names.

None of cit-Patents, Reddit, ogbn-products or the SNAP collection can be downloaded here
(reference: data/download_SNAP.sh needs wget), so the benchmarks run on synthetic CSR of
the named (N, nnz) with a comparable degree skew.  All generators are torch programs that
run on CPU (tests, small) or on the GPU (bench, full size), take a seed, and return
``(rowptr int32 [M+1], colind int32 [nnz])`` with columns sorted inside each row -- the
order the reference reader produces (util/util.hpp:75-102).  Duplicate (row, col) pairs
are kept ("general" MatrixMarket semantics, util.hpp:327) unless ``dedup=True``.
"""
import math

import torch

# (M == N, nnz) of the BASELINE.json configs
SHAPES = {
    "cit-Patents": (3_774_768, 16_518_948),
    "reddit": (232_965, 114_615_892),
    "ogbn-products": (2_449_029, 123_718_280),
    "rmat-10m": (10_000_000, 200_000_000),
}


def _gen(device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def coo_to_csr(rows, cols, M, N, dedup=False):
    """int64 COO (any order) -> (rowptr, colind) int32, sorted by (row, col)."""
    key = rows * N + cols
    del rows, cols
    key = torch.unique(key) if dedup else torch.sort(key).values
    rows = torch.div(key, N, rounding_mode="floor")
    colind = (key - rows * N).to(torch.int32)
    del key
    counts = torch.bincount(rows, minlength=M)
    del rows
    rowptr = torch.zeros(M + 1, dtype=torch.int64, device=colind.device)
    torch.cumsum(counts, 0, out=rowptr[1:])
    return rowptr.to(torch.int32), colind


def _sample_by_weight(w, n, g):
    """n ids drawn with probability proportional to w (inverse-CDF sampling)."""
    cdf = torch.cumsum(w.double(), 0)
    u = torch.rand(n, generator=g, device=w.device, dtype=torch.float64) * cdf[-1]
    return torch.searchsorted(cdf, u).clamp_(max=w.numel() - 1)


def uniform_csr(M, N, nnz, seed=0, device="cpu", dedup=False):
    """nnz entries uniformly at random over the M x N grid."""
    g = _gen(device, seed)
    rows = torch.randint(0, M, (nnz,), generator=g, device=device, dtype=torch.int64)
    cols = torch.randint(0, N, (nnz,), generator=g, device=device, dtype=torch.int64)
    return coo_to_csr(rows, cols, M, N, dedup)


def citation_like(N=SHAPES["cit-Patents"][0], nnz=SHAPES["cit-Patents"][1], seed=1, device="cpu",
                  local_fraction=0.7, zero_fraction=0.45):
    """Directed citation graph: power-law out-degree with many empty rows (mean nnz/N), and most
    targets are *older, nearby* ids (col < row, exponential lag) -- the locality that real
    cit-Patents has because patent ids are issued in time order.  local_fraction=0 gives the
    uniformly-random variant."""
    g = _gen(device, seed)
    w = torch.rand(N, generator=g, device=device).clamp_(min=1e-9).pow_(-1.0 / 2.2)  # Pareto(alpha=2.2)
    w.clamp_(max=200.0)
    w[torch.rand(N, generator=g, device=device) < zero_fraction] = 0
    rows = _sample_by_weight(w, nnz, g)
    del w
    lag = (-torch.log(torch.rand(nnz, generator=g, device=device, dtype=torch.float64).clamp_(min=1e-12))
           * (0.01 * N)).long() + 1
    local = torch.remainder(rows - lag, N)
    unif = torch.randint(0, N, (nnz,), generator=g, device=device, dtype=torch.int64)
    cols = torch.where(torch.rand(nnz, generator=g, device=device) < local_fraction, local, unif)
    del lag, local, unif
    return coo_to_csr(rows, cols, N, N)


def social_like(N, nnz, seed=2, device="cpu", sigma=1.0, locality=0.0, window=0.001):
    """Undirected heavy-tailed graph (Chung-Lu with log-normal weights), mirrored so that the CSR
    is symmetric (CSC == CSR).  ``locality`` is the fraction of edges whose second endpoint is
    drawn from a +-window*N neighbourhood of the first (community structure)."""
    g = _gen(device, seed)
    half = nnz // 2
    w = torch.exp(torch.randn(N, generator=g, device=device) * sigma)
    u = _sample_by_weight(w, half, g)
    v = _sample_by_weight(w, half, g)
    if locality > 0:
        span = max(1, int(window * N))
        near = torch.remainder(u + torch.randint(-span, span + 1, (half,), generator=g, device=device), N)
        v = torch.where(torch.rand(half, generator=g, device=device) < locality, near, v)
    del w
    rows = torch.cat([u, v])
    cols = torch.cat([v, u])
    del u, v
    return coo_to_csr(rows, cols, N, N)


def reddit_like(seed=2, device="cpu", scale=1.0):
    N, nnz = SHAPES["reddit"]
    return social_like(max(2, int(N * scale)), max(2, int(nnz * scale)), seed, device, sigma=1.0, locality=0.3, window=0.01)


def products_like(seed=3, device="cpu", scale=1.0):
    N, nnz = SHAPES["ogbn-products"]
    return social_like(max(2, int(N * scale)), max(2, int(nnz * scale)), seed, device, sigma=1.2, locality=0.5, window=0.001)


def rmat(N=SHAPES["rmat-10m"][0], nnz=SHAPES["rmat-10m"][1], seed=4, device="cpu", a=0.57, b=0.19, c=0.19, chunk=1 << 26):
    """R-MAT (a, b, c, d) over 2^ceil(log2 N) ids folded into [0, N) by modulo; duplicates kept."""
    g = _gen(device, seed)
    bits = max(1, math.ceil(math.log2(N)))
    keys = []
    for s in range(0, nnz, chunk):
        n = min(chunk, nnz - s)
        r = torch.zeros(n, dtype=torch.int64, device=device)
        cidx = torch.zeros(n, dtype=torch.int64, device=device)
        for _ in range(bits):
            u = torch.rand(n, generator=g, device=device)
            rbit = (u >= a + b)
            cbit = ((u >= a) & (u < a + b)) | (u >= a + b + c)
            r = (r << 1) | rbit.long()
            cidx = (cidx << 1) | cbit.long()
        keys.append(torch.remainder(r, N) * N + torch.remainder(cidx, N))
        del r, cidx
    key = torch.cat(keys) if len(keys) > 1 else keys[0]
    del keys
    rows = torch.div(key, N, rounding_mode="floor")
    cols = key - rows * N
    del key
    return coo_to_csr(rows, cols, N, N)


def degree_stats(rowptr):
    deg = (rowptr[1:] - rowptr[:-1]).long()
    return {
        "rows": int(deg.numel()), "nnz": int(deg.sum()), "mean": float(deg.float().mean()),
        "max": int(deg.max()), "empty_rows": int((deg == 0).sum()),
        "p99": int(torch.quantile(deg[:: max(1, deg.numel() // 1_000_000)].float(), 0.99)),
    }


def cli_dense(n_rows, n_cols, seed=1, device="cpu"):
    """Dense operand with the value set of the CLI's B = (rand()%100-50)/100 (spmm_test.cu:592-594):
    multiples of 0.01 in [-0.5, 0.49].  (Same distribution, torch's generator instead of glibc rand.)"""
    g = _gen(device, seed)
    return (torch.randint(0, 100, (n_rows, n_cols), generator=g, device=device) - 50).float() / 100


def write_mtx(path, rowptr, colind, N=None, field="pattern", symmetry="general", values=None):
    """Write a CSR as a MatrixMarket coordinate file (1-based), for driving the CLI on synthetic graphs."""
    import numpy as np
    rowptr = rowptr.cpu().numpy() if hasattr(rowptr, "cpu") else np.asarray(rowptr)
    colind = colind.cpu().numpy() if hasattr(colind, "cpu") else np.asarray(colind)
    M = rowptr.shape[0] - 1
    N = M if N is None else N
    if symmetry == "general" and field in ("pattern", "real"):
        # the library's parallel writer (gespmm_write_mtx): seconds instead of a minute at 10^7 entries
        from . import capi
        v = None if field == "pattern" else (np.ones(colind.shape[0], np.float32) if values is None else np.asarray(values, np.float32))
        capi.write_mtx(path, M, N, rowptr, colind, v)
        return
    rows = np.repeat(np.arange(M, dtype=np.int64), np.diff(rowptr)) + 1
    cols = colind.astype(np.int64) + 1
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate %s %s\n%%\n%d %d %d\n" % (field, symmetry, M, N, cols.shape[0]))
        if field == "pattern":
            np.savetxt(f, np.stack([rows, cols], 1), fmt="%d %d")
        else:
            v = np.ones(cols.shape[0]) if values is None else np.asarray(values)
            fmt = "%d %d %d" if field == "integer" else "%d %d %.9g"
            np.savetxt(f, np.stack([rows, cols, v], 1), fmt=fmt)
