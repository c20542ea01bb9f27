#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the SpMM hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload citpatents|rmat|reddit|products|pubmed] [--K 128] [--scale 1.0]

One "step" = one CSR x dense SpMM over the whole (row-sharded) matrix, C = A @ B, fp32, through
the reference-facing operator (``spmm.csr_spmm`` -> C ABI -> sm_100a kernel), the valued kernel
with A == 1 exactly as the reference CLI times it (spmm_test.cu:573-574, 756).

Default workload (N = 1 and N > 1): BASELINE.json configs[1], the cit-Patents shape-alike
(N = 3,774,768, nnz = 16,518,948, K = 128; synthetic, seeded -- the real file cannot be
downloaded), strong-scaled over N ranks by nnz-balanced row blocks with B replicated on every
rank before the timed region.  At N > 1 the same line carries an ``rmat`` record: configs[4]
(R-MAT 10M x 10M, nnz = 200M, K = 128) on 1 GPU and on the N GPUs, with the replication times.

Prints ONE JSON line (rank 0).  metric = GFLOP/s with flops = 2*nnz*K (spmm_test.cu:728,738).
  value     inputs resident in HBM, CUDA events on the launching stream, max over ranks; the
            median of 5 batches of --steps launches (the reference CLI times ITER back-to-back
            launches, spmm_test.cu:754-760; SURVEY.md 8d asks for the median of >= 5 batches)
  parity    every rank checks a sample of its rows of C against the oracle (bit for bit where the
            library sums in CSR order, 1e-4 of max(|G|, sum|a||b|) elsewhere); a mismatch exits 1
  e2e       same metric with HOST (pinned) buffers: per rank H2D of its CSR block and ITS ROW
            BLOCK of B, all-gather of B over NVLink, the operator, D2H of its block of C, all inside
            the timed region
  roofline  achieved = algorithmic bytes / t.  N = 1: bytes_min = 4(M+1) + 4nnz + 4nnz + 4NK + 4MK
            (SURVEY.md 8d).  N > 1: per rank 4(M_p+1) + 8 nnz_p + 4 K (distinct B rows its block
            references) + 4 M_p K, summed; peak = MEASURED_PEAKS.json hbm_gbs x N.  When the
            128-column slice of B that one pass gathers from fits the L2 the bound is "l2_gather":
            gathered bytes / t against the measured L2-resident 512-byte-row gather ceiling.
  cpu_baseline  the oracle's C restatement of the reference loop (OpenMP over rows, all host
            threads) on the same workload, plus torch.sparse.mm (north_star's named baseline)
--impl reference times that same CPU restatement as the reference arm (the reference has no CPU
implementation of its own beyond the VALIDATE golden loop, spmm_test.cu:595-605, which the
restatement follows; its GPU kernels are timed beside ours as ``reference_kernel_same_gpu``).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

WORKLOADS = {
    # name: (description, generator kwargs)
    "citpatents": "cit-Patents shape-alike (N=3774768, nnz=16518948), synthetic citation graph seed 1",
    "citpatents_uniform": "cit-Patents shape-alike, uniformly random columns (N=3774768, nnz=16518948, local_fraction=0), seed 1",
    "rmat": "R-MAT(0.57,0.19,0.19,0.05) N=10000000 nnz=200000000 seed 4",
    "reddit": "Reddit shape-alike (N=232965, nnz=114615892) symmetric seed 2",
    "products": "ogbn-products shape-alike (N=2449029, nnz=123718280) symmetric seed 3",
    "pubmed": "PubMed shape-alike (N=19717, nnz=88648) symmetric seed 11",
}


def make_graph(name, scale, device):
    from gespmm_b200 import graphs
    if name == "citpatents":
        N, nnz = graphs.SHAPES["cit-Patents"]
        return graphs.citation_like(N=max(2, int(N * scale)), nnz=max(2, int(nnz * scale)), seed=1, device=device)
    if name == "citpatents_uniform":
        N, nnz = graphs.SHAPES["cit-Patents"]
        return graphs.citation_like(N=max(2, int(N * scale)), nnz=max(2, int(nnz * scale)), seed=1, device=device, local_fraction=0.0)
    if name == "rmat":
        N, nnz = graphs.SHAPES["rmat-10m"]
        return graphs.rmat(N=max(2, int(N * scale)), nnz=max(2, int(nnz * scale)), seed=4, device=device)
    if name == "reddit":
        return graphs.reddit_like(seed=2, device=device, scale=scale)
    if name == "products":
        return graphs.products_like(seed=3, device=device, scale=scale)
    if name == "pubmed":
        return graphs.social_like(19717, 88648, seed=11, device=device, sigma=1.0, locality=0.3, window=0.01)
    raise SystemExit("unknown workload %s" % name)


def bytes_min(M, N, K, nnz, valued=True):
    return 4 * (M + 1) + 4 * nnz + (4 * nnz if valued else 0) + 4 * N * K + 4 * M * K


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.active = False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            while not self._stop_evt.is_set():
                if self.active:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for bit, name in names.items():
                            if mask & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                time.sleep(0.01)
        except Exception as e:  # NVML missing: report that, never fake
            self.error = repr(e)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if hasattr(self, "error"):
            out["error"] = self.error
        return out


def bind_near_gpu(index):
    """One process per GPU: run this rank (and first-touch its pinned host buffers) on the CPUs NVML reports as nearest
    to its GPU, so that the end-to-end copies do not cross the socket interconnect.  Returns the number of CPUs, or None."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(index))
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def physical_gpu_index(local_index):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_index])
        except Exception:
            return local_index
    return local_index


LONG_ROW = 4096               # GESPMM_LONG_ROW (include/gespmm.h)
L2_GATHER_PEAK_GBS = 17900.0  # bin/membench on this pool: random 512-byte rows out of an L2-resident table (profiles/r01_membench.txt)
L2_RESIDENT_BYTES = 126 * 2**20  # the L2: membench holds 17.5-17.9 TB/s up to a 100 MB table and 15.1 TB/s at 134 MB
N_BATCHES = 5


def host_threads():
    """Host threads this process may use -- NOT OMP_NUM_THREADS, which torchrun forces to 1 on every rank."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def make_config(args, M, N, nnz, stats):
    """The `config` object; the own arm and the reference arm build it from the same arguments."""
    world = args.gpus
    return {"workload": WORKLOADS[args.workload], "K": args.K, "M": M, "N": N, "nnz": nnz, "valued": not args.unvalued,
            "scale": args.scale,
            "sharding": "nnz-balanced contiguous row blocks over %d ranks, B replicated on every rank before the timed region" % world
                        if world > 1 else "none",
            "l2": "inputs larger than L2 (B + C >= %.2f GB per rank vs 126 MB)" % ((N * args.K + (M // world) * args.K) * 4 / 1e9),
            "degree_stats": stats}


def cpu_legs(rowptr, colind, B, K, want_torch=True):
    """The oracle restatement (all host threads) and torch.sparse.mm on the host cores, whole workload."""
    oracle = entry.load_oracle()
    rp, ci, Bh = rowptr.cpu().numpy(), colind.cpu().numpy(), B.cpu().numpy()
    ones = np.ones(ci.shape[0], np.float32)
    nnz = ci.shape[0]
    flops = 2.0 * nnz * K
    threads = host_threads()
    w = min(rp.shape[0] - 1, 1000)  # warm the thread pool up on a few rows
    oracle.spmm(rp[: w + 1], ci[: rp[w]], ones[: rp[w]], Bh, nthreads=threads)
    passes, t0 = 0, time.perf_counter()
    while passes < 3 or (time.perf_counter() - t0 < 10.0 and passes < 50):  # about 10 s of CPU work
        oracle.spmm(rp, ci, ones, Bh, fma=True, nthreads=threads)
        passes += 1
    dt = (time.perf_counter() - t0) / passes
    out = {"value": flops / dt / 1e9, "unit": "GFLOP/s", "cores": threads, "kind": "port",
           "sample": "whole workload, mean of %d passes (%.2f s each): oracle/spmm_oracle.c OpenMP over rows" % (passes, dt), "seconds": dt}
    if want_torch:
        try:
            torch.set_num_threads(threads)
            A = torch.sparse_csr_tensor(rowptr.cpu().long(), colind.cpu().long(), torch.ones(nnz), size=(rp.shape[0] - 1, Bh.shape[0]))
            Bt = torch.from_numpy(Bh)
            torch.sparse.mm(A, Bt)
            t0 = time.perf_counter()
            for _ in range(3):
                torch.sparse.mm(A, Bt)
            dtt = (time.perf_counter() - t0) / 3
            out["torch_sparse_mm"] = {"value": flops / dtt / 1e9, "unit": "GFLOP/s", "threads": torch.get_num_threads(),
                                      "cpu_count": os.cpu_count(), "seconds": dtt}
        except Exception as e:
            out["torch_sparse_mm"] = {"error": repr(e)}
    return out


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm (oracle port), ALL host threads (set explicitly: torchrun forces
    OMP_NUM_THREADS=1), same workload and the same `config` object as the own arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    entry.load_package()
    from gespmm_b200 import graphs
    oracle = entry.load_oracle()
    threads = host_threads()
    dev = "cuda" if torch.cuda.is_available() else "cpu"  # the GPU is only used to generate the synthetic graph faster
    rowptr, colind = make_graph(args.workload, args.scale, dev)
    stats = graphs.degree_stats(rowptr)
    rowptr, colind = rowptr.cpu(), colind.cpu()
    M = N = rowptr.numel() - 1
    nnz, K = colind.numel(), args.K
    B = graphs.cli_dense(N, K, seed=1, device="cpu")
    rp, ci, Bh = rowptr.numpy(), colind.numpy(), B.numpy()
    ones = None if args.unvalued else np.ones(nnz, np.float32)
    # bound each step to a prefix of the rows worth <= ~2 s of CPU work (whole matrix when it fits)
    r0 = min(M, 20000)
    t0 = time.perf_counter()
    oracle.spmm(rp[: r0 + 1], ci[: rp[r0]], None if ones is None else ones[: rp[r0]], Bh, nthreads=threads)
    probe = max(time.perf_counter() - t0, 1e-6)
    rate = max(rp[r0], 1) / probe  # nnz per second
    per_step_s = min(2.0, 120.0 / max(1, args.steps + args.warmup))  # the whole run stays within a few minutes
    budget_nnz = rate * per_step_s
    rows = M if nnz <= budget_nnz else int(np.searchsorted(rp, budget_nnz))
    rows = max(rows, 1)
    s_nnz = int(rp[rows])
    rp_s, ci_s, on_s = rp[: rows + 1], ci[:s_nnz], (None if ones is None else ones[:s_nnz])
    for _ in range(max(1, min(args.warmup, 3))):
        oracle.spmm(rp_s, ci_s, on_s, Bh, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.spmm(rp_s, ci_s, on_s, Bh, nthreads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    val = 2.0 * s_nnz * K / dt / 1e9
    sample = "rows [0,%d) of %d, nnz %d of %d per step; %d OpenMP threads" % (rows, M, s_nnz, nnz, threads)
    emit({
        "impl": "reference", "metric": "spmm_gflops", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args, M, N, nnz, stats),
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_REAL_STDOUT = None


def claim_stdout():
    """Route everything written to fd 1 (e.g. NCCL's version banner, library chatter) to stderr and keep
    the real stdout for the one JSON line the driver parses."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj):
    (_REAL_STDOUT or sys.stdout).write(json.dumps(obj) + "\n")
    (_REAL_STDOUT or sys.stdout).flush()


# ---- pieces of the own arm ------------------------------------------------------------------------------------------

class Ranks:
    """rank / world and the few collectives the bench needs (no-ops at world == 1)."""

    def __init__(self, dev):
        import torch.distributed as dist
        self.dist, self.dev = dist, dev
        self.rank, self.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def reduce(self, values, op="max", dtype=torch.float64):
        t = torch.tensor(values, device=self.dev, dtype=dtype)
        if self.world > 1:
            self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN, "sum": self.dist.ReduceOp.SUM}[op])
        return t.tolist()

    def gather(self, value, dtype=torch.float64):
        t = torch.zeros(self.world, device=self.dev, dtype=dtype)
        t[self.rank] = value
        if self.world > 1:
            self.dist.all_reduce(t)
        return t.tolist()


def timed_batches(step, steps, batches, R):
    """`batches` timed regions of exactly `steps` steps, each bracketed by a barrier + synchronize on both sides and
    timed with CUDA events on the launching stream.  Returns (per-batch ms per step as the MAX over ranks,
    per-rank ms per step of the median batch)."""
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mine = []
    for _ in range(batches):
        R.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        R.barrier()
        mine.append(e0.elapsed_time(e1) / steps)
    worst = R.reduce(mine, "max")
    order = sorted(range(batches), key=lambda i: worst[i])
    mid = order[batches // 2]
    return worst, mid, R.gather(mine[mid])


def check_parity(oracle, capi, rowptr, colind, val, B, C, K, sample=1536, seed=0):
    """A sample of the rows of C (the longest few + random ones) against the oracle on the same inputs.
    Bit for bit on the rows summed in CSR order (``capi.row_sum_is_sequential(K, row_nnz)``: the ctypes binding for bare
    C ABI calls, the ``spmm`` extension module for products made through the operator), within
    1e-4 * max(|G|, sum |a||b|) of the fp64 golden on the others.  Returns a dict of counts (this rank)."""
    M = rowptr.numel() - 1
    if M == 0:
        return {"rows_checked": 0, "rows_bitwise": 0, "rows_bad": 0, "max_rel": 0.0}
    deg = (rowptr[1:] - rowptr[:-1]).long()
    g = torch.Generator(device="cpu").manual_seed(seed)
    pick = torch.cat([torch.topk(deg, min(8, M)).indices.cpu(), torch.randint(0, M, (min(sample, M),), generator=g)])
    idx = torch.unique(pick).to(rowptr.device)
    lens = deg[idx]
    sub_rp = torch.zeros(idx.numel() + 1, dtype=torch.int64, device=rowptr.device)
    torch.cumsum(lens, 0, out=sub_rp[1:])
    total = int(sub_rp[-1])
    pos = torch.repeat_interleave(rowptr[idx].long() - sub_rp[:-1], lens) + torch.arange(total, device=rowptr.device)
    cols = colind[pos].long()
    uniq, inv = torch.unique(cols, return_inverse=True)
    rp_h, ci_h = sub_rp.to(torch.int32).cpu().numpy(), inv.to(torch.int32).cpu().numpy()
    v_h = None if val is None else val[pos].cpu().numpy()
    B_h = B[uniq].cpu().numpy() if total else np.zeros((1, K), np.float32)
    got = C[idx].cpu().numpy()
    want = oracle.spmm(rp_h, ci_h, v_h, B_h, fma=True, nthreads=host_threads())
    lens_h = lens.cpu().numpy()
    seq = (lens_h <= LONG_ROW) if capi.row_sum_is_sequential(K, 2) else (lens_h <= 1)
    bad = int((got[seq] != want[seq]).any(axis=1).sum()) if seq.any() else 0
    max_rel = 0.0
    if not seq.all():
        G, mag = oracle.spmm_f64(rp_h, ci_h, v_h, B_h, nthreads=host_threads())
        err = np.abs(got[~seq].astype(np.float64) - G[~seq]) / np.maximum(np.maximum(np.abs(G[~seq]), mag[~seq]), 1e-30)
        max_rel = float(err.max())
        bad += int((err > 1e-4).any(axis=1).sum())
    return {"rows_checked": int(idx.numel()), "rows_bitwise": int(seq.sum()), "rows_bad": bad, "max_rel": max_rel}


def merged_parity(R, p):
    s = R.reduce([p["rows_checked"], p["rows_bitwise"], p["rows_bad"]], "sum")
    return {"rows_checked": int(s[0]), "rows_bitwise_vs_oracle": int(s[1]), "rows_within_1e-4": int(s[0] - s[1]),
            "rows_bad": int(s[2]), "max_rel": R.reduce([p["max_rel"]], "max")[0], "ok": int(s[2]) == 0,
            "how": "per rank: the 8 longest + up to 1536 random rows of its block of C against oracle/spmm_oracle.c on the same inputs"}


def rank_bytes(sh, K, valued, distinct):
    """Algorithmic bytes of this rank's product: its CSR block, the B rows it references (once each), its block of C."""
    M_loc = sh.row_hi - sh.row_lo
    return 4 * (M_loc + 1) + 4 * sh.nnz_local * (2 if valued else 1) + 4 * K * distinct + 4 * M_loc * K


def measure_e2e(sh, B, K, steps, R, spmm):
    """HOST (pinned) inputs -> device, operator, C back to the host, per step, per rank.  A rank uploads its CSR block
    and ITS ROW BLOCK of B only; the blocks are assembled on every GPU by one all-gather over NVLink inside the timed
    region (at one rank: the whole of B comes over PCIe).  Steps alternate between two streams / two buffer sets so that
    step i's D2H of C overlaps step i+1's H2D of its inputs (PCIe is full duplex)."""
    dev, world, rank = sh.device, sh.world, sh.rank
    M_loc, N = sh.row_hi - sh.row_lo, sh.N
    blk = sh.even_b_block()
    lo, hi = min(N, rank * blk), min(N, (rank + 1) * blk)
    h_rp, h_ci = sh.rowptr.cpu().pin_memory(), sh.colind.cpu().pin_memory()
    h_val = None if sh.val is None else sh.val.cpu().pin_memory()
    h_B = torch.zeros(blk, K, dtype=torch.float32).pin_memory()
    h_B[: hi - lo].copy_(B[lo:hi])
    h2d = h_rp.numel() * 4 + h_ci.numel() * 4 + (0 if h_val is None else h_val.numel() * 4) + (hi - lo) * K * 4
    d2h = M_loc * K * 4
    sets = []
    for _ in range(2):
        sets.append({"stream": torch.cuda.Stream(), "rp": torch.empty_like(sh.rowptr), "ci": torch.empty_like(sh.colind),
                     "val": None if sh.val is None else torch.empty_like(sh.val),
                     "Bblk": torch.empty(blk, K, dtype=torch.float32, device=dev),
                     "Bfull": torch.empty(blk * world, K, dtype=torch.float32, device=dev),
                     "hC": torch.empty(M_loc, K, dtype=torch.float32).pin_memory()})

    def e2e_step(i):
        st = sets[i % 2]
        with torch.cuda.stream(st["stream"]):
            st["rp"].copy_(h_rp, non_blocking=True); st["ci"].copy_(h_ci, non_blocking=True)
            if st["val"] is not None:
                st["val"].copy_(h_val, non_blocking=True)
            if world == 1:
                st["Bfull"].copy_(h_B, non_blocking=True)
                Bf = st["Bfull"][:N]
            else:
                st["Bblk"].copy_(h_B, non_blocking=True)
                Bf = sh.all_gather_B_even(st["Bblk"], out=st["Bfull"])
            o = (spmm.csr_spmm_no_edge_value(st["rp"], st["ci"], Bf) if st["val"] is None
                 else spmm.csr_spmm(st["rp"], st["ci"], st["val"], Bf))
            st["hC"].copy_(o, non_blocking=True)

    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_steps = max(4, min(steps, 10))
    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    R.barrier()
    e0.record(stream)
    for st in sets:
        st["stream"].wait_event(e0)
    for i in range(n_steps):
        e2e_step(i)
    for st in sets:
        stream.wait_stream(st["stream"])
    e1.record(stream)
    torch.cuda.synchronize()
    R.barrier()
    ms = R.reduce([e0.elapsed_time(e1) / n_steps], "max")[0]
    assert torch.equal(sets[0]["hC"], sets[1]["hC"])
    tot = R.reduce([float(h2d), float(d2h)], "sum")
    return ms, int(tot[0]), int(tot[1]), n_steps, sets[0]["hC"]


def rmat_record(args, R, dev, oracle, capi, spmm):
    """BASELINE.json configs[4] inside the N > 1 line: R-MAT 10M x 10M, nnz = 200M, K = 128, nnz-balanced row blocks.
    1-GPU time = the whole matrix on rank 0; N-GPU time = max over ranks of the per-rank block product with B resident;
    replication of B measured both ways (NCCL broadcast from rank 0; all-gather of equal row blocks)."""
    from gespmm_b200 import graphs
    from gespmm_b200.sharding import RowShardedSpMM
    dist, rank, world = R.dist, R.rank, R.world
    K = 128
    scale = args.rmat_scale
    Nn, nnz_target = graphs.SHAPES["rmat-10m"]
    rowptr, colind = graphs.rmat(N=max(2, int(Nn * scale)), nnz=max(2, int(nnz_target * scale)), seed=4, device=dev)
    M = N = rowptr.numel() - 1
    nnz = colind.numel()
    chk = R.reduce([float(rowptr.long().sum()), float(colind.long().sum())], "max")
    chk2 = R.reduce([float(rowptr.long().sum()), float(colind.long().sum())], "min")
    assert chk == chk2, "ranks generated different R-MAT graphs"
    val = torch.ones(nnz, dtype=torch.float32, device=dev)
    sh = RowShardedSpMM(rowptr, colind, val, N, rank=rank, world=world, device=dev)
    if rank != 0:
        del rowptr, colind, val
    # B: every rank generates ITS equal row block (seeded by rank), then one all-gather replicates it
    blk = sh.even_b_block()
    lo, hi = min(N, rank * blk), min(N, (rank + 1) * blk)
    Bblk = torch.zeros(blk, K, device=dev)
    Bblk[: hi - lo] = graphs.cli_dense(hi - lo, K, seed=100 + rank, device=dev)
    Bfull = torch.empty(blk * world, K, device=dev)
    ag = []
    for _ in range(3):
        R.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        B = sh.all_gather_B_even(Bblk, out=Bfull)
        torch.cuda.synchronize()
        ag.append((time.perf_counter() - t0) * 1e3)
    bc = []
    Bb = torch.empty(N, K, device=dev)
    if rank == 0:
        Bb.copy_(B)
    for _ in range(3):
        R.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        dist.broadcast(Bb, src=0)
        torch.cuda.synchronize()
        bc.append((time.perf_counter() - t0) * 1e3)
    same_b = bool(torch.equal(Bb, B))
    del Bb
    ag_ms, bc_ms = R.reduce([min(ag[1:])], "max")[0], R.reduce([min(bc[1:])], "max")[0]

    def step():
        return sh.forward(B)
    for _ in range(3):
        C = step()
    steps = max(5, min(args.steps, 20))
    worst, mid, per_rank = timed_batches(step, steps, N_BATCHES, R)
    ms_n = worst[mid]
    C = step()
    torch.cuda.synchronize()
    par = merged_parity(R, check_parity(oracle, spmm, sh.rowptr, sh.colind, sh.val, B, C, K, seed=17 + rank))
    del C
    # what one step costs when B changes every step: all-gather then product, vs the replication pipelined with the
    # product panel by panel (RowShardedSpMM.forward_replicating: 4 column panels on a communication stream)
    def step_seq():
        return sh.forward(sh.all_gather_B_even(Bblk, out=Bfull))

    def step_pipe():
        return sh.forward_replicating(Bblk, chunks=2)

    def step_pipe4():
        return sh.forward_replicating(Bblk, chunks=4)
    repl = {}
    for name, fn in (("allgather_then_product", step_seq), ("pipelined_2_panels", step_pipe), ("pipelined_4_panels", step_pipe4)):
        for _ in range(2):
            Cp = fn()
        w2, m2, _ = timed_batches(fn, 5, 3, R)
        repl[name] = w2[m2]
    Cp = step_pipe()
    torch.cuda.synchronize()
    par_pipe = merged_parity(R, check_parity(oracle, spmm, sh.rowptr, sh.colind, sh.val, B, Cp, 64, seed=91 + rank))
    del Cp
    ms_1 = None
    if rank == 0:
        mx = int(spmm.max_row_nnz(rowptr))
        for _ in range(2):
            C1 = spmm.csr_spmm_ex(rowptr, colind, val, B, max_row_nnz=mx)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        for _ in range(3):
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                C1 = spmm.csr_spmm_ex(rowptr, colind, val, B, max_row_nnz=mx)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 5)
        ms_1 = sorted(times)[1]
        del C1
    R.barrier()
    ms_1 = R.reduce([ms_1 if ms_1 is not None else 0.0], "max")[0]
    distinct = sh.distinct_b_rows()
    bm = R.reduce([float(rank_bytes(sh, K, True, distinct))], "sum")[0]
    peak, _ = measured_peak_gbs()
    flops = 2.0 * nnz * K
    return {
        "workload": "R-MAT(0.57,0.19,0.19,0.05) N=%d nnz=%d seed 4, K=128, valued A == 1 (BASELINE.json configs[4])" % (N, nnz),
        "ms_1gpu": ms_1, "gflops_1gpu": flops / ms_1 / 1e6, "ms_per_step": ms_n, "value": flops / ms_n / 1e6, "unit": "GFLOP/s",
        "speedup": ms_1 / ms_n, "n_gpus": world, "per_rank_ms": [round(x, 4) for x in per_rank],
        "batches_ms_per_step": [round(x, 4) for x in worst], "steps": steps,
        "b_allgather_ms": ag_ms, "b_broadcast_ms": bc_ms, "b_replicas_identical": same_b,
        "ms_per_step_including_allgather": ms_n + ag_ms, "ms_per_step_including_broadcast": ms_n + bc_ms,
        "ms_per_step_replicating": repl, "parity_pipelined": par_pipe,
        "roofline": {"bound": "hbm", "bytes_min": bm, "achieved": bm / ms_n / 1e6, "peak": peak * world, "unit": "GB/s",
                     "frac": bm / ms_n / 1e6 / (peak * world),
                     "bytes_model": "per rank: CSR block + distinct B rows referenced + C block, summed over ranks"},
        "parity": par,
    }


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="citpatents", choices=sorted(WORKLOADS))
    ap.add_argument("--K", type=int, default=128)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--unvalued", action="store_true", help="time csr_spmm_no_edge_value instead of csr_spmm with A == 1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ref-kernel", action="store_true")
    ap.add_argument("--no-rmat", action="store_true", help="N > 1: skip the R-MAT 10M/200M record")
    ap.add_argument("--no-uniform", action="store_true", help="N = 1: skip the uniformly-random variant of the cit-Patents shape")
    ap.add_argument("--rmat-scale", type=float, default=1.0)
    ap.add_argument("--b-sharded", action="store_true",
                    help="N > 1: leave B row-sharded (no replication) and let the kernel gather remote rows over NVLink")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 3
        args.warmup = args.warmup if args.warmup is not None else 1
        return run_reference_arm(args)
    args.steps = args.steps if args.steps is not None else 200
    args.warmup = max(3, args.warmup if args.warmup is not None else 10)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (own arm) needs a CUDA device: this framework has no CPU path")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_near_gpu(physical_gpu_index(local)) if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None
    R = Ranks(dev)
    rank, world, dist = R.rank, R.world, R.dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.gpus = world

    entry.load_package()
    from gespmm_b200 import capi, graphs
    from gespmm_b200.op import spmm
    from gespmm_b200.sharding import RowShardedSpMM
    oracle = entry.load_oracle()

    K = args.K
    valued = not args.unvalued
    rowptr, colind = make_graph(args.workload, args.scale, dev)   # same seed on every rank -> same graph
    M = N = rowptr.numel() - 1
    nnz = colind.numel()
    sums = [float(rowptr.long().sum()), float(colind.long().sum())]
    assert R.reduce(sums, "max") == R.reduce(sums, "min"), "ranks generated different graphs"
    val = torch.ones(nnz, dtype=torch.float32, device=dev) if valued else None
    sh = RowShardedSpMM(rowptr, colind, val, N, rank=rank, world=world, device=dev)
    stats = graphs.degree_stats(rowptr) if rank == 0 else None
    B0 = graphs.cli_dense(N, K, seed=1, device=dev) if rank == 0 else None
    # replication of B before the timed region (`value` = inputs resident): one NCCL broadcast, timed on its own
    torch.cuda.synchronize()
    R.barrier()
    t0 = time.perf_counter()
    B = sh.broadcast_B(B0, K, root=0)
    torch.cuda.synchronize()
    bcast_ms = R.reduce([(time.perf_counter() - t0) * 1e3], "max")[0]
    rowptr_full, colind_full = (rowptr, colind) if rank == 0 else (None, None)
    del rowptr, colind, val
    M_loc, nnz_loc = sh.row_hi - sh.row_lo, sh.nnz_local
    per_rank_shape = [[int(a), int(b)] for a, b in zip(R.gather(M_loc), R.gather(nnz_loc))]

    remote_frac = None
    if args.b_sharded and world > 1:
        bb = sh.b_row_bounds()
        parts = sh.share_B_parts(B[bb[rank]:bb[rank + 1]].clone())
        remote = ((sh.colind < bb[rank]) | (sh.colind >= bb[rank + 1])).sum()
        t = R.reduce([float(remote), float(nnz_loc)], "sum")
        remote_frac = t[0] / t[1]

        def step():
            return sh.forward_sharded_B(parts, K, force=True)
    else:
        def step():
            return sh.forward(B)

    for _ in range(args.warmup):
        C = step()
    torch.cuda.synchronize()

    sampler = ClockSampler(physical_gpu_index(local)) if rank == 0 else None
    if sampler:
        sampler.start()
        sampler.active = True
    worst, mid, per_rank_ms = timed_batches(step, args.steps, N_BATCHES, R)
    if sampler:
        sampler.active = False
    ms = worst[mid]
    clocks = sampler.stop() if sampler else None
    flops = 2.0 * nnz * K
    value = flops / (ms * 1e-3) / 1e9
    long_rows_local = sh.max_row_nnz > LONG_ROW or sh.max_row_nnz < 0
    launches_per_step = 2 if (long_rows_local and nnz_loc > LONG_ROW) else 1

    # parity: a sample of every rank's rows against the oracle; a mismatch fails the run
    C = step()
    torch.cuda.synchronize()
    parity = merged_parity(R, check_parity(oracle, spmm, sh.rowptr, sh.colind, sh.val, B, C, K, seed=rank))

    # roofline
    peak, peak_src = measured_peak_gbs()
    distinct = sh.distinct_b_rows()
    distinct_all = R.gather(distinct)
    if world == 1:
        bm = float(bytes_min(M, N, K, nnz, valued))
        bytes_model = "SURVEY.md 8d: 4(M+1) + 4nnz (+4nnz valued) + 4NK + 4MK, every operand once"
    else:
        bm = R.reduce([float(rank_bytes(sh, K, valued, distinct))], "sum")[0]
        bytes_model = "per rank: CSR block + distinct B rows its block references + C block, summed over ranks"
    panel_bytes = N * min(K, 128) * 4
    gather_bytes = 4.0 * (M + world) + 4.0 * nnz * (2 if valued else 1) + 4.0 * nnz * K + 4.0 * M * K
    traffic, traffic_src = None, None
    if world == 1 and args.scale == 1.0:
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            traffic = tj.get("%s_K%d_%s" % (args.workload, K, "valued" if valued else "unvalued"))
            traffic_src = tj.get("_source")
        except Exception:
            pass
    if panel_bytes <= L2_RESIDENT_BYTES:
        achieved = gather_bytes / (ms * 1e-3) / 1e9
        roofline = {"bound": "l2_gather", "achieved": achieved, "peak": L2_GATHER_PEAK_GBS * world, "unit": "GB/s",
                    "frac": achieved / (L2_GATHER_PEAK_GBS * world), "traffic": traffic,
                    "peak_source": "bin/membench: random 512-byte rows out of an L2-resident table (profiles/r01_membench.txt); "
                                   "the %d-byte slice of B one 128-column pass gathers from fits the L2" % panel_bytes,
                    "bytes_model": "gathered bytes: CSR + nnz*K*4 + C (every nonzero fetches its B row from the L2)",
                    "bytes": gather_bytes, "hbm_bytes_min": bm, "hbm_frac": bm / (ms * 1e-3) / 1e9 / (peak * world)}
    else:
        achieved = bm / (ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "bytes_min": bm,
                    "bytes_model": bytes_model, "frac_of_8TBs_nominal": achieved / (8000.0 * world),
                    "distinct_b_rows_per_rank": [int(x) for x in distinct_all],
                    # secondary, labelled: no-B-reuse model, every nonzero gathers its own B row from DRAM
                    "bytes_gather_model": gather_bytes, "gather_model_gbs": gather_bytes / (ms * 1e-3) / 1e9}

    e2e = None
    if not args.no_e2e:
        del C
        e2e_ms, h2d, d2h, e2e_steps, hC = measure_e2e(sh, B, K, args.steps, R, spmm)
        e2e = {"value": flops / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
               "pcie_gbs_each_way_per_rank": [h2d / world / (e2e_ms * 1e-3) / 1e9, d2h / world / (e2e_ms * 1e-3) / 1e9],
               "api": "per rank: pinned host CSR block + its row block of B copied in, " +
                      ("one all-gather of B over NVLink, " if world > 1 else "") +
                      "spmm.csr_spmm, its block of C copied out to pinned host memory; steps alternate over two streams so "
                      "D2H of one overlaps H2D of the next"}
        # the host copy of C that came back is checked too
        Cd = step()
        assert torch.equal(hC, Cd.cpu()), "e2e result differs from the device-resident result"
        del Cd

    out = {
        "metric": "spmm_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        # cit-Patents K=128, GE-SpMM cachec2_2 = 140.4 GFLOP/s (matrix_id_info.xlsx DL19; the real
        # graph, GPU unstated, single GPU).  Only the default workload at K=128 has a published counterpart.
        "vs_baseline": (value / 140.4) if (args.workload == "citpatents" and K == 128 and args.scale == 1.0) else None,
        "vs_baseline_note": "BASELINE.md: 140.4 GFLOP/s = GE-SpMM cachec2_2 on the real cit-Patents, K=128, unstated 2019-era GPU; "
                            "this run is the seeded synthetic shape-alike of the same N and nnz",
        "dtype": "f32", "data": "synthetic",
        "config": make_config(args, M, N, nnz, stats),
        "timing": "median of %d batches of %d steps; each batch bracketed by barrier + synchronize, CUDA events, max over ranks" % (N_BATCHES, args.steps),
        "batches_ms_per_step": [round(x, 5) for x in worst],
        "roofline": roofline, "e2e": e2e, "parity": parity,
        "gpu_launches": args.steps * N_BATCHES * launches_per_step, "gpu_launches_per_step": launches_per_step,
        "clocks": clocks, "b_broadcast_ms": bcast_ms if world > 1 else None, "per_rank_ms": [round(x, 5) for x in per_rank_ms],
        "b_layout": ("row-sharded, remote rows gathered over NVLink inside the kernel (no replication); fraction of gathers "
                     "that are remote: %.3f" % remote_frac) if remote_frac is not None else "replicated on every rank",
        "per_rank_rows_nnz": per_rank_shape, "cpus_bound_near_gpu": numa, "per_rank_max_row_nnz": [int(x) for x in R.gather(sh.max_row_nnz)],
    }

    # reference kernel on the same GPU ("the bar on the B200 box"), whole matrix, rank 0 only
    if rank == 0 and not args.no_ref_kernel:
        if oracle.have_ref(oracle.REF_CLI_KERNELS) and (N * K < 2**31 and M * K < 2**31):
            try:
                L = oracle.ref_cli_kernels()
                ones = torch.ones(nnz, dtype=torch.float32, device=dev)
                Cr = torch.empty(M, K, device=dev)
                torch.cuda.synchronize()
                rms = L.ref_spmm_time_ms(2, 8, M, K, rowptr_full.data_ptr(), colind_full.data_ptr(), ones.data_ptr(),
                                         B.data_ptr(), Cr.data_ptr(), 3, max(5, min(args.steps, 50)))
                torch.cuda.synchronize()
                Cl = step()
                Crl = Cr[sh.row_lo:sh.row_hi]
                short = ((rowptr_full[1:] - rowptr_full[:-1]) <= LONG_ROW)[sh.row_lo:sh.row_hi]
                out["reference_kernel_same_gpu"] = {
                    "kernel": "spmm_test2<float>, tile_row 8 (spmm_test.cu:161-236, 756), 1 GPU, whole matrix",
                    "ms": rms, "value": flops / (rms * 1e-3) / 1e9, "unit": "GFLOP/s",
                    "bitwise_equal_to_ours": bool(torch.equal(Crl, Cl)),
                    "bitwise_equal_on_rows_up_to_GESPMM_LONG_ROW": bool(torch.equal(Crl[short], Cl[short])),
                    "max_abs_diff": float((Crl - Cl).abs().max())}
                # the reference's own kernel on the same inputs is part of the parity verdict of this run
                if spmm.row_sum_is_sequential(K, 2):   # (widths whose default walker re-associates are held to 1e-4, not to bits)
                    parity["reference_kernel_bitwise_on_rows_up_to_GESPMM_LONG_ROW"] = out["reference_kernel_same_gpu"]["bitwise_equal_on_rows_up_to_GESPMM_LONG_ROW"]
                    parity["ok"] = parity["ok"] and parity["reference_kernel_bitwise_on_rows_up_to_GESPMM_LONG_ROW"]
                del Cr, ones, Cl
            except Exception as e:
                out["reference_kernel_same_gpu"] = {"error": repr(e)}
        else:
            out["reference_kernel_same_gpu"] = {"unavailable": "oracle/_ref not built or int32 offsets would overflow"}
    if rank == 0 and not args.no_cpu and world == 1:  # rank 0 at N = 1 only
        out["cpu_baseline"] = cpu_legs(rowptr_full, colind_full, B, K)
    if world == 1 and args.workload == "citpatents" and args.scale == 1.0 and not args.no_uniform:
        # SURVEY.md 8d config 2 asks for the clustered AND the uniformly random graph: the same shape with every column
        # drawn uniformly (no L2 reuse to find) next to the headline -- the floor for a kernel that does not reorder the graph
        rp_u, ci_u = make_graph("citpatents_uniform", 1.0, dev)
        v_u = torch.ones(ci_u.numel(), dtype=torch.float32, device=dev) if valued else None
        mx_u = int(spmm.max_row_nnz(rp_u))

        def step_u():
            return spmm.csr_spmm_ex(rp_u, ci_u, v_u, B, max_row_nnz=mx_u)
        for _ in range(3):
            Cu = step_u()
        w_u, m_u, _ = timed_batches(step_u, min(args.steps, 50), 3, R)
        par_u = check_parity(oracle, spmm, rp_u, ci_u, v_u, B, Cu, K, sample=512, seed=3)
        bm_u = float(bytes_min(M, N, K, ci_u.numel(), valued))
        out["uniform_variant"] = {"workload": WORKLOADS["citpatents_uniform"], "ms_per_step": w_u[m_u],
                                  "value": 2.0 * ci_u.numel() * K / w_u[m_u] / 1e6, "unit": "GFLOP/s",
                                  "roofline_frac": bm_u / w_u[m_u] / 1e6 / peak, "rows_checked": par_u["rows_checked"],
                                  "rows_bad": par_u["rows_bad"]}
        parity["ok"] = parity["ok"] and par_u["rows_bad"] == 0
        del rp_u, ci_u, v_u, Cu

    # BASELINE.json configs[4] (R-MAT 10M/200M, row-sharded) rides on the N > 1 line
    if world > 1 and not args.no_rmat:
        del B, sh, rowptr_full, colind_full
        torch.cuda.empty_cache()
        try:
            out["rmat"] = rmat_record(args, R, dev, oracle, capi, spmm)
        except Exception as e:  # the headline workload's numbers stand; the record says what happened
            out["rmat"] = {"error": repr(e)}
        ok_r = out["rmat"].get("parity", {}).get("ok", False) if "error" not in out["rmat"] else False
    else:
        ok_r = True

    if rank == 0:
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not parity["ok"] or not ok_r:
        sys.stderr.write("bench.py: PARITY FAILURE %s\n" % json.dumps({"main": parity, "rmat": out.get("rmat", {}).get("parity")}))
        sys.exit(1)


if __name__ == "__main__":
    main()
