#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the SpMM hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload citpatents|rmat|reddit|products|pubmed] [--K 128] [--scale 1.0]

One "step" = one CSR x dense SpMM over the whole (row-sharded) matrix, C = A @ B, fp32, through
the reference-facing operator (``spmm.csr_spmm`` -> C ABI -> sm_100a kernel), the valued kernel
with A == 1 exactly as the reference CLI times it (spmm_test.cu:573-574, 756).

Default workload (N = 1 and N > 1): BASELINE.json configs[1], the cit-Patents shape-alike
(N = 3,774,768, nnz = 16,518,948, K = 128; synthetic, seeded -- the real file cannot be
downloaded), strong-scaled over N ranks by nnz-balanced row blocks with B replicated by one
NCCL broadcast before the timed region.  ``--workload rmat`` is configs[4] (10M x 10M,
nnz = 200M, K = 128).

Prints ONE JSON line (rank 0).  metric = GFLOP/s with flops = 2*nnz*K (spmm_test.cu:728,738).
  value     inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e       same metric with HOST (pinned) buffers: H2D of rowptr/colind/val/B, the operator,
            D2H of C, all inside the timed region
  roofline  achieved = bytes_min / t, bytes_min = 4(M+1) + 4nnz + 4nnz + 4NK + 4MK per rank
            (SURVEY.md 8d), peak = MEASURED_PEAKS.json hbm_gbs (x N ranks)
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
    "pubmed": "pubmed.mtx of the reference (tests/golden/pubmed_csr.npz)",
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
        z = np.load(os.path.join(ROOT, "tests", "golden", "pubmed_csr.npz"))
        return torch.from_numpy(z["rowptr"]).to(device), torch.from_numpy(z["colind"]).to(device)
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


def physical_gpu_index(local_index):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_index])
        except Exception:
            return local_index
    return local_index


def cpu_legs(rowptr, colind, B, K, want_torch=True):
    """The oracle restatement (all threads) and torch.sparse.mm on the host cores, whole workload."""
    oracle = entry.load_oracle()
    rp, ci, Bh = rowptr.cpu().numpy(), colind.cpu().numpy(), B.cpu().numpy()
    ones = np.ones(ci.shape[0], np.float32)
    nnz = ci.shape[0]
    flops = 2.0 * nnz * K
    threads = oracle.num_threads()
    w = min(rp.shape[0] - 1, 1000)  # warm the thread pool up on a few rows
    oracle.spmm(rp[: w + 1], ci[: rp[w]], ones[: rp[w]], Bh, nthreads=0)
    passes, t0 = 0, time.perf_counter()
    while passes < 3 or (time.perf_counter() - t0 < 10.0 and passes < 50):  # about 10 s of CPU work
        oracle.spmm(rp, ci, ones, Bh, fma=True, nthreads=0)
        passes += 1
    dt = (time.perf_counter() - t0) / passes
    out = {"value": flops / dt / 1e9, "unit": "GFLOP/s", "cores": threads, "kind": "port",
           "sample": "whole workload, mean of %d passes (%.2f s each): oracle/spmm_oracle.c OpenMP over rows" % (passes, dt), "seconds": dt}
    if want_torch:
        try:
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
    """--impl reference: the reference's CPU algorithm (oracle port), all host threads, same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    entry.load_package()
    from gespmm_b200 import graphs
    oracle = entry.load_oracle()
    dev = "cuda" if torch.cuda.is_available() else "cpu"  # the GPU is only used to generate the synthetic graph faster
    rowptr, colind = make_graph(args.workload, args.scale, dev)
    rowptr, colind = rowptr.cpu(), colind.cpu()
    M = N = rowptr.numel() - 1
    nnz, K = colind.numel(), args.K
    B = graphs.cli_dense(N, K, seed=1, device="cpu")
    rp, ci, Bh = rowptr.numpy(), colind.numpy(), B.numpy()
    ones = np.ones(nnz, np.float32)
    # bound each step to a prefix of the rows worth <= ~2 s of CPU work (whole matrix when it fits)
    t0 = time.perf_counter()
    oracle.spmm(rp[: min(M, 20000) + 1], ci[: rp[min(M, 20000)]], ones[: rp[min(M, 20000)]], Bh)
    probe = max(time.perf_counter() - t0, 1e-6)
    rate = max(rp[min(M, 20000)], 1) / probe  # nnz per second
    per_step_s = min(2.0, 120.0 / max(1, args.steps + args.warmup))  # the whole run stays within a few minutes
    budget_nnz = rate * per_step_s
    rows = M if nnz <= budget_nnz else int(np.searchsorted(rp, budget_nnz))
    rows = max(rows, 1)
    s_nnz = int(rp[rows])
    rp_s, ci_s, on_s = rp[: rows + 1], ci[:s_nnz], ones[:s_nnz]
    for _ in range(max(1, min(args.warmup, 3))):
        oracle.spmm(rp_s, ci_s, on_s, Bh)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.spmm(rp_s, ci_s, on_s, Bh)
    dt = (time.perf_counter() - t0) / args.steps
    val = 2.0 * s_nnz * K / dt / 1e9
    sample = "rows [0,%d) of %d, nnz %d of %d per step" % (rows, M, s_nnz, nnz)
    emit({
        "impl": "reference", "metric": "spmm_gflops", "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "K": K, "M": M, "nnz": nnz, "valued": True, "scale": args.scale},
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": oracle.num_threads(), "kind": "port", "sample": sample},
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
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    entry.load_package()
    from gespmm_b200 import graphs
    from gespmm_b200.op import spmm
    from gespmm_b200.sharding import RowShardedSpMM

    K = args.K
    rowptr, colind = make_graph(args.workload, args.scale, dev)   # same seed on every rank -> same graph
    M = N = rowptr.numel() - 1
    nnz = colind.numel()
    if world > 1:
        chk = torch.stack([rowptr.long().sum(), colind.long().sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks generated different graphs"
    val = None if args.unvalued else torch.ones(nnz, dtype=torch.float32, device=dev)
    sh = RowShardedSpMM(rowptr, colind, val, N, rank=rank, world=world, device=dev)
    stats = graphs.degree_stats(rowptr) if rank == 0 else None
    B0 = graphs.cli_dense(N, K, seed=1, device=dev) if rank == 0 else None
    # replication of B: ONE NCCL broadcast, timed on its own (not part of `value`: inputs are resident)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    B = sh.broadcast_B(B0, K, root=0)
    torch.cuda.synchronize()
    bcast_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([bcast_ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); bcast_ms = float(t)
    rowptr_full, colind_full = (rowptr, colind) if rank == 0 else (None, None)
    if rank != 0:
        del rowptr, colind
    M_loc, nnz_loc = sh.row_hi - sh.row_lo, sh.nnz_local
    per_rank_shape = [[M_loc, nnz_loc]]
    if world > 1:
        t = torch.zeros(world, 2, device=dev, dtype=torch.int64); t[rank, 0] = M_loc; t[rank, 1] = nnz_loc
        dist.all_reduce(t); per_rank_shape = t.tolist()

    remote_frac = None
    if args.b_sharded and world > 1:
        bb = sh.b_row_bounds()
        parts = sh.share_B_parts(B[bb[rank]:bb[rank + 1]].clone())
        remote = ((sh.colind < bb[rank]) | (sh.colind >= bb[rank + 1])).sum()
        t = torch.stack([remote.double(), torch.tensor(float(nnz_loc), device=dev, dtype=torch.float64)])
        dist.all_reduce(t)
        remote_frac = float(t[0] / t[1])

        def step():
            return sh.forward_sharded_B(parts, K)
    else:
        def step():
            return sh.forward(B)

    for _ in range(args.warmup):
        C = step()
    torch.cuda.synchronize()

    sampler = ClockSampler(physical_gpu_index(local)) if rank == 0 else None
    if sampler:
        sampler.start()
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.active = True
    e0.record(stream)
    for _ in range(args.steps):
        C = step()
    e1.record(stream)
    torch.cuda.synchronize()
    if sampler:
        sampler.active = False
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_total = float(t)
    per_rank_ms = [e0.elapsed_time(e1) / args.steps]
    if world > 1:
        t = torch.zeros(world, device=dev); t[rank] = per_rank_ms[0]; dist.all_reduce(t); per_rank_ms = [round(float(x), 4) for x in t]
    ms = ms_total / args.steps
    clocks = sampler.stop() if sampler else None
    flops = 2.0 * nnz * K
    value = flops / (ms * 1e-3) / 1e9

    # roofline: algorithmic bytes summed over ranks (each rank reads all of B), peak x ranks
    bm_local = torch.tensor([float(bytes_min(M_loc, N, K, nnz_loc, not args.unvalued))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(bm_local)
    bm = float(bm_local)
    peak, peak_src = measured_peak_gbs()
    achieved = bm / (ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("%s_K%d_%s" % (args.workload, K, "unvalued" if args.unvalued else "valued"))
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                "traffic": traffic, "traffic_unit": "bytes per step, ncu dram__bytes_read.sum + dram__bytes_write.sum of both kernels (profiles/traffic.json)",
                "peak_source": peak_src, "bytes_min": bm,
                "frac_of_8TBs_nominal": achieved / (8000.0 * world),
                # secondary, labelled: no-B-reuse model, every nonzero gathers its own B row from DRAM
                "bytes_gather_model": 4.0 * (M + world) + 4.0 * nnz * (1 if args.unvalued else 2) + 4.0 * nnz * K + 4.0 * M * K}
    roofline["gather_model_gbs"] = roofline["bytes_gather_model"] / (ms * 1e-3) / 1e9

    # e2e: HOST (pinned) inputs -> device, operator, C back to the host, per step, per rank.
    # Steps alternate between two streams / two buffer sets so that step i's D2H of C overlaps step
    # i+1's H2D of its inputs (PCIe is full duplex); every step still copies all of its inputs in and
    # all of its result out.
    e2e = None
    if not args.no_e2e:
        h_rp, h_ci = sh.rowptr.cpu().pin_memory(), sh.colind.cpu().pin_memory()
        h_val = None if sh.val is None else sh.val.cpu().pin_memory()
        h_B = B.cpu().pin_memory()
        h2d = h_rp.numel() * 4 + h_ci.numel() * 4 + (0 if h_val is None else h_val.numel() * 4) + h_B.numel() * 4
        d2h = M_loc * K * 4
        del C
        sets = []
        for _ in range(2):
            sets.append({
                "stream": torch.cuda.Stream(), "rp": torch.empty_like(sh.rowptr), "ci": torch.empty_like(sh.colind),
                "val": None if sh.val is None else torch.empty_like(sh.val), "B": torch.empty_like(B),
                "hC": torch.empty(M_loc, K, dtype=torch.float32).pin_memory()})

        def e2e_step(i):
            st = sets[i % 2]
            with torch.cuda.stream(st["stream"]):
                st["rp"].copy_(h_rp, non_blocking=True); st["ci"].copy_(h_ci, non_blocking=True)
                if st["val"] is not None:
                    st["val"].copy_(h_val, non_blocking=True)
                st["B"].copy_(h_B, non_blocking=True)
                o = (spmm.csr_spmm_no_edge_value(st["rp"], st["ci"], st["B"]) if st["val"] is None
                     else spmm.csr_spmm(st["rp"], st["ci"], st["val"], st["B"]))
                st["hC"].copy_(o, non_blocking=True)

        e2e_steps = max(4, min(args.steps, 10))
        for i in range(2):
            e2e_step(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record(stream)
        for st in sets:
            st["stream"].wait_event(e0)
        for i in range(e2e_steps):
            e2e_step(i)
        for st in sets:
            stream.wait_stream(st["stream"])
        e1.record(stream)
        torch.cuda.synchronize()
        e2e_ms = e0.elapsed_time(e1) / e2e_steps
        assert torch.equal(sets[0]["hC"], sets[1]["hC"])
        hb = torch.tensor([float(h2d), float(d2h)], device=dev, dtype=torch.float64)
        if world > 1:
            t = torch.tensor([e2e_ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_ms = float(t)
            dist.all_reduce(hb)
        e2e = {"value": flops / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(hb[0]),
               "d2h_bytes_per_step": int(hb[1]), "ms_per_step": e2e_ms, "steps": e2e_steps,
               "pcie_gbs_each_way": [float(hb[0]) / world / (e2e_ms * 1e-3) / 1e9, float(hb[1]) / world / (e2e_ms * 1e-3) / 1e9],
               "api": "spmm.csr_spmm on pinned host tensors copied in, C copied out to pinned host memory (per rank: its "
                      "row block + all of B); steps alternate over two streams so D2H of one overlaps H2D of the next"}
        C = sets[0]["hC"]  # host copy of the local result, for the bitwise check below
        del h_B, sets

    out = {
        "metric": "spmm_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        # cit-Patents K=128, GE-SpMM cachec2_2 = 140.4 GFLOP/s (matrix_id_info.xlsx DL19; the real
        # graph, GPU unstated, single GPU).  Only the default workload at K=128 has a published counterpart.
        "vs_baseline": (value / 140.4) if (args.workload == "citpatents" and K == 128 and args.scale == 1.0) else None,
        "vs_baseline_note": "BASELINE.md: 140.4 GFLOP/s = GE-SpMM cachec2_2 on the real cit-Patents, K=128, unstated 2019-era GPU; "
                            "this run is the seeded synthetic shape-alike of the same N and nnz",
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "K": K, "M": M, "N": N, "nnz": nnz, "valued": not args.unvalued,
                   "scale": args.scale, "sharding": "nnz-balanced contiguous row blocks, B replicated by one NCCL broadcast before the timed region" if world > 1 else "none",
                   "l2": "inputs larger than L2 (B + C = %.2f GB per rank vs 126 MB)" % ((N * K + M_loc * K) * 4 / 1e9),
                   "degree_stats": stats},
        "roofline": roofline, "e2e": e2e, "gpu_launches": args.steps * (2 if nnz_loc > 4096 else 1), "clocks": clocks,
        "b_broadcast_ms": bcast_ms if world > 1 else None, "per_rank_ms": per_rank_ms,
        "b_layout": ("row-sharded, remote rows gathered over NVLink inside the kernel (no replication); fraction of gathers "
                     "that are remote: %.3f" % remote_frac) if remote_frac is not None else "replicated on every rank",
        "per_rank_rows_nnz": per_rank_shape,
    }

    if rank == 0:
        # reference kernel on the same GPU ("the bar on the B200 box"), whole matrix, rank 0 only
        if not args.no_ref_kernel:
            oracle = entry.load_oracle()
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
                    same = bool(torch.equal(Crl, Cl))
                    short = ((rowptr_full[1:] - rowptr_full[:-1]) <= 4096)[sh.row_lo:sh.row_hi]
                    same_short = bool(torch.equal(Crl[short], Cl[short]))
                    maxdiff = float((Crl - Cl).abs().max())
                    out["reference_kernel_same_gpu"] = {
                        "kernel": "spmm_test2<float>, tile_row 8 (spmm_test.cu:161-236, 756), 1 GPU, whole matrix",
                        "ms": rms, "value": flops / (rms * 1e-3) / 1e9, "unit": "GFLOP/s", "bitwise_equal_to_ours": same,
                        "bitwise_equal_on_rows_up_to_GESPMM_LONG_ROW": same_short, "max_abs_diff": maxdiff}
                    del Cr, ones
                except Exception as e:
                    out["reference_kernel_same_gpu"] = {"error": repr(e)}
            else:
                out["reference_kernel_same_gpu"] = {"unavailable": "oracle/_ref not built or int32 offsets would overflow"}
        if not args.no_cpu and world == 1:  # rank 0 at N = 1 only
            out["cpu_baseline"] = cpu_legs(rowptr_full, colind_full, B, K)
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
