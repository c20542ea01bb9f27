"""Two-GPU run of the row-sharded path over NCCL (skipped on a one-GPU box): broadcast of B from rank 0,
all_gather of a row-sharded B, local products through the CUDA operator, C assembled and compared with
the oracle bit for bit (short rows) -- the same plumbing the gloo tests cover on CPU, with the real kernel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    entry.load_package()
    oracle = entry.load_oracle()
    from gespmm_b200 import capi, graphs
    from gespmm_b200.sharding import RowShardedSpMM
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        N, nnz, K = 50_000, 1_500_000, 128
        rowptr, colind = graphs.rmat(N=N, nnz=nnz, seed=11)  # CPU generator: identical on every rank
        val = torch.rand(nnz, generator=torch.Generator().manual_seed(5)) - 0.5
        sh = RowShardedSpMM(rowptr, colind, val, N, device=dev)
        B0 = graphs.cli_dense(N, K, seed=9) if rank == 0 else None
        B = sh.broadcast_B(B0, K, root=0)
        C_local = sh.forward(B)
        bb = sh.b_row_bounds()
        B2 = sh.all_gather_B(B[bb[rank]:bb[rank + 1]].clone())
        assert torch.equal(B2, B)
        # B left row-sharded: blocks exchanged as CUDA IPC handles, remote rows gathered over NVLink by the kernel
        mine = B[bb[rank]:bb[rank + 1]].clone()
        parts = sh.share_B_parts(mine)
        C_fused = sh.forward_sharded_B(parts, K)
        torch.cuda.synchronize()
        assert torch.equal(C_fused, C_local), "sharded-B fused product must equal the replicated-B product bit for bit"
        sh.release_B_parts()
        # B row-sharded in equal blocks: one all-gather; and the replication overlapped with the product, panel by panel
        blk = sh.even_b_block()
        lo, hi = min(N, rank * blk), min(N, (rank + 1) * blk)
        even = torch.zeros(blk, K, device=dev)
        even[: hi - lo] = B[lo:hi]
        assert torch.equal(sh.all_gather_B_even(even), B)
        for chunks in (1, 2, 4):
            C_pipe = sh.forward_replicating(even, chunks=chunks, sequential=True)
            torch.cuda.synchronize()
            short_loc = (sh.rowptr[1:] - sh.rowptr[:-1]) <= capi.LONG_ROW
            assert torch.equal(C_pipe[short_loc], C_local[short_loc]), "pipelined replication (sequential order) must give the plain product's bits"
            assert torch.allclose(C_pipe, C_local, rtol=1e-4, atol=1e-3)
        C_pipe = sh.forward_replicating(even, chunks=4)
        torch.cuda.synchronize()
        assert torch.allclose(C_pipe, C_local, rtol=1e-4, atol=1e-3)
        assert sh.max_row_nnz == int((sh.rowptr[1:] - sh.rowptr[:-1]).max())
        full = sh.gather_C(C_local, dst=0)
        if rank == 0:
            want = oracle.spmm(rowptr.numpy(), colind.numpy(), val.numpy(), B.cpu().numpy())
            short = np.diff(rowptr.numpy()) <= capi.LONG_ROW
            got = full.cpu().numpy()
            assert np.array_equal(got[short], want[short])
            G, mag = oracle.spmm_f64(rowptr.numpy(), colind.numpy(), val.numpy(), B.cpu().numpy())
            assert (np.abs(got - G) <= 1e-4 * np.maximum(np.abs(G), mag) + 1e-30).all()
            open(os.path.join(out_dir, "ok"), "w").write("ok")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_row_sharded_spmm_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")
