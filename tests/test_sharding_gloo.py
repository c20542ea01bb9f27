"""world_size-2 (and 3) runs of the multi-GPU host logic on CPU with the gloo backend:
row-block sharding, replication of B by broadcast and by all_gather of a row-sharded B,
and assembly of C.  The local product is injected (the oracle) because the product itself has
no CPU path; what is under test is the plumbing around it."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, valued, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    entry.load_package()
    oracle = entry.load_oracle()
    from gespmm_b200 import graphs
    from gespmm_b200.sharding import RowShardedSpMM
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, nnz, K = 3000, 40000, 24
        rowptr, colind = graphs.rmat(N=N, nnz=nnz, seed=11)          # same seed on every rank -> same graph
        val = torch.rand(nnz, generator=torch.Generator().manual_seed(5)) - 0.5 if valued else None

        def oracle_fn(rp, ci, v, B):
            return torch.from_numpy(oracle.spmm(rp.numpy(), ci.numpy(), None if v is None else v.numpy(), B.numpy(), nthreads=1))

        sh = RowShardedSpMM(rowptr, colind, val, N, spmm_fn=oracle_fn)
        assert sh.rank == rank and sh.world == world
        # (1) B owned by rank 0, NCCL-style broadcast
        B0 = graphs.cli_dense(N, K, seed=9) if rank == 0 else None
        B = sh.broadcast_B(B0, K, root=0)
        C_local = sh.forward(B)
        assert C_local.shape == (sh.row_hi - sh.row_lo, K)
        full = sh.gather_C(C_local, dst=0)
        # (2) B row-sharded like C (the state between layers): all_gather, then the same product
        bb = sh.b_row_bounds()
        B2 = sh.all_gather_B(B[bb[rank]:bb[rank + 1]].clone())
        assert torch.equal(B2, B)
        C_local2 = sh.forward(B2)
        assert torch.equal(C_local2, C_local)
        # (3) B row-sharded in EQUAL blocks (what each rank uploads in the end-to-end path): one all_gather_into_tensor
        #     into a padded buffer whose first N rows are B (gloo has no all_gather_into_tensor: emulate it with all_gather)
        blk = sh.even_b_block()
        assert blk * world >= N > blk * (world - 1)
        lo, hi = min(N, rank * blk), min(N, (rank + 1) * blk)
        mine = torch.zeros(blk, K)
        mine[: hi - lo] = B[lo:hi]
        if dist.get_backend() == "gloo":
            pieces = [torch.empty(blk, K) for _ in range(world)]
            dist.all_gather(pieces, mine)
            B3 = torch.cat(pieces)[:N]
        else:
            B3 = sh.all_gather_B_even(mine)
        assert torch.equal(B3, B)
        # (4) the rank's compulsory share of B and its longest row
        assert sh.distinct_b_rows() == int(torch.unique(sh.colind).numel()) <= N
        assert sh.max_row_nnz == -1   # only measured for device-resident blocks (the CUDA path)
        if rank == 0:
            want = oracle.spmm(rowptr.numpy(), colind.numpy(), None if val is None else val.numpy(), B.numpy())
            assert np.array_equal(full.numpy(), want)
            open(os.path.join(out_dir, "ok_%d_%d" % (world, int(valued))), "w").write("ok")
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,valued", [(2, False), (2, True), (3, True)])
def test_row_sharded_spmm_gloo(tmp_path, world, valued):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, valued, str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(tmp_path / ("ok_%d_%d" % (world, int(valued))))
