"""Row-block sharding of the CSR across GPUs + replication of the dense operand.

The reference is single-GPU (no NCCL/MPI call site anywhere); this module is
the multi-GPU half of BASELINE.json's north_star.  The path shards naturally by output rows:

    C[rows_p, :] = A[rows_p, :] @ B            p = 0..P-1, one process per GPU

* rows are split into P contiguous blocks balanced on key(r) = rowptr[r] + r (nonzeros
  gathered + rows stored -- the same cost model the kernel's task windows use), the block's
  rowptr is rebased to 0, colind stays global;
* every rank needs all of B (for power-law / random graphs a row block references every
  column), so B is replicated with ONE collective: ``broadcast`` from the owner, or
  ``all_gather`` when B is itself row-sharded (the state between GNN layers, since C is
  produced row-sharded);
* C needs no reduction (disjoint row blocks).

torch.distributed is the plumbing (backend nccl on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def partition_rows(rowptr, parts):
    """Boundaries b[0..parts] (b[0]=0, b[parts]=M) of contiguous row blocks balanced on rowptr[r]+r."""
    rp = rowptr.to(torch.int64)
    M = rp.numel() - 1
    key = rp + torch.arange(M + 1, dtype=torch.int64, device=rp.device)
    total = int(key[-1])
    targets = torch.tensor([(total * p) // parts for p in range(1, parts)], dtype=torch.int64, device=rp.device)
    inner = torch.searchsorted(key, targets).clamp_(max=M).tolist() if parts > 1 else []
    bounds = [0] + inner + [M]
    for i in range(1, len(bounds)):  # monotone even for degenerate inputs
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


def shard_csr(rowptr, colind, val, lo, hi):
    """Rows [lo, hi) as a CSR of their own: rebased rowptr, global column ids."""
    s, e = int(rowptr[lo]), int(rowptr[hi])
    rp = (rowptr[lo:hi + 1] - rowptr[lo]).to(torch.int32).contiguous()
    ci = colind[s:e].contiguous()
    v = None if val is None else val[s:e].contiguous()
    return rp, ci, v


def _cuda_spmm(rowptr, colind, val, B):
    from .op import spmm  # fails loudly if the extension is missing; rejects CPU tensors
    return spmm.csr_spmm_no_edge_value(rowptr, colind, B) if val is None else spmm.csr_spmm(rowptr, colind, val, B)


class RowShardedSpMM:
    """One rank's share of C = A @ B.

    ``rowptr/colind/val`` are the FULL matrix (any device); the constructor keeps only this
    rank's row block on ``device``.  ``spmm_fn(rowptr, colind, val, B) -> C`` defaults to the
    CUDA operator; the CPU (gloo) tests inject the oracle there to check the plumbing.
    """

    def __init__(self, rowptr, colind, val, n_cols, rank=None, world=None, device=None, group=None, spmm_fn=None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.device = torch.device(device) if device is not None else rowptr.device
        self.N = int(n_cols)
        self.M = rowptr.numel() - 1
        self.bounds = partition_rows(rowptr, self.world)
        self.row_lo, self.row_hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        rp, ci, v = shard_csr(rowptr, colind, val, self.row_lo, self.row_hi)
        self.rowptr, self.colind = rp.to(self.device), ci.to(self.device)
        self.val = None if v is None else v.to(self.device)
        self.nnz_local = int(self.colind.numel())
        self.spmm_fn = spmm_fn or _cuda_spmm

    # -- replication of the dense operand -------------------------------------------------------
    def broadcast_B(self, B, K, root=0):
        """B lives on ``root`` (pass None elsewhere); every rank gets the full [N, K]."""
        if self.rank != root or B is None:
            B = torch.empty(self.N, K, dtype=torch.float32, device=self.device)
        else:
            B = B.to(self.device).contiguous()
        if self.world > 1:
            dist.broadcast(B, src=root, group=self.group)
        return B

    def b_row_bounds(self):
        """Row blocks of a row-sharded B: the same boundaries as C's when A is square, else even."""
        if self.N == self.M:
            return self.bounds
        return [(self.N * p) // self.world for p in range(self.world + 1)]

    def all_gather_B(self, B_local, out=None):
        """B is row-sharded with b_row_bounds(); returns the full [N, K] on every rank."""
        K = B_local.shape[1]
        bb = self.b_row_bounds()
        full = out if out is not None else torch.empty(self.N, K, dtype=torch.float32, device=self.device)
        if self.world == 1:
            full.copy_(B_local)
            return full
        # uneven blocks: all_gather into per-rank views of the one output buffer (NCCL writes in place)
        views = [full[bb[p]:bb[p + 1]] for p in range(self.world)]
        sizes = {v.shape[0] for v in views}
        if len(sizes) == 1 and dist.get_backend(self.group) == "nccl":
            dist.all_gather_into_tensor(full, B_local.contiguous(), group=self.group)
        else:
            views[self.rank].copy_(B_local)
            works = [dist.broadcast(views[p], src=p, group=self.group, async_op=True)
                     for p in range(self.world) if views[p].shape[0] > 0]
            for w in works:
                w.wait()
        return full

    # -- the local product -----------------------------------------------------------------------
    def forward(self, B_full):
        """C[row_lo:row_hi, :] for this rank."""
        return self.spmm_fn(self.rowptr, self.colind, self.val, B_full)

    # -- no replication: B stays row-sharded, the kernel gathers remote rows over NVLink -------------
    def share_B_parts(self, B_local):
        """Exchange CUDA IPC handles of every rank's row block of B (rows b_row_bounds()[p] .. [p+1]).
        Returns the list of blocks as tensors: this rank's own tensor and peer-mapped views of the others
        (torch opens the handles with lazy peer access, so kernels on this device can read them over
        NVLink).  The owners must keep their blocks alive while the views are in use."""
        if self.world == 1:
            return [B_local]
        from torch.multiprocessing.reductions import reduce_tensor
        B_local = B_local.contiguous()
        payload = [None] * self.world
        dist.all_gather_object(payload, reduce_tensor(B_local), group=self.group)
        parts = []
        for q, (rebuild, args) in enumerate(payload):
            parts.append(B_local if q == self.rank else rebuild(*args))
        self._ipc_owner = B_local  # keep our exported block alive
        return parts

    def forward_sharded_B(self, parts, stream=None):
        """C[row_lo:row_hi, :] = A[row block] @ B with B given as row blocks (see share_B_parts): one fused
        kernel computes and pulls the remote B rows over NVLink; nothing is replicated."""
        from . import capi
        bb = self.b_row_bounds()
        K = next(p.shape[1] for p in parts if p.shape[0] > 0)
        M_loc = self.row_hi - self.row_lo
        C = torch.empty(M_loc, K, dtype=torch.float32, device=self.device)
        for q, p in enumerate(parts):
            assert p.dtype == torch.float32 and p.is_contiguous() and p.shape[0] == bb[q + 1] - bb[q]
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        with torch.cuda.device(self.device):
            capi.csr_spmm_f32_bparts(M_loc, self.N, K, self.nnz_local, self.rowptr.data_ptr(), self.colind.data_ptr(),
                                     None if self.val is None else self.val.data_ptr(),
                                     [p.data_ptr() if p.shape[0] else 0 for p in parts], bb, K, C.data_ptr(), K, st)
        return C

    def gather_C(self, C_local, dst=0):
        """Assemble the full C on ``dst`` (tests / validation only; the product leaves C sharded)."""
        K = C_local.shape[1]
        if self.world == 1:
            return C_local
        if self.rank == dst:
            full = torch.empty(self.M, K, dtype=C_local.dtype, device=C_local.device)
            full[self.row_lo:self.row_hi] = C_local
            for p in range(self.world):
                if p != dst and self.bounds[p + 1] > self.bounds[p]:
                    dist.recv(full[self.bounds[p]:self.bounds[p + 1]], src=p, group=self.group)
            return full
        if self.row_hi > self.row_lo:
            dist.send(C_local.contiguous(), dst=dst, group=self.group)
        return None
