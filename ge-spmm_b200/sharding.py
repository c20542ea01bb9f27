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


def _cuda_spmm(rowptr, colind, val, B, max_row_nnz=-1):
    from .op import spmm  # fails loudly if the extension is missing; rejects CPU tensors
    if max_row_nnz >= 0:  # a per-graph figure: lets the call skip the long-row kernel when no row is long
        return spmm.csr_spmm_ex(rowptr, colind, val, B, max_row_nnz=max_row_nnz)
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
        self.spmm_fn = spmm_fn
        # the longest row of this rank's block, once per graph (gespmm_max_row_nnz): products over a block without
        # long rows launch one kernel instead of two
        self.max_row_nnz = -1
        if spmm_fn is None and self.rowptr.is_cuda:
            from .op import spmm
            self.max_row_nnz = int(spmm.max_row_nnz(self.rowptr))

    def even_b_block(self):
        """Rows per rank when B is row-sharded in equal blocks (the last one may be short): ceil(N / world)."""
        return -(-self.N // self.world)

    def all_gather_B_even(self, B_block, out=None):
        """B row-sharded in EQUAL blocks of even_b_block() rows (rank p holds rows [p*blk, (p+1)*blk), zero-padded past
        N): ONE all_gather_into_tensor assembles all of B on every rank; returns the [N, K] view of the padded buffer."""
        blk, K = self.even_b_block(), B_block.shape[1]
        assert B_block.shape[0] == blk and B_block.is_contiguous()
        full = out if out is not None else torch.empty(blk * self.world, K, dtype=B_block.dtype, device=B_block.device)
        if self.world == 1:
            full[:blk].copy_(B_block)
        else:
            dist.all_gather_into_tensor(full, B_block, group=self.group)
        return full[:self.N]

    def distinct_b_rows(self):
        """How many different rows of B this rank's block references (its compulsory share of B)."""
        return int(torch.unique(self.colind).numel()) if self.nnz_local else 0

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
        if self.spmm_fn is not None:
            return self.spmm_fn(self.rowptr, self.colind, self.val, B_full)
        return _cuda_spmm(self.rowptr, self.colind, self.val, B_full, self.max_row_nnz)

    # -- replication overlapped with the product: column panels on a communication stream ----------------
    def forward_replicating(self, B_block, chunks=2, sequential=False, out=None):
        """C[row_lo:row_hi, :] = A[row block] @ B when B CHANGES EVERY STEP and arrives row-sharded in equal blocks
        (``even_b_block()`` rows per rank, zero-padded): the replication is not a separate phase.  The local block is
        re-laid out as ``chunks`` column panels; panel j is all-gathered over NVLink on a communication stream while the
        product of panel j - 1 (a [N, K / chunks] operand with row stride K / chunks, written into columns
        [j K / chunks, (j + 1) K / chunks) of C with row stride K -- the C ABI takes strides) runs on the caller's stream.
        One step then costs about max(all-gather, product) + one panel of the other instead of their sum.
        Two panels of 64 columns are the default: the 64-column walker is as fast per column as the 128-column one, the
        32-column one is not.  K / chunks < 128 takes the narrow-B walkers: pass ``sequential=True`` for the bits of the plain product
        (GESPMM_FLAG_SEQUENTIAL), else rows of more than one nonzero are re-associated (within ~2e-6 of max|C|)."""
        from . import capi
        blk, K = self.even_b_block(), B_block.shape[1]
        assert B_block.shape[0] == blk and K % chunks == 0 and (K // chunks) % 4 == 0
        Kc = K // chunks
        dev = self.device
        M_loc = self.row_hi - self.row_lo
        cur = torch.cuda.current_stream(dev)
        if not hasattr(self, "_comm_stream"):
            self._comm_stream = torch.cuda.Stream(device=dev)
        comm = self._comm_stream
        # panel-major copies: local [chunks, blk, Kc], assembled [chunks, world * blk, Kc]
        local = B_block.view(blk, chunks, Kc).permute(1, 0, 2).contiguous()
        full = torch.empty(chunks, blk * self.world, Kc, dtype=torch.float32, device=dev)
        C = out if out is not None else torch.empty(M_loc, K, dtype=torch.float32, device=dev)
        comm.wait_stream(cur)
        events = []
        with torch.cuda.stream(comm):
            for j in range(chunks):
                if self.world > 1:
                    dist.all_gather_into_tensor(full[j], local[j], group=self.group)
                else:
                    full[j, :blk].copy_(local[j])
                ev = torch.cuda.Event()
                ev.record(comm)
                events.append(ev)
        local.record_stream(comm); full.record_stream(comm)
        opts = capi.opts(sequential=sequential, max_row_nnz=self.max_row_nnz)
        with torch.cuda.device(dev):
            for j in range(chunks):
                cur.wait_event(events[j])
                capi.csr_spmm_f32_ex(M_loc, self.N, Kc, self.nnz_local, self.rowptr.data_ptr(), self.colind.data_ptr(),
                                     None if self.val is None else self.val.data_ptr(), full[j].data_ptr(), Kc,
                                     C.data_ptr() + 4 * j * Kc, K, opts, cur.cuda_stream)
        return C

    # -- no replication: B stays row-sharded, the kernel gathers remote rows over NVLink -------------
    def share_B_parts(self, B_local):
        """Put this rank's row block of B (rows b_row_bounds()[p] .. [p+1]) into an exportable device
        allocation, exchange the CUDA IPC handles, and map the other ranks' blocks for kernels on this device
        (peer access over NVLink is enabled by the mapping).  Returns the blocks' device addresses (ints),
        index = owner rank; valid until release_B_parts()."""
        from . import capi
        B_local = B_local.contiguous()
        if self.world == 1:
            self._parts_keep = (B_local, None, [])
            return [B_local.data_ptr()]
        with torch.cuda.device(self.device):
            # torch's allocator may hand out virtual-memory segments that cudaIpc* cannot export: own cudaMalloc
            own, handle = capi.ipc_alloc(B_local.numel() * 4)
            holder = type("CudaBuf", (), {})()
            holder.__cuda_array_interface__ = {"shape": tuple(B_local.shape), "typestr": "<f4", "data": (own, False), "version": 2}
            block = torch.as_tensor(holder, device=self.device) if B_local.numel() else B_local
            if B_local.numel():
                block.copy_(B_local)
            torch.cuda.synchronize(self.device)
            payload = [None] * self.world
            dist.all_gather_object(payload, handle, group=self.group)
            ptrs, opened = [], []
            for q, h in enumerate(payload):
                if q == self.rank:
                    ptrs.append(own)
                else:
                    opened.append(capi.ipc_open(h))
                    ptrs.append(opened[-1])
        self._parts_keep = (block, own, opened)
        dist.barrier(group=self.group)
        return ptrs

    def release_B_parts(self):
        from . import capi
        keep = getattr(self, "_parts_keep", None)
        if keep:
            torch.cuda.synchronize(self.device)
            if self.world > 1:
                dist.barrier(group=self.group)  # nobody unmaps while a peer may still be reading
            with torch.cuda.device(self.device):
                for base in keep[2]:
                    capi.ipc_close(base)
                if self.world > 1:
                    dist.barrier(group=self.group)  # every peer has unmapped before the owner frees
                if keep[1]:
                    capi.ipc_free(keep[1])
            self._parts_keep = None

    def forward_sharded_B(self, part_ptrs, K, stream=None, force=False):
        """C[row_lo:row_hi, :] = A[row block] @ B with B given as row blocks (see share_B_parts): one fused
        kernel computes and pulls the remote B rows over NVLink; nothing is replicated.
        Measured on 8 x B200: with one peer the remote 512-byte gathers run at ~640 GB/s and the fused kernel
        beats replicate-then-multiply on graphs with locality; with three or more peers they fall to ~130 GB/s per rank and
        replication (all_gather_B_even / forward_replicating) wins everywhere -- so beyond two ranks this path must be
        asked for explicitly (``force=True``)."""
        from . import capi
        if self.world > 2 and not force:
            raise RuntimeError("forward_sharded_B is gated to world_size <= 2 (random 512-byte peer gathers reach only ~130 GB/s "
                               "per rank with >= 3 peers: DESIGN.md section 5); use forward_replicating / all_gather_B_even + forward, "
                               "or pass force=True")
        bb = self.b_row_bounds()
        M_loc = self.row_hi - self.row_lo
        C = torch.empty(M_loc, K, dtype=torch.float32, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        with torch.cuda.device(self.device):
            capi.csr_spmm_f32_bparts(M_loc, self.N, K, self.nnz_local, self.rowptr.data_ptr(), self.colind.data_ptr(),
                                     None if self.val is None else self.val.data_ptr(),
                                     [p if bb[q + 1] > bb[q] else 0 for q, p in enumerate(part_ptrs)], bb, K,
                                     C.data_ptr(), K, st)
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
