"""SPMMFunction and GCNConv: mirror of the reference's pytorch-custom/op.py.

Same names, signatures and argument meaning as op.py:8-36 (SPMMFunction) and op.py:77-152
(GCNConv); the SpMM itself is the sm_100a kernel behind ``spmm.csr_spmm`` /
``spmm.csr_spmm_no_edge_value`` (csrc/spmm.cpp -> include/gespmm.h).

Differences, all deliberate:
  * the extension is prebuilt in-tree (``build.py``) instead of JIT-compiled at import
    (op.py:6); if it is missing it is built once with ``make``; if that fails the import
    fails -- there is no fallback implementation;
  * ``glorot`` / ``zeros`` are restated here (op.py:75 imports them from torch_geometric,
    which this image does not have);
  * forward does not stash ``feat`` in ctx (op.py:16 keeps [N,K] alive for nothing);
  * ``normalize=False`` works (op.py:131-134 calls ``rowptr.shape(0)``, a TypeError);
  * the "[I] Treat edge weight as no_grad." notice (op.py:31) is printed once, not per call;
  * ``GCNConv(..., fuse_norm=True)`` (new, off by default) runs the two degree-normalisation
    passes around the aggregation and the bias add (op.py:142-147: ``x * out_deg_norm`` before,
    ``* in_deg_norm`` and ``+ bias`` after -- three extra passes over [N, K]) INSIDE the kernel:
    the gathered row of x is scaled by its column's out_deg_norm, the finished row by in_deg_norm
    and offset by the bias on its way out (gespmm_opts.row_scale / col_scale / bias), each a
    separately rounded fp32 operation in the unfused order, so the fused layer's output is
    bit-identical to the unfused one's; the unvalued kernel stays unvalued (no per-edge weights);
  * the summation order is a per-call choice (``set_sequential``), not an environment variable:
    by default K <= 64 uses the faster re-associating walker (within ~2e-6 of max|out| of the
    reference's order, deterministic); ``set_sequential(True)`` makes every row of at most 4096
    nonzeros bit-identical to the reference kernels at every K.
"""
import importlib
import importlib.util
import math
import os
import sys

import torch
from torch.nn import Parameter

from . import build as _build


def _load_spmm():
    if not os.path.exists(_build.EXT):
        _build.build()
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("spmm", _build.EXT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules.setdefault("gespmm_b200.spmm", mod)
    return mod


spmm = _load_spmm()

_warned_no_grad = False
_sequential = False


def set_sequential(on=True):
    """Per-process default of this module's calls for GESPMM_FLAG_SEQUENTIAL (see include/gespmm.h): sum every row of
    at most 4096 nonzeros in the reference's strictly sequential CSR order, for every K.  Returns the old value."""
    global _sequential
    old, _sequential = _sequential, bool(on)
    return old


def _product(rowptr, colind, val, feat, row_scale=None, col_scale=None, bias=None):
    if _sequential or row_scale is not None or col_scale is not None or bias is not None:
        return spmm.csr_spmm_ex(rowptr, colind, val, feat, sequential=_sequential, row_scale=row_scale,
                                col_scale=col_scale, bias=bias)
    return spmm.csr_spmm_no_edge_value(rowptr, colind, feat) if val is None else spmm.csr_spmm(rowptr, colind, val, feat)


class SPMMFunction(torch.autograd.Function):
    """out = A @ feat with A given as CSR (forward) and CSC (backward: A^T @ grad_out)."""

    @staticmethod
    def forward(ctx, rowptr, colind, colptr, rowind, feat, edge_weight_csr=None, edge_weight_csc=None):
        out = _product(rowptr, colind, edge_weight_csr, feat)
        ctx.backward_csc = (colptr, rowind, edge_weight_csr, edge_weight_csc)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        global _warned_no_grad
        colptr, rowind, edge_weight_csr, edge_weight_csc = ctx.backward_csc
        grad_out = grad_out.contiguous()
        if edge_weight_csr is not None:
            if edge_weight_csc is None:
                raise RuntimeError(
                    "Backward of SPMM require edge values in both src-first and dst-first order, "
                    "and do not support gradients for edge values. Call with SPMMFunction.apply(rowptr, colind, "
                    "colptr, rowind, in_feat, edge_value_row_first, edge_value_col_first")
            grad_feat = _product(colptr, rowind, edge_weight_csc, grad_out)
            if not _warned_no_grad:
                print("[I] Treat edge weight as no_grad.")
                _warned_no_grad = True
        else:
            grad_feat = _product(colptr, rowind, None, grad_out)
        return None, None, None, None, grad_feat, None, None


def _fuse_gather_scale(nnz, n_feat_rows, K):
    """Whether the per-gathered-row scale should ride inside the kernel or run as its own pass over feat.
    Inside, it costs a 4-byte gather, a shuffle and K/32 multiplies per NONZERO; outside, one read + write of [N, K].
    Measured on B200 (profiles/r02_gcn_layer_fused.txt): at 4 nonzeros per row the fused layer is 2.3-3.9x the unfused
    one, at 50 per row it wins at K = 128 (DRAM-bound walker) and loses a little at K <= 64 (instruction-bound walkers),
    at ~500 per row it loses 20-25 %.  The row scale and the bias are always fused: they cost nothing per nonzero.
    Same bits either way."""
    per_row = nnz / max(1, n_feat_rows)
    return per_row < 16 or (K > 64 and per_row < 128)


class FusedSPMMFunction(torch.autograd.Function):
    """out = (A @ (feat * col_scale)) * row_scale + bias in ONE kernel (gespmm_opts.row_scale / col_scale / bias), i.e.
    GCNConv's ``x * out_deg_norm`` -> SPMMFunction -> ``* in_deg_norm`` -> ``+ bias`` (op.py:142-147) without the three
    element-wise passes; bit-identical to them.  Backward: grad_feat = (A^T @ (grad_out * row_scale)) * col_scale, the
    same fused kernel on the CSC arrays with the two scales swapped; grad_bias = column sums of grad_out.  The scales
    (functions of the graph) and the edge weights get no gradient, like the reference's edge weights."""

    @staticmethod
    def forward(ctx, rowptr, colind, colptr, rowind, feat, row_scale, col_scale, bias=None, edge_weight_csr=None,
                edge_weight_csc=None):
        if edge_weight_csr is not None and edge_weight_csc is None:
            raise RuntimeError("edge values are needed in both src-first and dst-first order (see SPMMFunction)")
        rs = None if row_scale is None else row_scale.reshape(-1).contiguous()
        cs = None if col_scale is None else col_scale.reshape(-1).contiguous()
        b = None if bias is None else bias.detach().contiguous()
        if cs is not None and not _fuse_gather_scale(colind.numel(), feat.shape[0], feat.shape[1]):
            out = _product(rowptr, colind, edge_weight_csr, feat * cs[:, None], rs, None, b)   # dense graph: scale feat in one pass
        else:
            out = _product(rowptr, colind, edge_weight_csr, feat, rs, cs, b)
        ctx.backward_csc = (colptr, rowind, edge_weight_csc, rs, cs, bias is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        colptr, rowind, edge_weight_csc, rs, cs, has_bias = ctx.backward_csc
        grad_out = grad_out.contiguous()
        grad_feat = None
        if ctx.needs_input_grad[4]:  # (A^T @ (grad_out * row_scale)) * col_scale: the same kernel, the two scales swapped
            if rs is not None and not _fuse_gather_scale(rowind.numel(), grad_out.shape[0], grad_out.shape[1]):
                grad_feat = _product(colptr, rowind, edge_weight_csc, grad_out * rs[:, None], cs, None, None)
            else:
                grad_feat = _product(colptr, rowind, edge_weight_csc, grad_out, cs, rs, None)
        grad_bias = grad_out.sum(dim=0) if (has_bias and ctx.needs_input_grad[7]) else None
        return None, None, None, None, grad_feat, None, None, grad_bias, None, None


def glorot(tensor):
    """torch_geometric.nn.inits.glorot: U(-a, a), a = sqrt(6 / (fan_in + fan_out))."""
    if tensor is not None:
        stdv = math.sqrt(6.0 / (tensor.size(-2) + tensor.size(-1)))
        tensor.data.uniform_(-stdv, stdv)


def zeros(tensor):
    if tensor is not None:
        tensor.data.fill_(0)


class GCNConv(torch.nn.Module):
    """x' = D_in^-1/2 A D_out^-1/2 (x W) + b, aggregation through SPMMFunction (op.py:77-152)."""

    def __init__(self, in_channels, out_channels, improved=False, cached=False, bias=True, normalize=True,
                 fuse_norm=False, **kwargs):
        super().__init__()
        self.fuse_norm = fuse_norm
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.improved = improved
        self.cached = cached
        self.normalize = normalize
        self.weight = Parameter(torch.empty(in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        glorot(self.weight)
        zeros(self.bias)
        self.cached_result = None
        self.cached_num_edges = None

    @staticmethod
    def in_deg_sqrt(indptr):
        return (1 / torch.sqrt((indptr[1:] - indptr[:-1]).float())).unsqueeze(dim=1)

    @staticmethod
    def out_deg_sqrt(indptr):
        return (1 / torch.sqrt((indptr[1:] - indptr[:-1]).float())).unsqueeze(dim=1)

    def forward(self, x, rowptr, colind, colptr, rowind, edge_weight_csr=None, edge_weight_csc=None):
        x = torch.matmul(x, self.weight)
        if not self.cached or self.cached_result is None:
            if self.normalize:
                in_deg_norm = self.in_deg_sqrt(rowptr)
                out_deg_norm = self.out_deg_sqrt(colptr)
            else:
                in_deg_norm = torch.ones(rowptr.shape[0] - 1, 1, dtype=x.dtype, device=x.device)
                out_deg_norm = torch.ones(colptr.shape[0] - 1, 1, dtype=x.dtype, device=x.device)
            self.cached_result = in_deg_norm, out_deg_norm
        in_deg_norm, out_deg_norm = self.cached_result
        if self.normalize and self.fuse_norm:
            # one kernel: gathered rows scaled by out_deg_norm, finished rows by in_deg_norm, bias added on the way out
            return FusedSPMMFunction.apply(rowptr, colind, colptr, rowind, x, in_deg_norm, out_deg_norm, self.bias,
                                           edge_weight_csr, edge_weight_csc)
        if self.normalize:
            x = x * out_deg_norm
        aggr_out = SPMMFunction.apply(rowptr, colind, colptr, rowind, x, edge_weight_csr, edge_weight_csc)
        if self.normalize:
            aggr_out = aggr_out * in_deg_norm
        if self.bias is not None:
            aggr_out = aggr_out + self.bias
        return aggr_out

    def __repr__(self):
        return "{}({}, {})".format(self.__class__.__name__, self.in_channels, self.out_channels)
