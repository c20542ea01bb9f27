"""2-/3-layer GCN training loop on a PubMed-shaped problem: the caller of the hot path that the
reference ships as pytorch-custom/gcn_custom.py (2 layers) and gcn_custom_2layer.py (3 layers).

Same model (GCNConv -> relu -> dropout -> GCNConv -> log_softmax), optimiser (Adam, lr 0.01, weight
decay 5e-4 on the first layer), 200 epochs under torch.autograd.profiler and the same per-epoch log
line.  Planetoid cannot be downloaded here, so:
  * the graph is read from a MatrixMarket file with the library's reader (``--mtx``; e.g. the PubMed
    adjacency the reference bundles as data/misc/pubmed.mtx) or, without one, is a seeded
    PubMed-shaped synthetic graph (19,717 nodes, 88,648 directed edges, symmetric); self-loops are
    added either way (gcn_custom.py:33-34);
  * features (500 columns, row-normalised like T.NormalizeFeatures) and the 3 class labels are
    synthetic: labels come from a planted 2-hop linear teacher, so there is something to learn;
  * the CSC arrays come from spmm.csr2csc on the GPU instead of scipy's tocsc() (gcn_custom.py:39-46).

    python ge-spmm_b200/gcn_custom.py --n-hidden 64 [--layers 3] [--epochs 200] [--fuse-norm] [--profile] [--mtx FILE]
"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def load_problem(device, n_features=500, n_classes=3, seed=0, mtx=None):
    from gespmm_b200 import capi, graphs
    from gespmm_b200.op import spmm
    if mtx is not None:
        _, _, rp, ci, _ = capi.read_mtx(mtx)
        rowptr, colind = torch.from_numpy(rp), torch.from_numpy(ci)
    else:
        rowptr, colind = graphs.social_like(19717, 88648, seed=seed + 11, sigma=1.0, locality=0.3, window=0.01)
    n = rowptr.numel() - 1
    rows = torch.repeat_interleave(torch.arange(n), (rowptr[1:] - rowptr[:-1]).long())
    loop = torch.arange(n)
    rowptr, colind = graphs.coo_to_csr(torch.cat([rows, loop]), torch.cat([colind.long(), loop]), n, n, dedup=True)
    g = {"rowptr": rowptr.to(device), "colind": colind.to(device)}
    nnz = colind.numel()
    g["value_csr"] = torch.ones(nnz, device=device)
    g["colptr"] = torch.empty(n + 1, dtype=torch.int32, device=device)
    g["rowind"] = torch.empty(nnz, dtype=torch.int32, device=device)
    g["value_csc"] = spmm.csr2csc(g["rowptr"], g["colind"], g["colptr"], g["rowind"], g["value_csr"])
    gen = torch.Generator().manual_seed(seed)
    x = torch.rand(n, n_features, generator=gen) * (torch.rand(n, n_features, generator=gen) < 0.1)
    x = x / x.sum(1, keepdim=True).clamp(min=1e-12)
    x = x.to(device)
    teacher = torch.randn(n_features, n_classes, generator=gen).to(device)
    h = spmm.csr_spmm_no_edge_value(g["rowptr"], g["colind"], spmm.csr_spmm_no_edge_value(g["rowptr"], g["colind"], x @ teacher))
    y = h.argmax(1)
    perm = torch.randperm(n, generator=gen).to(device)
    masks = {"train": perm[:60], "val": perm[60:560], "test": perm[560:1560]}  # Planetoid split sizes
    return g, x, y, masks, n_features, n_classes


class Net(torch.nn.Module):
    def __init__(self, n_in, n_hidden, n_out, layers=2, fuse_norm=False):
        super().__init__()
        from gespmm_b200.op import GCNConv
        dims = [n_in] + [n_hidden] * (layers - 1) + [n_out]
        self.convs = torch.nn.ModuleList(GCNConv(a, b, cached=True, normalize=True, fuse_norm=fuse_norm)
                                         for a, b in zip(dims[:-1], dims[1:]))
        self.reg_params = self.convs[0].parameters()
        self.non_reg_params = [p for c in self.convs[1:] for p in c.parameters()]

    def forward(self, x, g):
        a = (g["rowptr"], g["colind"], g["colptr"], g["rowind"], g["value_csr"], g["value_csc"])
        for conv in self.convs[:-1]:
            x = F.dropout(F.relu(conv(x, *a)), training=self.training)
        return F.log_softmax(self.convs[-1](x, *a), dim=1)


def run(n_hidden=64, layers=2, epochs=200, fuse_norm=False, profile=False, device="cuda", log=print, mtx=None):
    entry.load_package()
    device = torch.device(device)
    g, x, y, masks, n_in, n_out = load_problem(device, mtx=mtx)
    torch.manual_seed(0)
    model = Net(n_in, n_hidden, n_out, layers, fuse_norm).to(device)
    opt = torch.optim.Adam([dict(params=model.reg_params, weight_decay=5e-4),
                            dict(params=model.non_reg_params, weight_decay=0)], lr=0.01)

    def train():
        model.train()
        opt.zero_grad()
        loss = F.nll_loss(model(x, g)[masks["train"]], y[masks["train"]])
        loss.backward()
        opt.step()
        return float(loss)

    @torch.no_grad()
    def test():
        model.eval()
        pred = model(x, g).argmax(1)
        return [float((pred[m] == y[m]).float().mean()) for m in (masks["train"], masks["val"], masks["test"])]

    best_val = test_acc = 0.0
    losses = []
    ctx = torch.autograd.profiler.profile(use_device="cuda") if profile else None
    if ctx:
        ctx.__enter__()
    for epoch in range(1, epochs + 1):
        losses.append(train())
        tr, va, te = test()
        if va > best_val:
            best_val, test_acc = va, te
        log("Epoch: {:03d}, Train: {:.4f}, Val: {:.4f}, Test: {:.4f}".format(epoch, tr, best_val, test_acc))
    if ctx:
        ctx.__exit__(None, None, None)
        log(ctx.key_averages().table(sort_by="cuda_time_total", row_limit=12))
    return {"first_loss": losses[0], "last_loss": losses[-1], "best_val": best_val, "test": test_acc}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-hidden", type=int, default=64, help="number of hidden features")
    ap.add_argument("--layers", type=int, default=2, choices=[2, 3])
    ap.add_argument("--epochs", type=int, default=200)
    ap.add_argument("--fuse-norm", action="store_true")
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--mtx", default=None, help="MatrixMarket adjacency (default: a PubMed-shaped synthetic graph)")
    a = ap.parse_args()
    print(run(a.n_hidden, a.layers, a.epochs, a.fuse_norm, a.profile, mtx=a.mtx))
