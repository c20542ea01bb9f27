"""B200-native CSR x dense SpMM behind the operator surface of hgyhungry/ge-spmm.

The directory is named ``ge-spmm_b200`` (not an importable identifier); load it with
``__graft_entry__.load_package()`` or put this directory's parent loader on the path --
it registers itself as the package ``gespmm_b200``.

Layout (only what the hot path needs):
  csrc/        CUDA kernels + C ABI (``include/gespmm.h``), the ``spmm`` PyTorch extension
               shim, the ``spmm_test`` CLI, the .mtx reader
  op.py        ``SPMMFunction`` / ``GCNConv`` -- mirror of the reference's pytorch-custom/op.py
  capi.py      ctypes binding of ``lib/libgespmm.so`` (no torch types)
  sharding.py  row-block sharding of the CSR + NCCL replication of B for 2/4/8 GPUs
  graphs.py    seeded synthetic CSR generators (shape-alikes of the BASELINE.json graphs)
  build.py     in-tree build (``make`` in this directory)

There is no CPU fallback anywhere in this package: every compute entry point needs the
compiled extension and a CUDA device and fails loudly otherwise.
"""
from . import build  # noqa: F401  (cheap: no torch import, no compilation at import time)

__all__ = ["build", "capi", "op", "sharding", "graphs"]
