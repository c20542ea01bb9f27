#!/usr/bin/env python
"""Run bin/spmm_test on the reference's three bundled matrices (written back to .mtx from tests/golden)
with the reference kernel as the baseline cell; prints the CLI's JSON lines.  GPU box."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

entry.load_package()
from gespmm_b200 import build, graphs  # noqa: E402

oracle = entry.load_oracle()
with tempfile.TemporaryDirectory() as d:
    for name in ("cora", "citeseer", "pubmed"):
        z = np.load(os.path.join(ROOT, "tests", "golden", name + "_csr.npz"))
        path = os.path.join(d, name + ".mtx")
        graphs.write_mtx(path, z["rowptr"], z["colind"])
        cmd = [build.CLI, path, "0", "--json", "--iters", "200", "--out", os.path.join(d, "out.csv")]
        if oracle.have_ref(oracle.REF_CLI_KERNELS):
            cmd += ["--baseline-lib", oracle.REF_CLI_KERNELS]
        out = subprocess.run(cmd + sys.argv[1:], stdout=subprocess.PIPE, text=True).stdout
        print("\n".join(l for l in out.splitlines() if l.startswith("{")), flush=True)
