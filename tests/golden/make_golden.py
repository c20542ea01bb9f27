"""Regenerates tests/golden/*.npz and expected_mtx.json.  Run HERE (the authoring container),
where /root/reference exists and oracle/_ref has been built (`make -C oracle ref`):

    python tests/golden/make_golden.py

What it records, and from where:
  * {cora,citeseer,pubmed}_csr.npz : rowptr/colind of the reference's three bundled matrices
    (/root/reference/data/misc/*.mtx) as parsed by THE REFERENCE'S OWN READER
    (oracle/_ref/libref_readmtx.so = util/util.hpp readMtx<float>, compiled unmodified) and
    converted COO->CSR as spmm_test.cu:557-581 does.  The GPU box has no /root/reference;
    these files are how the real graphs travel.
  * expected_mtx.json : for every hand-made edge_*.mtx in this directory, the (row, col, val)
    triplets the reference reader returns, so that the oracle restatement and the product
    reader can be pinned to the reference without it.
  * known_answers.json : nnz / max degree / empty rows of the bundled matrices,
    and fp32 checksums of C = A @ B (A == 1, B = CLI recipe with seed 1, K = 32) computed by the
    oracle restatement in the reference's summation order.
"""
import glob
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

REF_DATA = "/root/reference/data/misc"
REFERENCE_UB = {"edge_symmetric_tail_selfloop.mtx"}


def main():
    assert os.path.isdir(REF_DATA), "run where /root/reference exists"
    assert oracle.have_ref(oracle.REF_READMTX), "make -C oracle ref first"
    known = {}
    for name in ("cora", "citeseer", "pubmed"):
        nr, nc, r, c, v = oracle.ref_read_mtx(os.path.join(REF_DATA, name + ".mtx"))
        indptr, indices, data = oracle.coo_to_csr(nr, r, c)
        np.savez_compressed(os.path.join(HERE, name + "_csr.npz"), rowptr=indptr, colind=indices,
                            shape=np.array([nr, nc], np.int64))
        deg = np.diff(indptr)
        K = 32
        B = oracle.fill_B_cli(nc * K, 1).reshape(nc, K)
        C = oracle.spmm(indptr, indices, data, B, fma=True)
        known[name] = {
            "nrows": nr, "ncols": nc, "nnz": int(len(r)), "max_degree": int(deg.max()),
            "empty_rows": int((deg == 0).sum()),
            "C_K32_seed1_sum_f64": float(C.astype(np.float64).sum()),
            "C_K32_seed1_abs_sum_f64": float(np.abs(C.astype(np.float64)).sum()),
            "C_K32_seed1_crc": int(np.bitwise_xor.reduce(C.view(np.uint32).ravel().astype(np.uint64) *
                                                         (np.arange(C.size, dtype=np.uint64) % 65521 + 1))),
        }
    expected = {}
    for path in sorted(glob.glob(os.path.join(HERE, "edge_*.mtx"))):
        name = os.path.basename(path)
        if name in REFERENCE_UB:
            # the reference reader is undefined here (see the file's comment): record the documented
            # post-condition (mirror, sort, drop self-loops and duplicates) via the restatement
            nr, nc, r, c, v = oracle.read_mtx(path)
        else:
            # fresh process per file: the reference reader's heap behaviour must not leak between files
            out = subprocess.run([sys.executable, "-c",
                                  "import sys, json; sys.path.insert(0, %r); from oracle import oracle; "
                                  "nr, nc, r, c, v = oracle.ref_read_mtx(%r); "
                                  "print(json.dumps([nr, nc, r.tolist(), c.tolist(), [float(x) for x in v]]))" % (ROOT, path)],
                                 check=True, stdout=subprocess.PIPE, text=True).stdout
            nr, nc, r, c, v = json.loads(out)
            r, c, v = np.array(r), np.array(c), np.array(v)
        expected[name] = {"nrows": nr, "ncols": nc, "row": r.tolist(), "col": c.tolist(),
                          "val": [float(x) for x in v], "reference_defined": name not in REFERENCE_UB}
    with open(os.path.join(HERE, "expected_mtx.json"), "w") as f:
        json.dump(expected, f, indent=1)
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(known, f, indent=1)
    print(json.dumps(known, indent=1))
    for k, e in expected.items():
        print(k, e["nrows"], e["ncols"], list(zip(e["row"], e["col"], e["val"])))


if __name__ == "__main__":
    main()
