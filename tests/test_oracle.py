"""The oracle against everything that pins it: the reference's own reader compiled from
/root/reference (when present), the committed golden fixtures, the pure-numpy restatement,
closed-form identities, and (BASELINE.json config 1) torch.sparse.mm on the CPU."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

REF_DATA = "/root/reference/data/misc"
BUNDLED = ("cora", "citeseer", "pubmed")


def _rand_csr(rng, M, N, nnz):
    rows = np.sort(rng.integers(0, M, nnz))
    cols = rng.integers(0, N, nnz).astype(np.int32)
    rowptr = np.zeros(M + 1, np.int32)
    np.add.at(rowptr, rows + 1, 1)
    return np.cumsum(rowptr).astype(np.int32), cols


# ---- reader ----------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", BUNDLED)
def test_reader_restatement_matches_reference_reader_on_bundled(oracle, name, known_answers):
    path = os.path.join(REF_DATA, name + ".mtx")
    if not (os.path.exists(path) and oracle.have_ref(oracle.REF_READMTX)):
        pytest.skip("needs /root/reference and oracle/_ref (authoring container only)")
    nr, nc, r, c, v = oracle.ref_read_mtx(path)
    nr2, nc2, r2, c2, v2 = oracle.read_mtx(path)
    assert (nr, nc) == (nr2, nc2) == (known_answers[name]["nrows"], known_answers[name]["ncols"])
    assert len(r) == known_answers[name]["nnz"]
    assert np.array_equal(r, r2) and np.array_equal(c, c2)


@pytest.mark.parametrize("name", BUNDLED)
def test_golden_csr_known_answers(golden_csr, known_answers, name):
    rowptr, colind, shape = golden_csr(name)
    ka = known_answers[name]
    deg = np.diff(rowptr)
    assert shape == (ka["nrows"], ka["ncols"])
    assert rowptr[0] == 0 and rowptr[-1] == len(colind) == ka["nnz"]
    assert deg.max() == ka["max_degree"] and (deg == 0).sum() == ka["empty_rows"]
    # readMtx post-conditions for a symmetric file: sorted, no self-loops, no duplicates, symmetric
    rows = np.repeat(np.arange(shape[0]), deg)
    key = rows.astype(np.int64) * shape[1] + colind
    assert (np.diff(key) > 0).all()
    assert (rows != colind).all()
    assert np.array_equal(np.sort(colind.astype(np.int64) * shape[0] + rows), key)


def test_reader_restatement_on_edge_cases(oracle, expected_mtx):
    for fname, exp in expected_mtx.items():
        nr, nc, r, c, v = oracle.read_mtx(os.path.join(GOLDEN, fname))
        assert (nr, nc) == (exp["nrows"], exp["ncols"]), fname
        assert r.tolist() == exp["row"] and c.tolist() == exp["col"], fname
        if "symmetric" not in fname:  # the reference leaves values misaligned after makeSymmetric (util.hpp:268-283)
            got = sorted(zip(r.tolist(), c.tolist(), v.tolist()))
            want = sorted(zip(exp["row"], exp["col"], exp["val"]))
            assert got == want, fname


def test_edge_case_fixtures_cover_reference_reader(oracle, expected_mtx):
    """expected_mtx.json is regenerated from the reference reader: must match it when it is here."""
    if not oracle.have_ref(oracle.REF_READMTX):
        pytest.skip("oracle/_ref not built")
    files = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "edge_*.mtx")))
    assert files == sorted(expected_mtx)
    for fname in files:
        if not expected_mtx[fname]["reference_defined"]:
            continue  # reference reads past the end of its vectors on this input (see the file's comment)
        nr, nc, r, c, v = oracle.ref_read_mtx(os.path.join(GOLDEN, fname))
        assert r.tolist() == expected_mtx[fname]["row"] and c.tolist() == expected_mtx[fname]["col"], fname


def test_coo_to_csr(oracle):
    row = np.array([0, 0, 2, 2, 2, 5], np.int32)
    col = np.array([1, 4, 0, 2, 3, 5], np.int32)
    indptr, indices, data = oracle.coo_to_csr(6, row, col)
    assert indptr.tolist() == [0, 2, 2, 5, 5, 5, 6]
    assert indices.tolist() == col.tolist()
    assert (data == 1.0).all()  # spmm_test.cu:574


# ---- SpMM ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", BUNDLED)
def test_oracle_checksums_on_golden(oracle, golden_csr, known_answers, name):
    rowptr, colind, shape = golden_csr(name)
    K = 32
    B = oracle.fill_B_cli(shape[1] * K, 1).reshape(shape[1], K)
    C = oracle.spmm(rowptr, colind, np.ones(len(colind), np.float32), B, fma=True)
    ka = known_answers[name]
    assert float(C.astype(np.float64).sum()) == ka["C_K32_seed1_sum_f64"]
    assert float(np.abs(C.astype(np.float64)).sum()) == ka["C_K32_seed1_abs_sum_f64"]
    # A == 1: valued (either rounding mode) and unvalued must agree bit for bit
    assert np.array_equal(C, oracle.spmm(rowptr, colind, None, B))
    assert np.array_equal(C, oracle.spmm(rowptr, colind, np.ones(len(colind), np.float32), B, fma=False))


def test_c_restatement_equals_numpy_restatement(oracle):
    rng = np.random.default_rng(0)
    rowptr, colind = _rand_csr(rng, 40, 30, 300)
    val = rng.standard_normal(300).astype(np.float32)
    B = rng.standard_normal((30, 7)).astype(np.float32)
    for fma in (False, True):
        assert np.array_equal(oracle.spmm(rowptr, colind, val, B, fma=fma), oracle.spmm_numpy(rowptr, colind, val, B, fma=fma))
    assert np.array_equal(oracle.spmm(rowptr, colind, None, B), oracle.spmm_numpy(rowptr, colind, None, B))
    # threading does not change the result (rows are independent)
    assert np.array_equal(oracle.spmm(rowptr, colind, val, B, nthreads=1), oracle.spmm(rowptr, colind, val, B, nthreads=4))


def test_identities(oracle):
    rng = np.random.default_rng(1)
    M = N = 257
    rowptr, colind = _rand_csr(rng, M, N, 4000)
    # B == 1  =>  C[r, :] = degree(r)   (Gunrock's validation identity, gunrock-test/app/spmm/spmm_test.cuh:119-141)
    C = oracle.spmm(rowptr, colind, None, np.ones((N, 16), np.float32))
    assert np.array_equal(C, np.repeat(np.diff(rowptr)[:, None], 16, 1).astype(np.float32))
    # A == I  =>  C == B bitwise
    B = rng.standard_normal((N, 33)).astype(np.float32)
    eye_ptr, eye_ind = np.arange(N + 1, dtype=np.int32), np.arange(N, dtype=np.int32)
    assert np.array_equal(oracle.spmm(eye_ptr, eye_ind, None, B), B)
    assert np.array_equal(oracle.spmm(eye_ptr, eye_ind, np.ones(N, np.float32), B), B)
    # empty rows give zeros; an empty matrix gives all zeros
    z = oracle.spmm(np.zeros(M + 1, np.int32), np.zeros(0, np.int32), None, B)
    assert z.shape == (M, 33) and not z.any()


def test_max_reduce_restatement(oracle):
    """dgl-custom/binary_reduce_max.cu semantics: start -10000, acc > x ? acc : x, empty rows keep the start value."""
    rng = np.random.default_rng(3)
    rowptr, colind = _rand_csr(rng, 50, 40, 300)
    rowptr[10:] -= rowptr[10] - rowptr[9]  # make row 9 empty by shifting (keeps monotone)
    rowptr = np.maximum.accumulate(rowptr).astype(np.int32)
    colind = colind[: rowptr[-1]]
    B = rng.standard_normal((40, 9)).astype(np.float32)
    C = oracle.spmm_max(rowptr, colind, None, B)
    for r in range(50):
        cols = colind[rowptr[r]:rowptr[r + 1]]
        want = np.maximum(B[cols].max(0), -10000.0) if len(cols) else np.full(9, -10000.0, np.float32)
        assert np.array_equal(C[r], want.astype(np.float32))
    assert np.array_equal(oracle.spmm_max(rowptr, colind, None, B - 20000.0), np.full((50, 9), -10000.0, np.float32))  # the sentinel's flaw
    Cinf = oracle.spmm_max(rowptr, colind, None, B - 20000.0, init=-np.inf)
    nonempty = np.diff(rowptr) > 0
    assert np.isneginf(Cinf[~nonempty]).all() and (Cinf[nonempty] < -10000).all()


def test_fp32_within_tolerance_of_fp64_golden(oracle):
    rng = np.random.default_rng(2)
    rowptr, colind = _rand_csr(rng, 500, 400, 60000)
    val = rng.standard_normal(60000).astype(np.float32)
    B = (rng.integers(0, 100, (400, 64)) - 50).astype(np.float32) / 100
    C = oracle.spmm(rowptr, colind, val, B)
    G, mag = oracle.spmm_f64(rowptr, colind, val, B)
    assert (np.abs(C - G) <= 1e-4 * np.maximum(np.abs(G), mag) + 1e-30).all()


def test_reference_dispatch_table(oracle):
    # spmm_kernel.cu:186-205
    assert oracle.ref_dispatch(1000, 16) == (0, (125, 1), (16, 8), 0)
    assert oracle.ref_dispatch(1000, 32) == (1, (250, 1), (32, 4), 512)
    assert oracle.ref_dispatch(1000, 63) == (1, (250, 2), (32, 4), 512)
    assert oracle.ref_dispatch(1000, 64) == (2, (125, 1), (32, 8), 1024)
    assert oracle.ref_dispatch(3774768, 128) == (2, (471846, 2), (32, 8), 1024)  # SURVEY 8a, cit-Patents


def test_fill_B_cli_value_set(oracle):
    B = oracle.fill_B_cli(10000, 1)
    assert np.array_equal(B, oracle.fill_B_cli(10000, 1))
    q = np.round(B * 100)
    assert np.abs(q - B * 100).max() < 1e-4 and q.min() >= -50 and q.max() <= 49


def test_config1_torch_sparse_mm_cpu(oracle):
    """BASELINE.json configs[0]: random CSR 1k x 1k, nnz = 10k, K = 32, fp32, torch.sparse.mm on CPU."""
    g = torch.Generator().manual_seed(0)
    M = N = 1000
    flat = torch.randperm(M * N, generator=g)[:10000].sort().values
    rows, cols = flat // N, flat % N
    rowptr = torch.zeros(M + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=M), 0)
    B = torch.rand(N, 32, generator=g) - 0.5
    A = torch.sparse_csr_tensor(rowptr, cols, torch.ones(10000), size=(M, N))
    C_torch = torch.sparse.mm(A, B).numpy()
    rp, ci = rowptr.numpy().astype(np.int32), cols.numpy().astype(np.int32)
    C_oracle = oracle.spmm(rp, ci, None, B.numpy())
    G, mag = oracle.spmm_f64(rp, ci, None, B.numpy())
    for C in (C_torch, C_oracle):
        assert (np.abs(C - G) <= 1e-4 * np.maximum(np.abs(G), mag) + 1e-30).all()
    assert np.abs(C_torch - C_oracle).max() <= 1e-5
