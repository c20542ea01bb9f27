"""Parity tests proper: the CUDA path (through the `spmm` extension / the C ABI) against the
oracle, against the reference's own kernels compiled from /root/reference (oracle/_ref, when
it travelled), and -- at BASELINE.json's full sizes -- through size-independent properties.

Bar (BASELINE.json north_star): within 1e-4 relative fp32 of the reference.  What is actually
asserted is stronger: BIT-EXACT for every row the library sums in the reference's sequential order
(gespmm_row_sum_is_sequential: rows of at most GESPMM_LONG_ROW nonzeros, unless the sub-warp walker
for K <= 64 is in use), and |diff| <= 1e-4 * max(|ref|, sum|a||b|) for the re-associated rows.
"""
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

RTOL = 1e-4  # BASELINE.json north_star: "within 1e-4 relative fp32"
LONG = 4096  # GESPMM_LONG_ROW (include/gespmm.h); test_host checks capi.LONG_ROW against the header


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def spmm(pkg):
    from gespmm_b200.op import spmm
    return spmm


@pytest.fixture(scope="module")
def ref_ext(oracle):
    if not oracle.have_ref(oracle.REF_EXT):
        pytest.skip("oracle/_ref/ref_spmm not built (needs /root/reference at build time)")
    return oracle.ref_extension()


@pytest.fixture(scope="module")
def ref_cli(oracle):
    if not oracle.have_ref(oracle.REF_CLI_KERNELS):
        pytest.skip("oracle/_ref/libref_cli_kernels.so not built")
    return oracle.ref_cli_kernels()


def _run(spmm, dev, rowptr, colind, val, B):
    rp = torch.as_tensor(rowptr, dtype=torch.int32, device=dev)
    ci = torch.as_tensor(colind, dtype=torch.int32, device=dev)
    Bd = torch.as_tensor(B, dtype=torch.float32, device=dev)
    if val is None:
        out = spmm.csr_spmm_no_edge_value(rp, ci, Bd)
    else:
        out = spmm.csr_spmm(rp, ci, torch.as_tensor(val, dtype=torch.float32, device=dev), Bd)
    torch.cuda.synchronize()
    return out


def _sequential_rows(rowptr, K):
    """Rows the OPERATOR sums in the reference's order (spmm.row_sum_is_sequential: gespmm_row_sum_is_sequential_ex for
    the options the extension passes -- it hands the library a padding workspace for widths that are not multiples of 4)."""
    from gespmm_b200.op import spmm as q
    deg = np.diff(rowptr)
    assert q.row_sum_is_sequential(K, 1) and q.row_sum_is_sequential(K, 0)
    assert not q.row_sum_is_sequential(K, LONG + 1)
    if q.row_sum_is_sequential(K, 2):
        assert q.row_sum_is_sequential(K, LONG)
        return deg <= LONG
    return deg <= 1


def _check(oracle, rowptr, colind, val, B, C):
    """bit-exact on sequentially summed rows, 1e-4 of sum|a||b| on the others; returns number of long rows."""
    C = C.cpu().numpy()
    want = oracle.spmm(rowptr, colind, val, B, fma=True)
    seq = _sequential_rows(rowptr, B.shape[1])
    assert C.shape == want.shape
    assert np.array_equal(C[seq], want[seq]), "sequentially summed rows must be bit-identical to the oracle"
    if not seq.all():
        G, mag = oracle.spmm_f64(rowptr, colind, val, B)
        err = np.abs(C.astype(np.float64) - G)
        assert (err <= RTOL * np.maximum(np.abs(G), mag) + 1e-30).all()
    return int((np.diff(rowptr) > LONG).sum())


def _check_with(oracle, rowptr, colind, val, B, C, sequential):
    """_check for a call whose summation order the caller states itself (C ABI calls with their own options)."""
    C = C.cpu().numpy()
    want = oracle.spmm(rowptr, colind, val, B, fma=True)
    deg = np.diff(rowptr)
    seq = (deg <= LONG) if sequential else (deg <= 1)
    assert np.array_equal(C[seq], want[seq])
    G, mag = oracle.spmm_f64(rowptr, colind, val, B)
    assert (np.abs(C.astype(np.float64) - G) <= RTOL * np.maximum(np.abs(G), mag) + 1e-30).all()


def _rand_csr(rng, M, N, nnz, empty_frac=0.0):
    rows = rng.integers(0, M, nnz)
    if empty_frac:
        keep = rng.random(M) >= empty_frac
        alive = np.flatnonzero(keep)
        rows = alive[rng.integers(0, len(alive), nnz)]
    rows = np.sort(rows)
    cols = rng.integers(0, N, nnz).astype(np.int32)
    rowptr = np.zeros(M + 1, np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return np.cumsum(rowptr).astype(np.int32), cols


# ---- golden graphs (the reference's bundled matrices, via tests/golden) ----------------------------

@pytest.mark.parametrize("name", ["cora", "citeseer", "pubmed"])
@pytest.mark.parametrize("K", [16, 32, 33, 64, 100, 128, 256, 512, 640])
def test_bundled_graphs_match_oracle_and_reference_kernels(spmm, dev, oracle, golden_csr, request, name, K):
    rowptr, colind, shape = golden_csr(name)
    rng = np.random.default_rng(K)
    B = oracle.fill_B_cli(shape[1] * K, seed=K).reshape(shape[1], K)   # the CLI's recipe (spmm_test.cu:592-594)
    val = rng.standard_normal(len(colind)).astype(np.float32)
    for v in (None, val, np.ones(len(colind), np.float32)):
        C = _run(spmm, dev, rowptr, colind, v, B)
        _check(oracle, rowptr, colind, v, B, C)
        if oracle.have_ref(oracle.REF_EXT):
            ref = request.getfixturevalue("ref_ext")
            rp, ci, Bd = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, B))
            R = ref.csr_spmm_no_edge_value(rp, ci, Bd) if v is None else ref.csr_spmm(rp, ci, torch.as_tensor(v, device=dev), Bd)
            torch.cuda.synchronize()
            seq = torch.from_numpy(_sequential_rows(rowptr, K)).to(dev)
            assert torch.equal(C[seq], R[seq]), "must be bit-identical to pytorch-custom/spmm_kernel.cu on the same inputs"


def test_reference_cli_kernels_agree_bitwise(spmm, dev, oracle, golden_csr, ref_cli):
    """spmm_test0..4<float> via spmmWrapper (spmm_test.cu:456-492), tile_row 4 as under VALIDATE (:688)."""
    rowptr, colind, shape = golden_csr("pubmed")
    K = 256
    B = oracle.fill_B_cli(shape[1] * K, seed=3).reshape(shape[1], K)
    ones = np.ones(len(colind), np.float32)
    C = _run(spmm, dev, rowptr, colind, ones, B)
    rp, ci, v, Bd = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, ones, B))
    for method in range(5):
        R = torch.full((shape[0], K), float("nan"), device=dev)
        rc = ref_cli.ref_spmm_wrapper(method, 4, shape[0], K, rp.data_ptr(), ci.data_ptr(), v.data_ptr(), Bd.data_ptr(), R.data_ptr())
        torch.cuda.synchronize()
        assert rc == 0
        assert torch.equal(C, R), "method %d" % method


# ---- shapes, tails, empties ------------------------------------------------------------------------

@pytest.mark.parametrize("K", [1, 2, 3, 4, 7, 8, 31, 36, 60, 96, 127, 132, 192, 384, 516, 1024, 1100])
def test_k_sweep_random_with_empty_rows(spmm, dev, oracle, K):
    rng = np.random.default_rng(100 + K)
    M, N, nnz = 3001, 2777, 40000
    rowptr, colind = _rand_csr(rng, M, N, nnz, empty_frac=0.4)
    B = rng.standard_normal((N, K)).astype(np.float32)
    val = rng.standard_normal(nnz).astype(np.float32)
    for v in (None, val):
        _check(oracle, rowptr, colind, v, B, _run(spmm, dev, rowptr, colind, v, B))


def test_degenerate_shapes(spmm, dev, oracle):
    rng = np.random.default_rng(5)
    B = rng.standard_normal((10, 128)).astype(np.float32)
    # all rows empty: every C element must still be written (zeros), like the reference
    C = _run(spmm, dev, np.zeros(1001, np.int32), np.zeros(0, np.int32), None, B)
    assert C.shape == (1000, 128) and not C.any()
    # zero rows
    C = _run(spmm, dev, np.zeros(1, np.int32), np.zeros(0, np.int32), None, B)
    assert C.shape == (0, 128)
    # one row, one column, one nonzero
    C = _run(spmm, dev, np.array([0, 1], np.int32), np.array([0], np.int32), np.array([2.5], np.float32), np.ones((1, 1), np.float32))
    assert C.item() == 2.5
    # a single row holding everything, N == 1
    C = _run(spmm, dev, np.array([0, 777], np.int32), np.zeros(777, np.int32), None, np.full((1, 8), 0.5, np.float32))
    assert (C == 388.5).all()
    # trailing and leading empty rows around one dense row
    rowptr = np.array([0] * 40 + [50] * 61, np.int32)
    colind = np.arange(50, dtype=np.int32) % 10
    _check(oracle, rowptr, colind, None, B, _run(spmm, dev, rowptr, colind, None, B))


@pytest.mark.parametrize("K", [32, 128, 200, 512])
def test_long_rows_segmented_path(spmm, dev, oracle, K):
    """Rows above GESPMM_LONG_ROW nonzeros: deterministic segmented sum, within tolerance; neighbours bit-exact."""
    rng = np.random.default_rng(K)
    M, N = 600, 5000
    deg = rng.integers(0, 12, M)
    deg[[0, 17, 18, 300, 301, 599]] = [70000, LONG + 1, LONG, 12345, 32768, 40001]  # 32768+: whole-cluster path
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    colind = rng.integers(0, N, rowptr[-1]).astype(np.int32)
    B = oracle.fill_B_cli(N * K, seed=1).reshape(N, K)
    val = rng.standard_normal(rowptr[-1]).astype(np.float32)
    for v in (None, val):
        C1 = _run(spmm, dev, rowptr, colind, v, B)
        assert _check(oracle, rowptr, colind, v, B, C1) == 5
        C2 = _run(spmm, dev, rowptr, colind, v, B)
        assert torch.equal(C1, C2), "segmented path must be deterministic"


@pytest.mark.parametrize("K", [4, 8, 12, 16, 20, 24, 32, 36, 48, 60, 64])
def test_subwarp_walker_for_narrow_B(spmm, dev, oracle, pkg, gespmm_env, K):
    """GESPMM_VARIANT=2: 2 / 4 / 8 nonzeros per warp-wide gather for K <= 64 / 32 / 16.  Integer-valued operands make
    every fp32 sum exact, so the result must equal the oracle bit for bit whatever the association; real-valued
    operands must stay within 1e-4 of the fp64 golden, be deterministic, and rows of <= 1 nonzero stay bit-exact.
    Empty rows, rows ending at every position of a quad, long (> 4096) and huge (>= 32768) rows, max-reduce."""
    from gespmm_b200 import capi
    gespmm_env.delenv("GESPMM_SEQUENTIAL", raising=False)   # (it would win over GESPMM_VARIANT)
    gespmm_env.setenv("GESPMM_VARIANT", "2")
    assert not capi.row_sum_is_sequential(K, 2) and capi.row_sum_is_sequential(K, 1)
    assert capi.row_sum_is_sequential(128, LONG)  # wider products are not affected
    rng = np.random.default_rng(500 + K)
    M, N = 2500, 3000
    deg = rng.integers(0, 9, M)
    deg[rng.random(M) < 0.3] = 0
    deg[[5, 6, 900, 2499]] = [40000, 4097, 700, 33000]
    deg[1000:1100] = rng.integers(20, 200, 100)
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    nnz = int(rowptr[-1])
    colind = rng.integers(0, N, nnz).astype(np.int32)
    Bi = rng.integers(-8, 9, (N, K)).astype(np.float32)
    vi = rng.integers(-2, 3, nnz).astype(np.float32)
    for v in (None, vi):
        C = _run(spmm, dev, rowptr, colind, v, Bi).cpu().numpy()
        assert np.array_equal(C, oracle.spmm(rowptr, colind, v, Bi, fma=True)), "exact (integer) sums must not depend on the order"
    Bf = rng.standard_normal((N, K)).astype(np.float32)
    vf = rng.standard_normal(nnz).astype(np.float32)
    for v in (None, vf):
        C1 = _run(spmm, dev, rowptr, colind, v, Bf)
        assert _check(oracle, rowptr, colind, v, Bf, C1) == 3
        assert torch.equal(C1, _run(spmm, dev, rowptr, colind, v, Bf)), "must be deterministic"
    rp, ci, v, Bd = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, vf, Bf))
    for vv in (None, v):
        C = torch.full((M, K), float("nan"), device=dev)
        capi.csr_spmm_max_f32(M, N, K, nnz, rp.data_ptr(), ci.data_ptr(), None if vv is None else vv.data_ptr(), Bd.data_ptr(), K,
                              C.data_ptr(), K, -10000.0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(C.cpu().numpy(), oracle.spmm_max(rowptr, colind, None if vv is None else vf, Bf, init=-10000.0))
    # not a multiple of 4: up to K = 16 the same walker runs on 4-byte slices (re-associated, within tolerance; exact on
    # integer-valued operands), above that the sequential scalar walker (bit-identical)
    B3 = rng.standard_normal((N, K - 1)).astype(np.float32)
    assert capi.row_sum_is_sequential(K - 1, 2) == (K - 1 > 16) == spmm.row_sum_is_sequential(K - 1, 2)
    assert spmm.row_sum_is_sequential(K - 1, LONG, True)
    for v in (None, vf):
        _check(oracle, rowptr, colind, v, B3, _run(spmm, dev, rowptr, colind, v, B3))
    B3i = rng.integers(-8, 9, (N, K - 1)).astype(np.float32)
    for v in (None, vi):
        assert np.array_equal(_run(spmm, dev, rowptr, colind, v, B3i).cpu().numpy(), oracle.spmm(rowptr, colind, v, B3i, fma=True))


@pytest.mark.parametrize("K", [4, 8, 16, 24, 32, 36, 48, 64])
def test_row_parallel_walker_for_narrow_B_is_sequential(spmm, dev, oracle, pkg, gespmm_env, K):
    """GESPMM_VARIANT=4: the lane groups own disjoint rows, each summed in CSR order -- bit-identical to the oracle
    (real-valued operands) on every row up to GESPMM_LONG_ROW, whatever the balance of the 32-row runs: empty rows,
    single-row runs between long rows, rows much longer than their neighbours; max-reduce bit-identical everywhere."""
    from gespmm_b200 import capi
    gespmm_env.setenv("GESPMM_VARIANT", "4")
    assert capi.row_sum_is_sequential(K, LONG) and not capi.row_sum_is_sequential(K, LONG + 1)
    rng = np.random.default_rng(700 + K)
    M, N = 2600, 3000
    deg = rng.integers(0, 9, M)
    deg[rng.random(M) < 0.3] = 0
    deg[[5, 6, 7, 900, 2599]] = [40000, 4097, 4100, 3000, 33000]   # row 6 sits alone between two long rows
    deg[1000:1100] = rng.integers(20, 400, 100)
    deg[1500:1532] = 0
    deg[1510] = 1
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    nnz = int(rowptr[-1])
    colind = rng.integers(0, N, nnz).astype(np.int32)
    Bf = rng.standard_normal((N, K)).astype(np.float32)
    vf = rng.standard_normal(nnz).astype(np.float32)
    for v in (None, vf):
        C1 = _run(spmm, dev, rowptr, colind, v, Bf)
        assert _check(oracle, rowptr, colind, v, Bf, C1) == 4
        assert torch.equal(C1, _run(spmm, dev, rowptr, colind, v, Bf)), "must be deterministic"
    rp, ci, v, Bd = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, vf, Bf))
    for vv in (None, v):
        C = torch.full((M, K), float("nan"), device=dev)
        capi.csr_spmm_max_f32(M, N, K, nnz, rp.data_ptr(), ci.data_ptr(), None if vv is None else vv.data_ptr(), Bd.data_ptr(), K,
                              C.data_ptr(), K, -10000.0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(C.cpu().numpy(), oracle.spmm_max(rowptr, colind, None if vv is None else vf, Bf, init=-10000.0))


def _mixed_graph(rng, M=2600, N=3000):
    """Empty rows, short rows, a block of medium rows, long (> 4096) and huge (>= 32768) rows."""
    deg = rng.integers(0, 9, M)
    deg[rng.random(M) < 0.3] = 0
    deg[[5, 6, 900, M - 1]] = [40000, 4097, 3000, 33000]
    deg[1000:1100] = rng.integers(20, 400, 100)
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    colind = rng.integers(0, N, int(rowptr[-1])).astype(np.int32)
    return rowptr, colind, M, N


@pytest.mark.parametrize("K", [3, 4, 7, 13, 16, 32, 48, 64])
def test_sequential_order_is_a_per_call_option(spmm, dev, oracle, pkg, gespmm_env, K):
    """csr_spmm_ex(sequential=True) -> GESPMM_FLAG_SEQUENTIAL: bit-identical to the oracle on every row up to
    GESPMM_LONG_ROW at the widths whose default walker re-associates; the plain call next to it, same process, same
    environment, is not affected (the order is no longer a process-wide switch)."""
    from gespmm_b200 import capi
    for name in ("GESPMM_VARIANT", "GESPMM_SEQUENTIAL"):   # the process default under test is the library's own
        gespmm_env.delenv(name, raising=False)
    rng = np.random.default_rng(900 + K)
    rowptr, colind, M, N = _mixed_graph(rng)
    Bf = rng.standard_normal((N, K)).astype(np.float32)
    vf = rng.standard_normal(len(colind)).astype(np.float32)
    rp, ci, vd, Bd = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, vf, Bf))
    short = np.diff(rowptr) <= LONG
    assert capi.row_sum_is_sequential(K, LONG, capi.opts(sequential=True)) and not capi.row_sum_is_sequential(K, 2)
    for v, vnp in ((None, None), (vd, vf)):
        Cs = spmm.csr_spmm_ex(rp, ci, v, Bd, sequential=True)
        Cd = spmm.csr_spmm_ex(rp, ci, v, Bd)
        torch.cuda.synchronize()
        want = oracle.spmm(rowptr, colind, vnp, Bf, fma=True)
        assert np.array_equal(Cs.cpu().numpy()[short], want[short])
        plain = spmm.csr_spmm_no_edge_value(rp, ci, Bd) if v is None else spmm.csr_spmm(rp, ci, v, Bd)
        assert torch.equal(Cd, plain)
        _check(oracle, rowptr, colind, vnp, Bf, Cd)


@pytest.mark.parametrize("K", [3, 7, 8, 16, 32, 41, 64, 100, 128, 200, 256])
def test_fused_scales_and_bias_match_the_separate_passes_bitwise(spmm, dev, oracle, pkg, K):
    """gespmm_opts.row_scale / col_scale / bias (GCNConv's passes around the aggregation, pytorch-custom/op.py:142-147):
    out = (A @ (B * cs)) * rs + bias must carry the bits of the four separate passes -- every product and sum is rounded
    separately, in that order -- for every walker (ring, sub-warp, row-parallel, scalar), valued and unvalued, on empty,
    short, long and huge rows, and with any subset of the three vectors."""
    rng = np.random.default_rng(1200 + K)
    rowptr, colind, M, N = _mixed_graph(rng)
    Bf = rng.standard_normal((N, K)).astype(np.float32)
    vf = rng.standard_normal(len(colind)).astype(np.float32)
    rs = (1.0 / np.sqrt(1.0 + np.diff(rowptr))).astype(np.float32)
    cs = (0.25 + rng.random(N)).astype(np.float32)
    bias = rng.standard_normal(K).astype(np.float32)
    rp, ci, vd, Bd, rsd, csd, bd = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, vf, Bf, rs, cs, bias))
    short = np.diff(rowptr) <= LONG
    for sequential in (False, True):
        for v, vnp in ((None, None), (vd, vf)):
            for use in ((1, 1, 1), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0)):
                r_, c_, b_ = (rsd if use[0] else None), (csd if use[1] else None), (bd if use[2] else None)
                got = spmm.csr_spmm_ex(rp, ci, v, Bd, sequential=sequential, row_scale=r_, col_scale=c_, bias=b_)
                x = Bd * csd[:, None] if use[1] else Bd
                y = spmm.csr_spmm_ex(rp, ci, v, x.contiguous(), sequential=sequential)
                if use[0]:
                    y = y * rsd[:, None]
                if use[2]:
                    y = y + bd
                torch.cuda.synchronize()
                assert torch.equal(got, y), "fused != separate passes (K=%d sequential=%s valued=%s use=%s)" % (K, sequential, v is not None, use)
        if spmm.row_sum_is_sequential(K, 2, sequential):   # against the oracle, on the rows summed in CSR order
            xs = (Bf * cs[:, None]).astype(np.float32)
            want = (oracle.spmm(rowptr, colind, vf, xs, fma=True) * rs[:, None]).astype(np.float32) + bias
            got = spmm.csr_spmm_ex(rp, ci, vd, Bd, sequential=sequential, row_scale=rsd, col_scale=csd, bias=bd).cpu().numpy()
            assert np.array_equal(got[short], want[short])


@pytest.mark.parametrize("K", [17, 31, 41, 47, 63, 65, 127, 130])
def test_widths_that_are_not_multiples_of_4_padded_and_unpadded(spmm, dev, oracle, pkg, K):
    """K % 4 != 0 above 16.  Bare C ABI call: the ring walker on 4-byte slices, sequential order, bit-identical to the oracle.
    With gespmm_opts.workspace (what the operator always passes) and a graph dense enough (this one: nnz / (M + N) = 15):
    padded copies of B and C, the 16-byte-slice walkers in sequential order -- the same bits; a sparser graph stays
    unpadded -- the same bits again; strided operands; a workspace that is too small is refused."""
    from gespmm_b200 import capi
    rng = np.random.default_rng(7000 + K)
    rowptr, colind, M, N = _mixed_graph(rng)
    nnz = len(colind)
    Bf = rng.standard_normal((N, K)).astype(np.float32)
    vf = rng.standard_normal(nnz).astype(np.float32)
    rp, ci, vd, Bd = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, vf, Bf))
    st = torch.cuda.current_stream().cuda_stream
    short = np.diff(rowptr) <= LONG
    want = oracle.spmm(rowptr, colind, vf, Bf, fma=True)
    # bare call, no workspace
    assert capi.row_sum_is_sequential(K, LONG)
    C0 = torch.full((M, K), float("nan"), device=dev)
    capi.csr_spmm_f32(M, N, K, nnz, rp.data_ptr(), ci.data_ptr(), vd.data_ptr(), Bd.data_ptr(), K, C0.data_ptr(), K, st)
    torch.cuda.synchronize()
    assert np.array_equal(C0.cpu().numpy()[short], want[short])
    # with a workspace, sequential flag: bit-identical again; default: within tolerance
    need = capi.pad_workspace_bytes(M, N, K, nnz)
    assert need >= 4 * (M + N) * ((K + 3) // 4 * 4) and capi.pad_workspace_bytes(M, N, K + (4 - K % 4), nnz) == 0
    assert capi.pad_workspace_bytes(M, N, K, 4 * (M + N) - 1) == 0   # too sparse for the padding to pay: nothing to allocate
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    for seq in (True, False):
        C1 = torch.full((M, K), float("nan"), device=dev)
        o = capi.opts(sequential=seq, workspace=ws.data_ptr(), workspace_bytes=need)
        capi.csr_spmm_f32_ex(M, N, K, nnz, rp.data_ptr(), ci.data_ptr(), vd.data_ptr(), Bd.data_ptr(), K, C1.data_ptr(), K, o, st)
        torch.cuda.synchronize()
        assert capi.row_sum_is_sequential(K, LONG, o)
        assert np.array_equal(C1.cpu().numpy()[short], want[short])
        _check_with(oracle, rowptr, colind, vf, Bf, C1, True)
    # a graph too sparse for the padding to pay takes the 4-byte-slice walker even with a workspace: same bits
    rp_s, ci_s = _rand_csr(rng, 20000, N, 30000, empty_frac=0.3)
    need_s = capi.pad_workspace_bytes(20000, N, K, 10**9)
    ws_s = torch.empty(need_s, dtype=torch.uint8, device=dev)
    Csp = torch.full((20000, K), float("nan"), device=dev)
    rps, cis = torch.as_tensor(rp_s, device=dev), torch.as_tensor(ci_s, device=dev)
    capi.csr_spmm_f32_ex(20000, N, K, len(ci_s), rps.data_ptr(), cis.data_ptr(), None, Bd.data_ptr(), K, Csp.data_ptr(), K,
                         capi.opts(workspace=ws_s.data_ptr(), workspace_bytes=need_s), st)
    torch.cuda.synchronize()
    assert np.array_equal(Csp.cpu().numpy(), oracle.spmm(rp_s, ci_s, None, Bf))
    with pytest.raises(capi.GespmmError):   # a workspace that is too small is refused, not overrun
        capi.csr_spmm_f32_ex(M, N, K, nnz, rp.data_ptr(), ci.data_ptr(), vd.data_ptr(), Bd.data_ptr(), K, C1.data_ptr(), K,
                             capi.opts(workspace=ws.data_ptr(), workspace_bytes=need - 256), st)
    # the operator (hands the library a workspace on its own): bit-identical with and without the sequential flag
    for Cs in (spmm.csr_spmm_ex(rp, ci, vd, Bd, sequential=True), spmm.csr_spmm(rp, ci, vd, Bd)):
        torch.cuda.synchronize()
        assert np.array_equal(Cs.cpu().numpy()[short], want[short])
        _check(oracle, rowptr, colind, vf, Bf, Cs)
    # strided B and C through the workspace path
    ldb, ldc = K + 5, K + 3
    Bs = torch.zeros(N, ldb, device=dev); Bs[:, :K] = Bd
    Cst = torch.full((M, ldc), -7.0, device=dev)
    capi.csr_spmm_f32_ex(M, N, K, nnz, rp.data_ptr(), ci.data_ptr(), vd.data_ptr(), Bs.data_ptr(), ldb, Cst.data_ptr(), ldc,
                         capi.opts(sequential=True, workspace=ws.data_ptr(), workspace_bytes=need), st)
    torch.cuda.synchronize()
    assert np.array_equal(Cst[:, :K].cpu().numpy()[short], want[short]) and bool((Cst[:, K:] == -7.0).all())


def test_longest_row_lets_the_call_skip_the_long_row_kernel(spmm, dev, oracle, pkg):
    """gespmm_max_row_nnz + gespmm_opts.max_row_nnz: same bits with and without the hint; a graph with long rows still
    gets its long-row kernel when the hint says so."""
    rng = np.random.default_rng(77)
    rowptr, colind = _rand_csr(rng, 20000, 20000, 90000, empty_frac=0.3)
    rowptr2, colind2, M2, N2 = _mixed_graph(rng)
    for rpn, cin, N in ((rowptr, colind, 20000), (rowptr2, colind2, N2)):
        rp, ci = torch.as_tensor(rpn, device=dev), torch.as_tensor(cin, device=dev)
        B = torch.randn(N, 128, device=dev)
        mx = spmm.max_row_nnz(rp)
        assert mx == int(np.diff(rpn).max())
        a = spmm.csr_spmm_ex(rp, ci, None, B, max_row_nnz=mx)
        b = spmm.csr_spmm_no_edge_value(rp, ci, B)
        torch.cuda.synchronize()
        assert torch.equal(a, b)
    assert spmm.max_row_nnz(torch.zeros(1, dtype=torch.int32, device=dev)) == 0


@pytest.mark.parametrize("K", [68, 128, 200, 256, 384])
def test_bulk_walker_tma_gathers_match_the_oracle_bitwise(spmm, dev, oracle, pkg, K):
    """GESPMM_WALKER_BULK: one cp.async.bulk (TMA) per gathered B row, completion on mbarriers.  Same flat stream and
    summation order as the ring walker: bit-identical to the oracle on every row up to GESPMM_LONG_ROW (valued and
    unvalued; empty rows, partial last panels, long and huge rows through the long-row kernel), deterministic, and equal
    to the default walker's result everywhere."""
    from gespmm_b200 import capi
    rng = np.random.default_rng(4000 + K)
    rowptr, colind, M, N = _mixed_graph(rng)
    nnz = len(colind)
    Bf = rng.standard_normal((N, K)).astype(np.float32)
    vf = rng.standard_normal(nnz).astype(np.float32)
    rp, ci, vd, Bd = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, vf, Bf))
    st = torch.cuda.current_stream().cuda_stream
    short = np.diff(rowptr) <= LONG
    for v, vnp in ((None, None), (vd, vf)):
        vp = None if v is None else v.data_ptr()
        C0 = torch.empty(M, K, device=dev); C1 = torch.full((M, K), float("nan"), device=dev); C2 = torch.full((M, K), float("nan"), device=dev)
        capi.csr_spmm_f32(M, N, K, nnz, rp.data_ptr(), ci.data_ptr(), vp, Bd.data_ptr(), K, C0.data_ptr(), K, st)
        for C in (C1, C2):
            capi.csr_spmm_f32_ex(M, N, K, nnz, rp.data_ptr(), ci.data_ptr(), vp, Bd.data_ptr(), K, C.data_ptr(), K,
                                 capi.opts(walker=capi.WALKER_BULK), st)
        torch.cuda.synchronize()
        assert torch.equal(C1, C2), "must be deterministic"
        assert torch.equal(C0, C1), "bulk walker != default walker"
        assert np.array_equal(C1.cpu().numpy()[short], oracle.spmm(rowptr, colind, vnp, Bf, fma=True)[short])
    # a graph of short rows only (no long-row kernel), many tasks
    rowptr2, colind2 = _rand_csr(rng, 30000, N, 150000, empty_frac=0.4)
    rp2, ci2 = torch.as_tensor(rowptr2, device=dev), torch.as_tensor(colind2, device=dev)
    C = torch.full((30000, K), float("nan"), device=dev)
    capi.csr_spmm_f32_ex(30000, N, K, len(colind2), rp2.data_ptr(), ci2.data_ptr(), None, Bd.data_ptr(), K, C.data_ptr(), K,
                         capi.opts(walker=capi.WALKER_BULK, max_row_nnz=int(np.diff(rowptr2).max())), st)
    torch.cuda.synchronize()
    assert np.array_equal(C.cpu().numpy(), oracle.spmm(rowptr2, colind2, None, Bf))


@pytest.mark.parametrize("walker", [0, 5])
@pytest.mark.parametrize("policy", [16, 32, 1 + 4 * 1 + 16 * 1, 2 + 4 * 1 + 16 * 1, 0 + 4 * 1, 3 + 4 * 3 + 16 * 3])
def test_l2_priority_steering_does_not_change_results(dev, oracle, pkg, policy, walker):
    """gespmm_opts.l2_policy: the C stores (every walker) and the gathered rows (bulk walker: a TMA bulk copy takes an L2
    cache-policy operand; cp.async with one traps on sm_100a, so the ring walker ignores the gather priorities) carry L2
    eviction priorities; the sums are the same bits."""
    from gespmm_b200 import capi
    rng = np.random.default_rng(31)
    rowptr, colind, M, N = _mixed_graph(rng)
    nnz = len(colind)
    for K in (128, 200, 256):
        Bf = rng.standard_normal((N, K)).astype(np.float32)
        vf = rng.standard_normal(nnz).astype(np.float32)
        rp, ci, vd, Bd = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, vf, Bf))
        st = torch.cuda.current_stream().cuda_stream
        for v in (None, vd):
            C0 = torch.empty(M, K, device=dev); C1 = torch.full((M, K), float("nan"), device=dev)
            vp = None if v is None else v.data_ptr()
            capi.csr_spmm_f32(M, N, K, nnz, rp.data_ptr(), ci.data_ptr(), vp, Bd.data_ptr(), K, C0.data_ptr(), K, st)
            capi.csr_spmm_f32_ex(M, N, K, nnz, rp.data_ptr(), ci.data_ptr(), vp, Bd.data_ptr(), K, C1.data_ptr(), K,
                                 capi.opts(walker=walker, l2_policy=policy, l2_window_rows=500), st)
            torch.cuda.synchronize()
            assert torch.equal(C0, C1)


def test_skewed_rmat_graph(spmm, dev, oracle, pkg):
    from gespmm_b200 import graphs
    rowptr, colind = graphs.rmat(N=200_000, nnz=4_000_000, seed=4)
    rowptr, colind = rowptr.numpy(), colind.numpy()
    assert np.diff(rowptr).max() > LONG
    K = 64
    B = oracle.fill_B_cli(200_000 * K, seed=2).reshape(200_000, K)
    _check(oracle, rowptr, colind, None, B, _run(spmm, dev, rowptr, colind, None, B))


# ---- raw C ABI: strides, alignment, streams, host buffers --------------------------------------------

def test_c_abi_strides_alignment_and_stream(dev, oracle, pkg):
    from gespmm_b200 import capi
    rng = np.random.default_rng(9)
    M, N, nnz, K = 700, 650, 9000, 128
    rowptr, colind = _rand_csr(rng, M, N, nnz, empty_frac=0.2)
    val = rng.standard_normal(nnz).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    want = torch.from_numpy(oracle.spmm(rowptr, colind, val, B))
    rp, ci, v = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, val))
    stream = torch.cuda.Stream()
    for ldb, ldc, off in [(K, K, 0), (K + 4, K + 8, 0), (K + 3, K + 1, 0), (K, K, 1)]:
        Bbuf = torch.zeros(N * ldb + 4, device=dev)
        Cbuf = torch.full((M * ldc + 4,), float("nan"), device=dev)
        Bview = Bbuf[off:off + N * ldb].view(N, ldb)
        Bview[:, :K] = torch.from_numpy(B).to(dev)
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            capi.csr_spmm_f32(M, N, K, nnz, rp.data_ptr(), ci.data_ptr(), v.data_ptr(),
                              Bbuf.data_ptr() + 4 * off, ldb, Cbuf.data_ptr() + 4 * off, ldc, stream.cuda_stream)
        stream.synchronize()
        Cview = Cbuf[off:off + M * ldc].view(M, ldc)
        assert torch.equal(Cview[:, :K].cpu(), want), (ldb, ldc, off)
        assert torch.isnan(Cview[:, K:]).all(), "padding columns of C must not be written"


@pytest.mark.parametrize("K", [128, 64, 200, 512])
def test_c_abi_b_given_as_row_blocks(dev, oracle, pkg, K):
    """gespmm_csr_spmm_f32_bparts: B split into separately allocated row blocks (here all on one device; across
    GPUs in test_sharding_nccl_gpu.py) gives the same bits as one contiguous B, long and huge rows included."""
    from gespmm_b200 import capi
    rng = np.random.default_rng(31 + K)
    M, N = 900, 4000
    deg = rng.integers(0, 10, M); deg[[3, 500]] = [40000, 6000]
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    nnz = int(rowptr[-1])
    colind = rng.integers(0, N, nnz).astype(np.int32)
    val = rng.standard_normal(nnz).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    rp, ci, v, Bd = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, val, B))
    for bounds in ([0, N], [0, 1000, 1000, 2500, N], [0, 1, 2, 3, 4, 5, 6, 7, N]):
        parts = [Bd[a:b].clone() for a, b in zip(bounds[:-1], bounds[1:])]
        for vv in (None, v):
            C = torch.full((M, K), float("nan"), device=dev)
            capi.csr_spmm_f32_bparts(M, N, K, nnz, rp.data_ptr(), ci.data_ptr(), None if vv is None else vv.data_ptr(),
                                     [p.data_ptr() if p.shape[0] else 0 for p in parts], bounds, K, C.data_ptr(), K,
                                     torch.cuda.current_stream().cuda_stream)
            whole = torch.empty(M, K, device=dev)
            capi.csr_spmm_f32(M, N, K, nnz, rp.data_ptr(), ci.data_ptr(), None if vv is None else vv.data_ptr(), Bd.data_ptr(), K,
                              whole.data_ptr(), K, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            if K > 64 and not (os.environ.get("GESPMM_VARIANT") or os.environ.get("GESPMM_SEQUENTIAL")):
                assert torch.equal(C, whole), (K, bounds)   # the ring walker either way: long rows are segmented identically too
            else:
                seq = torch.from_numpy(_sequential_rows(rowptr, K)).to(dev)
                assert torch.equal(C[seq], whole[seq]), (K, bounds)
            _check(oracle, rowptr, colind, None if vv is None else val, B, whole)
            # the sharded-B walker is always the sequential ring walker
            want = oracle.spmm(rowptr, colind, None if vv is None else val, B, fma=True)
            short = np.diff(rowptr) <= LONG
            assert np.array_equal(C.cpu().numpy()[short], want[short])
    with pytest.raises(capi.GespmmError):  # blocks must cover [0, N)
        capi.csr_spmm_f32_bparts(M, N, K, nnz, rp.data_ptr(), ci.data_ptr(), None, [Bd.data_ptr()], [0, N - 1], K, C.data_ptr(), K)


@pytest.mark.parametrize("K", [8, 33, 64, 128, 200, 512, 640])
def test_max_reduce_matches_oracle_bitwise(dev, oracle, pkg, K):
    """gespmm_csr_spmm_max_f32 (dgl-custom/binary_reduce_max.cu): max is order-independent, so every row --
    short, long (> 4096) and huge (>= 32768, cluster path) -- is bit-identical to the sequential restatement."""
    from gespmm_b200 import capi
    rng = np.random.default_rng(41 + K)
    M, N = 700, 3000
    deg = rng.integers(0, 10, M); deg[[2, 300, 699]] = [40000, 5000, 4097]
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    nnz = int(rowptr[-1])
    colind = rng.integers(0, N, nnz).astype(np.int32)
    val = rng.standard_normal(nnz).astype(np.float32)
    B = (rng.standard_normal((N, K)) * 3).astype(np.float32)
    rp, ci, v, Bd = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, val, B))
    for vv, init in ((None, -10000.0), (None, float("-inf")), (v, -10000.0)):
        C = torch.full((M, K), float("nan"), device=dev)
        capi.csr_spmm_max_f32(M, N, K, nnz, rp.data_ptr(), ci.data_ptr(), None if vv is None else vv.data_ptr(), Bd.data_ptr(), K,
                              C.data_ptr(), K, init, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        want = oracle.spmm_max(rowptr, colind, None if vv is None else val, B, init=init)
        assert np.array_equal(C.cpu().numpy(), want), (K, init)


def test_c_abi_host_buffers(oracle, pkg):
    from gespmm_b200 import capi
    rng = np.random.default_rng(10)
    rowptr, colind = _rand_csr(rng, 500, 400, 6000)
    val = rng.standard_normal(6000).astype(np.float32)
    B = rng.standard_normal((400, 96)).astype(np.float32)
    assert np.array_equal(capi.csr_spmm_host(rowptr, colind, val, B), oracle.spmm(rowptr, colind, val, B))
    assert np.array_equal(capi.csr_spmm_host(rowptr, colind, None, B), oracle.spmm(rowptr, colind, None, B))


def test_cuda_graph_capture(spmm, dev, oracle):
    rng = np.random.default_rng(11)
    rowptr, colind = _rand_csr(rng, 2000, 2000, 30000)
    B = rng.standard_normal((2000, 128)).astype(np.float32)
    rp, ci, Bd = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, B))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        spmm.csr_spmm_no_edge_value(rp, ci, Bd)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = spmm.csr_spmm_no_edge_value(rp, ci, Bd)
    Bd.copy_(torch.from_numpy(B * 2))
    g.replay()
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), oracle.spmm(rowptr, colind, None, B * 2))


def test_concurrent_calls_from_host_threads_and_streams(spmm, dev, oracle):
    """Re-entrancy: four host threads, each on its own stream, run different products at once (each thread gets its
    own helper stream for the long-row kernel); every result matches the oracle."""
    import threading
    rng = np.random.default_rng(77)
    jobs = []
    for t in range(4):
        M, N, K = 1500 + 100 * t, 1200, [64, 128, 256, 128][t]
        deg = rng.integers(0, 20, M); deg[t] = 9000 + t  # one long row each
        rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
        colind = rng.integers(0, N, rowptr[-1]).astype(np.int32)
        B = rng.standard_normal((N, K)).astype(np.float32)
        jobs.append((rowptr, colind, B))
    results, errors = [None] * 4, []

    def work(i):
        try:
            rowptr, colind, B = jobs[i]
            st = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(st):
                rp, ci, Bd = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, B))
                out = None
                for _ in range(20):
                    out = spmm.csr_spmm_no_edge_value(rp, ci, Bd)
            st.synchronize()
            results[i] = out.cpu().numpy()
            from gespmm_b200 import capi
            capi.thread_cleanup()   # this thread's helper stream / events go with it ...
            out2 = spmm.csr_spmm_no_edge_value(rp, ci, Bd)   # ... and are re-created on demand
            torch.cuda.synchronize()
            assert torch.equal(out2, out)
            capi.thread_cleanup()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    for (rowptr, colind, B), got in zip(jobs, results):
        want = oracle.spmm(rowptr, colind, None, B)
        short = _sequential_rows(rowptr, B.shape[1])
        assert np.array_equal(got[short], want[short])
        assert np.allclose(got[~short], want[~short], rtol=1e-4, atol=1e-3)


def test_operator_rejects_bad_arguments(spmm, dev):
    rp = torch.zeros(5, dtype=torch.int32, device=dev); ci = torch.zeros(0, dtype=torch.int32, device=dev)
    B = torch.zeros(4, 8, device=dev)
    with pytest.raises(RuntimeError, match="int32"):
        spmm.csr_spmm_no_edge_value(rp.long(), ci, B)
    with pytest.raises(RuntimeError, match="float32"):
        spmm.csr_spmm_no_edge_value(rp, ci, B.double())
    with pytest.raises(RuntimeError, match="contiguous"):
        spmm.csr_spmm_no_edge_value(rp, ci, torch.zeros(8, 4, device=dev).t())
    with pytest.raises(RuntimeError, match="CUDA"):
        spmm.csr_spmm_no_edge_value(rp, ci, B.cpu())
    with pytest.raises(RuntimeError, match="same length"):
        spmm.csr_spmm(rp, ci, torch.zeros(3, device=dev), B)


# ---- full-size properties (BASELINE.json sizes; the oracle would take minutes) ----------------------

def test_full_size_citpatents_shape_properties(spmm, dev, oracle, pkg):
    """N = 3,774,768, nnz = 16,518,948, K = 128 (configs[1]).  Integer-valued B makes every fp32 sum
    exact, so (i) B == 1 gives the degree, (ii) column sums of C equal in-degree-weighted column sums of
    B computed independently, (iii) a sample of rows equals the oracle bit for bit."""
    from gespmm_b200 import graphs
    N, nnz = graphs.SHAPES["cit-Patents"]
    K = 128
    rowptr, colind = graphs.citation_like(seed=1, device=dev)
    assert rowptr.numel() == N + 1 and colind.numel() == nnz
    deg = (rowptr[1:] - rowptr[:-1]).float()
    ones = torch.ones(N, K, device=dev)
    C = spmm.csr_spmm_no_edge_value(rowptr, colind, ones)
    assert torch.equal(C, deg[:, None].expand(-1, K))
    del C, ones
    g = torch.Generator(device=dev).manual_seed(3)
    B = torch.randint(-8, 9, (N, K), generator=g, device=dev).float()
    val = torch.randint(-2, 3, (nnz,), generator=g, device=dev).float()
    C = spmm.csr_spmm(rowptr, colind, val, B)
    wsum = torch.zeros(N, device=dev, dtype=torch.float64).index_add_(0, colind.long(), val.double())
    assert torch.equal(C.double().sum(0), wsum @ B.double())
    rows = torch.randint(0, N, (2000,), generator=g, device=dev).sort().values
    rp_h, ci_h, v_h, B_h = rowptr.cpu().numpy(), colind.cpu().numpy(), val.cpu().numpy(), B.cpu().numpy()
    for r in rows.cpu().tolist()[::20]:
        s, e = rp_h[r], rp_h[r + 1]
        want = oracle.spmm(np.array([0, e - s], np.int32), ci_h[s:e], v_h[s:e], B_h)
        assert np.array_equal(C[r].cpu().numpy(), want[0])


def _sample_rows_vs_oracle(oracle, rowptr, colind, val, B, C, rows):
    """Rows `rows` of C against the oracle, one small CSR per row (full-size graphs: the whole product is too slow on the CPU)."""
    rp_h, ci_h = rowptr.cpu().numpy(), colind.cpu().numpy()
    v_h = None if val is None else val.cpu().numpy()
    B_h = B.cpu().numpy()
    for r in rows:
        s, e = int(rp_h[r]), int(rp_h[r + 1])
        want = oracle.spmm(np.array([0, e - s], np.int32), ci_h[s:e], None if v_h is None else v_h[s:e], B_h)
        assert np.array_equal(C[r].cpu().numpy(), want[0]), r


def test_full_size_reddit_shape_through_spmmfunction(dev, oracle, pkg, request):
    """BASELINE.json configs[2] at full size: the Reddit shape (232,965 nodes, 114,615,892 edges, symmetric), K = 256,
    forward AND backward through SPMMFunction.apply (op.py:8-36; symmetric graph, so the CSC arrays are the CSR arrays).
    (i) feat == 1: forward gives the degree exactly, and so does the gradient of sum(out); (ii) integer-valued features:
    the column sums of out equal the degree-weighted column sums of feat computed independently, exactly; (iii) real-valued
    features: sampled rows (the longest included) equal the oracle bit for bit up to GESPMM_LONG_ROW nonzeros and within 1e-4
    above; (iv) the reference's own extension, built from its sources, gives the same bits on every row up to GESPMM_LONG_ROW."""
    from gespmm_b200 import graphs
    from gespmm_b200.op import SPMMFunction
    N, nnz = graphs.SHAPES["reddit"]
    K = 256
    rowptr, colind = graphs.reddit_like(seed=2, device=dev)
    assert rowptr.numel() == N + 1 and colind.numel() == nnz
    deg = (rowptr[1:] - rowptr[:-1])
    x = torch.ones(N, K, device=dev, requires_grad=True)
    y = SPMMFunction.apply(rowptr, colind, rowptr, colind, x)
    assert torch.equal(y.detach(), deg.float()[:, None].expand(-1, K))
    y.sum().backward()
    assert torch.equal(x.grad, deg.float()[:, None].expand(-1, K))       # A^T 1 = in-degree = degree (symmetric)
    del x, y
    g = torch.Generator(device=dev).manual_seed(5)
    xi = torch.randint(-4, 5, (N, K), generator=g, device=dev).float()
    yi = SPMMFunction.apply(rowptr, colind, rowptr, colind, xi)
    assert torch.equal(yi.double().sum(0), deg.double() @ xi.double())
    del xi, yi
    xf = torch.randn(N, K, generator=g, device=dev)
    yf = SPMMFunction.apply(rowptr, colind, rowptr, colind, xf)
    order = torch.argsort(deg, descending=True).cpu().tolist()
    short_rows = [r for r in order if int(deg[r]) <= LONG][:3] + torch.randint(0, N, (40,), generator=torch.Generator().manual_seed(1)).tolist()
    _sample_rows_vs_oracle(oracle, rowptr, colind, None, xf, yf, [r for r in short_rows if int(deg[r]) <= LONG])
    for r in order[:2]:                                                    # the two longest rows: segmented, within tolerance
        s, e = int(rowptr[r]), int(rowptr[r + 1])
        G, mag = oracle.spmm_f64(np.array([0, e - s], np.int32), colind[s:e].cpu().numpy(), None, xf.cpu().numpy())
        assert (np.abs(yf[r].cpu().numpy().astype(np.float64) - G[0]) <= RTOL * np.maximum(np.abs(G[0]), mag[0]) + 1e-30).all()
    if oracle.have_ref(oracle.REF_EXT):
        ref = request.getfixturevalue("ref_ext")
        yr = ref.csr_spmm_no_edge_value(rowptr, colind, xf)
        torch.cuda.synchronize()
        short = deg <= LONG
        assert torch.equal(yf[short], yr[short]), "must be bit-identical to pytorch-custom/spmm_kernel.cu at configs[2]'s full size"


@pytest.mark.parametrize("K", [32, 47, 64, 128])
def test_full_size_products_shape_properties(spmm, dev, oracle, pkg, K):
    """BASELINE.json configs[3] at full size: the ogbn-products shape (2,449,029 nodes, ~123.7 M edges), K from the sweep
    plus the graph's real class count (47).  feat == 1 gives the degree exactly (integer sums are exact in any order);
    real-valued features: sampled rows against the oracle, bit for bit where the operator sums in CSR order (K = 47, 128),
    within 1e-4 where it re-associates (K = 32, 64)."""
    from gespmm_b200 import graphs
    rowptr, colind = graphs.products_like(seed=3, device=dev)
    N = rowptr.numel() - 1
    deg = (rowptr[1:] - rowptr[:-1])
    C = spmm.csr_spmm_no_edge_value(rowptr, colind, torch.ones(N, K, device=dev))
    assert torch.equal(C, deg.float()[:, None].expand(-1, K))
    del C
    g = torch.Generator(device=dev).manual_seed(9)
    B = torch.randn(N, K, generator=g, device=dev)
    val = torch.randn(colind.numel(), generator=g, device=dev)
    C = spmm.csr_spmm(rowptr, colind, val, B)
    rows = [r for r in torch.randint(0, N, (60,), generator=torch.Generator().manual_seed(2)).tolist() if int(deg[r]) <= LONG]
    if spmm.row_sum_is_sequential(K, 2):
        _sample_rows_vs_oracle(oracle, rowptr, colind, val, B, C, rows)
    else:
        B_h = B.cpu().numpy()
        for r in rows:
            s, e = int(rowptr[r]), int(rowptr[r + 1])
            G, mag = oracle.spmm_f64(np.array([0, e - s], np.int32), colind[s:e].cpu().numpy(), val[s:e].cpu().numpy(), B_h)
            assert (np.abs(C[r].cpu().numpy().astype(np.float64) - G[0]) <= RTOL * np.maximum(np.abs(G[0]), mag[0]) + 1e-30).all()


def test_full_size_rmat_shape_properties(spmm, dev, oracle, pkg):
    """BASELINE.json configs[4] on one GPU at full size: R-MAT 10M x 10M, 200M nonzeros (duplicates kept), K = 128.
    feat == 1 gives the degree exactly on every row -- the hub row of 275,760 nonzeros (cluster path) included --; sampled
    rows of a real-valued product equal the oracle bit for bit."""
    from gespmm_b200 import graphs
    rowptr, colind = graphs.rmat(seed=4, device=dev)
    N, K = rowptr.numel() - 1, 128
    assert N == 10_000_000 and colind.numel() == 200_000_000
    deg = (rowptr[1:] - rowptr[:-1])
    assert int(deg.max()) > 32768                                           # exercises the whole-cluster long-row path
    C = spmm.csr_spmm_no_edge_value(rowptr, colind, torch.ones(N, K, device=dev))
    assert torch.equal(C, deg.float()[:, None].expand(-1, K))
    del C
    B = torch.randn(N, K, generator=torch.Generator(device=dev).manual_seed(4), device=dev)
    C = spmm.csr_spmm_no_edge_value(rowptr, colind, B)
    rows = [r for r in torch.randint(0, N, (60,), generator=torch.Generator().manual_seed(6)).tolist() if int(deg[r]) <= LONG]
    _sample_rows_vs_oracle(oracle, rowptr, colind, None, B, C, rows)


# ---- csr2csc, autograd, GCNConv ------------------------------------------------------------------------

@pytest.mark.parametrize("M,N,nnz", [(1, 1, 1), (50, 70, 0), (300, 200, 5000), (5000, 70000, 200000), (2000, 3, 30000)])
def test_csr2csc_matches_scipy(spmm, dev, M, N, nnz):
    import scipy.sparse as sp
    rng = np.random.default_rng(M + N)
    rowptr, colind = _rand_csr(rng, M, N, nnz)
    val = rng.standard_normal(nnz).astype(np.float32)
    A = sp.csr_matrix((val, colind, rowptr), shape=(M, N))
    rp, ci, v = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, val))
    colptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    rowind = torch.empty(nnz, dtype=torch.int32, device=dev)
    csc_val = spmm.csr2csc(rp, ci, colptr, rowind, v)
    torch.cuda.synchronize()
    # stable: within a column rows ascend and duplicates keep CSR order -- build the expectation by a stable argsort
    order = np.argsort(colind, kind="stable")
    rows = np.repeat(np.arange(M), np.diff(rowptr))
    assert np.array_equal(colptr.cpu().numpy(), np.concatenate([[0], np.cumsum(np.bincount(colind, minlength=N))]))
    assert np.array_equal(rowind.cpu().numpy(), rows[order])
    assert np.array_equal(csc_val.cpu().numpy(), val[order])
    AT = sp.csr_matrix((csc_val.cpu().numpy(), rowind.cpu().numpy(), colptr.cpu().numpy()), shape=(N, M))
    assert abs(AT - A.T).max() == 0 if nnz else True


def test_autograd_forward_backward(dev, oracle, pkg):
    from gespmm_b200.op import SPMMFunction, spmm
    rng = np.random.default_rng(21)
    M, N, nnz, K = 400, 300, 5000, 48
    rowptr, colind = _rand_csr(rng, M, N, nnz)
    val = rng.standard_normal(nnz).astype(np.float32)
    rp, ci, v = (torch.as_tensor(x, device=dev) for x in (rowptr, colind, val))
    colptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    rowind = torch.empty(nnz, dtype=torch.int32, device=dev)
    v_csc = spmm.csr2csc(rp, ci, colptr, rowind, v)
    rows = torch.repeat_interleave(torch.arange(M, device=dev), (rp[1:] - rp[:-1]).long())
    A = torch.zeros(M, N, device=dev, dtype=torch.float64).index_put_((rows, ci.long()), v.double(), accumulate=True)
    A1 = torch.zeros(M, N, device=dev, dtype=torch.float64).index_put_((rows, ci.long()), torch.ones(nnz, device=dev, dtype=torch.float64), accumulate=True)
    x = torch.randn(N, K, device=dev, requires_grad=True)
    w = torch.randn(M, K, device=dev)
    for ew_csr, ew_csc, Ad in ((None, None, A1), (v, v_csc, A)):
        x.grad = None
        y = SPMMFunction.apply(rp, ci, colptr, rowind, x, ew_csr, ew_csc)
        (y * w).sum().backward()
        assert torch.allclose(y.double(), Ad @ x.detach().double(), rtol=1e-4, atol=1e-4)
        assert torch.allclose(x.grad.double(), Ad.t() @ w.double(), rtol=1e-4, atol=1e-4)
    y = SPMMFunction.apply(rp, ci, colptr, rowind, x, v, None)
    with pytest.raises(RuntimeError, match="both src-first and dst-first"):  # op.py:22-27
        y.sum().backward()


def test_gcnconv_matches_dense(dev, pkg):
    from gespmm_b200 import graphs
    from gespmm_b200.op import GCNConv
    torch.manual_seed(0)
    N = 1000
    srp, sci = graphs.social_like(N, 16000, seed=8)
    # add a ring so that every node has degree > 0: the reference's 1/sqrt(deg) has no zero guard (op.py:104-109)
    rows = torch.repeat_interleave(torch.arange(N), (srp[1:] - srp[:-1]).long())
    ring = torch.arange(N)
    rp, ci = graphs.coo_to_csr(torch.cat([rows, ring, (ring + 1) % N]), torch.cat([sci.long(), (ring + 1) % N, ring]), N, N, dedup=True)
    rp, ci = rp.to(dev), ci.to(dev)
    deg = (rp[1:] - rp[:-1])
    assert (deg > 0).all()
    conv = GCNConv(32, 16).to(dev)
    with torch.no_grad():
        conv.bias.uniform_(-1, 1)
    x = torch.randn(N, 32, device=dev, requires_grad=True)
    out = conv(x, rp, ci, rp, ci)
    rows = torch.repeat_interleave(torch.arange(N, device=dev), deg.long())
    A = torch.zeros(N, N, device=dev).index_put_((rows, ci.long()), torch.ones(ci.numel(), device=dev), accumulate=True)
    dn = deg.float().rsqrt()[:, None]
    want = (A @ ((x @ conv.weight) * dn)) * dn + conv.bias
    assert torch.allclose(out, want, rtol=1e-4, atol=1e-4)
    out.sum().backward()
    assert x.grad is not None and conv.weight.grad is not None and torch.isfinite(x.grad).all()


@pytest.mark.parametrize("per_row,K", [(4, 32), (4, 128), (60, 32), (60, 128), (300, 64), (300, 128), (300, 47)])
def test_fused_function_routes_give_the_unfused_bits(dev, pkg, per_row, K):
    """FusedSPMMFunction scales the gathered rows inside the kernel on sparse graphs and in a pass of its own on dense ones
    (op._fuse_gather_scale); the row scale and the bias are always inside.  Whatever the route, forward and backward carry
    the bits of x * cs -> SPMMFunction -> * rs -> + bias (op.py:142-147)."""
    from gespmm_b200 import graphs
    from gespmm_b200.op import FusedSPMMFunction, SPMMFunction, _fuse_gather_scale
    N = 3000
    rp, ci = graphs.social_like(N, N * per_row, seed=per_row + K, device=dev)
    assert _fuse_gather_scale(ci.numel(), N, K) == (per_row < 16 or (K > 64 and per_row < 128))
    g = torch.Generator(device=dev).manual_seed(K)
    rs = (torch.rand(N, 1, generator=g, device=dev) + 0.5)
    cs = (torch.rand(N, 1, generator=g, device=dev) + 0.5)
    for valued in (False, True):
        ew = (torch.rand(ci.numel(), generator=g, device=dev) + 0.5) if valued else None
        x1 = torch.randn(N, K, generator=g, device=dev).requires_grad_(True)
        x2 = x1.detach().clone().requires_grad_(True)
        b1 = torch.randn(K, generator=g, device=dev).requires_grad_(True)
        b2 = b1.detach().clone().requires_grad_(True)
        y1 = SPMMFunction.apply(rp, ci, rp, ci, x1 * cs, ew, ew) * rs + b1          # symmetric graph: CSC == CSR
        y2 = FusedSPMMFunction.apply(rp, ci, rp, ci, x2, rs, cs, b2, ew, ew)
        assert torch.equal(y1, y2)
        w = torch.randn(N, K, generator=g, device=dev)
        y1.backward(w); y2.backward(w)
        assert torch.equal(x1.grad, x2.grad)
        assert torch.allclose(b1.grad, b2.grad, rtol=1e-5, atol=1e-4)


def test_gcnconv_fused_norm_and_training_loop(dev, pkg, golden_csr, tmp_path):
    """fuse_norm runs both degree normalisations and the bias add inside the kernel (gespmm_opts.row_scale / col_scale /
    bias): the layer's output and its gradients carry the BITS of the unfused layer (op.py:142-147), valued and unvalued;
    the 2-layer training loop (the reference's gcn_custom.py, on the real PubMed adjacency read from a .mtx) learns."""
    import importlib.util
    from gespmm_b200 import graphs
    from gespmm_b200.op import GCNConv
    spec = importlib.util.spec_from_file_location("gcn_custom", os.path.join(ROOT, "ge-spmm_b200", "gcn_custom.py"))
    gcn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gcn)
    rowptr, colind, shape = golden_csr("pubmed")
    mtx = str(tmp_path / "pubmed.mtx")
    graphs.write_mtx(mtx, rowptr, colind, N=shape[1])
    g, x, y, masks, n_in, n_out = gcn.load_problem(dev, mtx=mtx)
    assert g["rowptr"].numel() - 1 == 19717 and g["colind"].numel() == 88648 + 19717  # + self-loops
    for valued in (True, False):
        a = (g["rowptr"], g["colind"], g["colptr"], g["rowind"]) + ((g["value_csr"], g["value_csc"]) if valued else ())
        for width in (32, 128, 3):
            torch.manual_seed(1)
            plain = GCNConv(n_in, width, cached=True).to(dev)
            fused = GCNConv(n_in, width, cached=True, fuse_norm=True).to(dev)
            with torch.no_grad():
                plain.bias.uniform_(-1, 1)
            fused.load_state_dict(plain.state_dict())
            xg1, xg2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
            o1, o2 = plain(xg1, *a), fused(xg2, *a)
            assert torch.equal(o1, o2), "fused layer output must be bit-identical (width %d, valued %s)" % (width, valued)
            w = torch.randn_like(o1)
            (o1 * w).sum().backward(); (o2 * w).sum().backward()
            assert torch.equal(xg1.grad, xg2.grad) and torch.equal(plain.weight.grad, fused.weight.grad)
            assert torch.allclose(plain.bias.grad, fused.bias.grad, rtol=1e-5, atol=1e-5)
    for fuse in (False, True):
        res = gcn.run(n_hidden=16, layers=2, epochs=40, fuse_norm=fuse, log=lambda *_: None, mtx=mtx)
        assert res["last_loss"] < 0.7 * res["first_loss"] and res["best_val"] > 0.5, res
    res = gcn.run(n_hidden=16, layers=2, epochs=5, fuse_norm=True, log=lambda *_: None)  # the synthetic default graph
    assert res["last_loss"] == res["last_loss"]


# ---- CLI ---------------------------------------------------------------------------------------------

def test_cli_contract(tmp_path, oracle, pkg, golden_csr):
    """./spmm_test <mtx> [dev]: stdout lines and CSV cell order of the reference (spmm_test.cu:536,583,635,722,738,762)."""
    from gespmm_b200 import build, graphs
    rowptr, colind, shape = golden_csr("pubmed")
    mtx = str(tmp_path / "pubmed.mtx")
    graphs.write_mtx(mtx, rowptr, colind)
    cmd = [build.CLI, mtx, "0", "--iters", "20", "--validate", "--json"]
    if oracle.have_ref(oracle.REF_CLI_KERNELS):
        cmd += ["--baseline-lib", oracle.REF_CLI_KERNELS]
    res = subprocess.run(cmd, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    out = res.stdout
    assert "reading file ..." in out and "read file ok. N=19717 nnz=88648" in out
    assert "max_ncols = 512" in out and "running tests..." in out
    assert "mismatches = 0" in out
    if oracle.have_ref(oracle.REF_CLI_KERNELS):
        assert "differing bitwise from the reference kernel = 0" in out
    cells = open(tmp_path / "spmm_test_out.out").read().strip(",").split(",")
    assert len(cells) == 6  # (baseline, ours) for K = 128, 256, 512
    vals = [float(c) for c in cells]
    assert all(v > 0 for v in vals[1::2])
    res = subprocess.run([build.CLI, str(tmp_path / "nope.mtx")], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert res.returncode != 0 and "not found" in res.stdout
