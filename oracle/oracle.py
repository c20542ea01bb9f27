"""Python face of the oracle: ctypes loader for oracle/liboracle.so (the C restatement in
spmm_oracle.c), a numpy restatement for small cases, and loaders for the reference's own
code built under oracle/_ref/ (see oracle/Makefile).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under ge-spmm_b200/ may import this module.

Parity pin: checked against the reference reader compiled from /root/reference
(oracle/_ref/libref_readmtx.so) on the three bundled matrices and on hand-made edge cases
(tests/golden/, tests/test_oracle.py), and on the GPU box against the reference kernels
compiled from /root/reference (oracle/_ref/ref_spmm*.so, libref_cli_kernels.so) bit for bit
(tests/test_spmm_gpu.py).
"""
import ctypes
import importlib.util
import os
import subprocess
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_READMTX = os.path.join(REF_DIR, "libref_readmtx.so")
REF_CLI_KERNELS = os.path.join(REF_DIR, "libref_cli_kernels.so")
REF_CLI_BIN = os.path.join(REF_DIR, "spmm_test_ref")
REF_EXT = os.path.join(REF_DIR, "ref_spmm" + sysconfig.get_config_var("EXT_SUFFIX"))

_lib = None
_i64, _p = ctypes.c_int64, ctypes.c_void_p


def build(ref=True, verbose=False):
    """make -C oracle [ref]; the ref target only when /root/reference is present."""
    targets = ["all"]
    if ref and os.path.isdir("/root/reference"):
        targets.append("ref")
    res = subprocess.run(["make", "-C", HERE] + targets, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building the oracle failed")


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build(ref=False)
        L = ctypes.CDLL(LIB)
        L.oracle_spmm_valued_f32.argtypes = [_i64, _i64, _p, _p, _p, _p, _i64, _p, _i64, ctypes.c_int, ctypes.c_int]
        L.oracle_spmm_unvalued_f32.argtypes = [_i64, _i64, _p, _p, _p, _i64, _p, _i64, ctypes.c_int]
        L.oracle_spmm_max_f32.argtypes = [_i64, _i64, _p, _p, _p, _p, _i64, _p, _i64, ctypes.c_float, ctypes.c_int]
        L.oracle_spmm_f64.argtypes = [_i64, _i64, _p, _p, _p, _p, _i64, _p, _p, ctypes.c_int]
        L.oracle_ref_dispatch.argtypes = [_i64, _i64, _p, _p, _p]
        L.oracle_ref_dispatch.restype = ctypes.c_int
        L.oracle_coo_to_csr.argtypes = [_i64, _i64, _p, _p, _p, _p, _p]
        L.oracle_fill_B_cli.argtypes = [_p, _i64, ctypes.c_uint]
        L.oracle_read_mtx.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                      ctypes.POINTER(ctypes.c_int64), _p, _p, _p, _i64]
        L.oracle_read_mtx.restype = ctypes.c_int
        L.oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def num_threads():
    return lib().oracle_num_threads()


def spmm(rowptr, colind, val, B, fma=True, nthreads=0):
    """fp32 C = A @ B in the reference's per-element order (C restatement).  val=None -> unvalued."""
    rowptr, colind, B = _c(rowptr, np.int32), _c(colind, np.int32), _c(B, np.float32)
    M, K = rowptr.shape[0] - 1, B.shape[1]
    C = np.empty((M, K), np.float32)
    if val is None:
        lib().oracle_spmm_unvalued_f32(M, K, rowptr.ctypes.data, colind.ctypes.data, B.ctypes.data, K, C.ctypes.data, K, nthreads)
    else:
        val = _c(val, np.float32)
        lib().oracle_spmm_valued_f32(M, K, rowptr.ctypes.data, colind.ctypes.data, val.ctypes.data, B.ctypes.data, K,
                                     C.ctypes.data, K, 1 if fma else 0, nthreads)
    return C


def spmm_max(rowptr, colind, val, B, init=-10000.0, nthreads=0):
    """Max-reduce restatement (dgl-custom/binary_reduce_max.cu); init = -10000 is the reference's max_init()."""
    rowptr, colind, B = _c(rowptr, np.int32), _c(colind, np.int32), _c(B, np.float32)
    M, K = rowptr.shape[0] - 1, B.shape[1]
    C = np.empty((M, K), np.float32)
    v = None if val is None else _c(val, np.float32)
    lib().oracle_spmm_max_f32(M, K, rowptr.ctypes.data, colind.ctypes.data, None if v is None else v.ctypes.data,
                              B.ctypes.data, K, C.ctypes.data, K, float(init), nthreads)
    return C


def spmm_f64(rowptr, colind, val, B, nthreads=0, with_rowabs=True):
    """fp64 golden and sum|a||b| per element (the error scale for tolerance tests)."""
    rowptr, colind, B = _c(rowptr, np.int32), _c(colind, np.int32), _c(B, np.float32)
    M, K = rowptr.shape[0] - 1, B.shape[1]
    C = np.empty((M, K), np.float64)
    A = np.empty((M, K), np.float64) if with_rowabs else None
    v = None if val is None else _c(val, np.float32)
    lib().oracle_spmm_f64(M, K, rowptr.ctypes.data, colind.ctypes.data, None if v is None else v.ctypes.data,
                          B.ctypes.data, K, C.ctypes.data, None if A is None else A.ctypes.data, nthreads)
    return C, A


def spmm_numpy(rowptr, colind, val, B, fma=False):
    """Pure numpy/python restatement of the same loop (spmm_test.cu:595-605); small cases only.
    fp32 throughout; fma=False rounds the product and the sum separately (host golden)."""
    rowptr, colind, B = np.asarray(rowptr), np.asarray(colind), np.asarray(B, dtype=np.float32)
    M, K = len(rowptr) - 1, B.shape[1]
    C = np.zeros((M, K), np.float32)
    for i in range(M):
        acc = np.zeros(K, np.float32)
        for p in range(rowptr[i], rowptr[i + 1]):
            b = B[colind[p]]
            if val is None:
                acc = (acc + b).astype(np.float32)
            elif fma:
                acc = (np.float64(val[p]) * b.astype(np.float64) + acc.astype(np.float64)).astype(np.float32)
            else:
                acc = (acc + (np.float32(val[p]) * b).astype(np.float32)).astype(np.float32)
        C[i] = acc
    return C


def ref_dispatch(m, k):
    """(kernel id, grid, block, smem) the reference launcher picks for (m, k) (spmm_kernel.cu:186-205)."""
    g, b, s = (ctypes.c_int64 * 2)(), (ctypes.c_int64 * 2)(), ctypes.c_int64()
    kid = lib().oracle_ref_dispatch(m, k, g, b, ctypes.byref(s))
    return kid, tuple(g), tuple(b), s.value


def coo_to_csr(nrows, row, col):
    row, col = _c(row, np.int32), _c(col, np.int32)
    nnz = row.shape[0]
    indptr, indices, data = np.empty(nrows + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.float32)
    lib().oracle_coo_to_csr(nrows, nnz, row.ctypes.data, col.ctypes.data, indptr.ctypes.data, indices.ctypes.data, data.ctypes.data)
    return indptr, indices, data


def fill_B_cli(n, seed):
    B = np.empty(n, np.float32)
    lib().oracle_fill_B_cli(B.ctypes.data, n, seed)
    return B


def _read_with(fn, path):
    nr, nc, nv = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64()
    rc = fn(os.fsencode(path), ctypes.byref(nr), ctypes.byref(nc), ctypes.byref(nv), None, None, None, 0)
    if rc != 0:
        raise IOError("mtx read failed with code %d: %s" % (rc, path))
    n = nv.value
    r, c, v = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float32)
    rc = fn(os.fsencode(path), ctypes.byref(nr), ctypes.byref(nc), ctypes.byref(nv), r.ctypes.data, c.ctypes.data, v.ctypes.data, n)
    if rc != 0:
        raise IOError("mtx read failed with code %d: %s" % (rc, path))
    return nr.value, nc.value, r, c, v


def read_mtx(path):
    """(nrows, ncols, row, col, val) COO with readMtx's post-conditions -- the C restatement."""
    return _read_with(lib().oracle_read_mtx, path)


# ---- the reference's own code, compiled from /root/reference into oracle/_ref ---------------------

def have_ref(path):
    return os.path.exists(path)


def ref_read_mtx(path, isolated=True):
    """The reference's readMtx<float> itself (util/util.hpp:286-333).

    Its compaction scan reads one element past its vectors whenever the sorted entry list ends in a removed entry
    (util.hpp:268-283, `back<=nvals`; the bundled cora.mtx does: its last entry is a self-loop) and then drops one real
    entry if the stale int it finds there happens to be -1 -- so what it returns depends on the heap it runs in, and
    inside a long-lived test process it very occasionally returns 10,555 entries for cora instead of 10,556.
    isolated=True (default) therefore runs it in a fresh interpreter, whose heap is the same every time."""
    if isolated:
        import io
        import subprocess
        import sys
        code = ("import sys, numpy as np; sys.path.insert(0, %r); from oracle import oracle; "
                "nr, nc, r, c, v = oracle.ref_read_mtx(%r, isolated=False); "
                "np.savez(sys.stdout.buffer, shape=np.array([nr, nc]), r=r, c=c, v=v)" % (os.path.dirname(HERE), os.fspath(path)))
        res = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if res.returncode != 0:
            raise IOError("reference reader failed on %s: %s" % (path, res.stderr.decode()[-300:]))
        z = np.load(io.BytesIO(res.stdout))
        return int(z["shape"][0]), int(z["shape"][1]), z["r"], z["c"], z["v"]
    L = ctypes.CDLL(REF_READMTX)
    L.ref_read_mtx.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                               ctypes.POINTER(ctypes.c_int64), _p, _p, _p, _i64]
    L.ref_read_mtx.restype = ctypes.c_int
    return _read_with(L.ref_read_mtx, path)


def ref_cli_kernels():
    """ctypes handle on spmmWrapper / spmm_test0..4<float> of the reference CLI (GPU only)."""
    L = ctypes.CDLL(REF_CLI_KERNELS)
    L.ref_spmm_wrapper.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _p, _p, _p, _p, _p]
    L.ref_spmm_wrapper.restype = ctypes.c_int
    L.ref_spmm_time_ms.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _p, _p, _p, _p, _p,
                                   ctypes.c_int, ctypes.c_int]
    L.ref_spmm_time_ms.restype = ctypes.c_float
    return L


def ref_extension():
    """The reference PyTorch extension (pytorch-custom/spmm.cpp + spmm_kernel.cu) as module ref_spmm."""
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("ref_spmm", REF_EXT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
