"""ctypes binding of the C ABI in include/gespmm.h (lib/libgespmm.so).

This is the binding a non-PyTorch host would write; it carries no torch types.  Device
pointers are plain integers (e.g. ``tensor.data_ptr()``), streams are ``cudaStream_t`` as
integers (0 / None = legacy default stream, which the reference launches on:
pytorch-custom/spmm_kernel.cu:189,196,203).
"""
import ctypes
import os

import numpy as np

from . import build as _build

OK = 0
ERR_INVALID_ARG, ERR_CUDA, ERR_TOO_LARGE, ERR_IO, ERR_WORKSPACE, ERR_NOMEM = -1, -2, -3, -4, -5, -6
LONG_ROW = 4096

# every symbol include/gespmm.h declares
SYMBOLS = (
    "gespmm_version", "gespmm_error_string", "gespmm_csr_spmm_f32", "gespmm_csr_spmm_f32_ex", "gespmm_opts_init", "gespmm_max_row_nnz",
    "gespmm_reload_env", "gespmm_thread_cleanup", "gespmm_pad_workspace_bytes", "gespmm_row_sum_is_sequential_ex", "gespmm_csr_spmm_f32_host", "gespmm_csr_spmm_f32_bparts", "gespmm_csr_spmm_max_f32", "gespmm_row_sum_is_sequential", "gespmm_enable_peer_access", "gespmm_ipc_open", "gespmm_ipc_close", "gespmm_ipc_alloc", "gespmm_ipc_free",
    "gespmm_csr2csc_workspace_bytes", "gespmm_csr2csc_f32", "gespmm_read_mtx", "gespmm_free_host",
    "gespmm_write_csr", "gespmm_read_csr", "gespmm_read_mtx_cached", "gespmm_write_mtx",
)

_lib = None

FLAG_SEQUENTIAL, FLAG_NO_OVERLAP = 0x1, 0x2
WALKER_AUTO, WALKER_RING, WALKER_REGISTER, WALKER_SUBWARP, WALKER_ROWS, WALKER_BULK = 0, 1, 2, 3, 4, 5


class Opts(ctypes.Structure):
    """gespmm_opts (include/gespmm.h); make one with opts(...)."""
    _fields_ = [("struct_size", ctypes.c_uint32), ("flags", ctypes.c_uint32), ("max_row_nnz", ctypes.c_int64),
                ("row_scale", ctypes.c_void_p), ("col_scale", ctypes.c_void_p), ("bias", ctypes.c_void_p),
                ("walker", ctypes.c_int32), ("task_keys", ctypes.c_int32), ("long_row", ctypes.c_int32),
                ("panel_v", ctypes.c_int32), ("l2_policy", ctypes.c_int32), ("l2_window_rows", ctypes.c_int32),
                ("hot_columns", ctypes.c_void_p), ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t)]


class GespmmError(RuntimeError):
    def __init__(self, code, what):
        self.code = code
        super().__init__("%s failed: %s (code %d)" % (what, error_string(code), code))


def lib():
    """Load lib/libgespmm.so; there is no fallback if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(_build.LIB):
            raise ImportError(
                "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(this package has no CPU or pure-PyTorch path)" % _build.LIB)
        L = ctypes.CDLL(_build.LIB)
        i64, p, sz = ctypes.c_int64, ctypes.c_void_p, ctypes.c_size_t
        L.gespmm_version.restype = ctypes.c_int
        L.gespmm_error_string.restype = ctypes.c_char_p
        L.gespmm_error_string.argtypes = [ctypes.c_int]
        L.gespmm_csr_spmm_f32.restype = ctypes.c_int
        L.gespmm_csr_spmm_f32.argtypes = [i64, i64, i64, i64, p, p, p, p, i64, p, i64, p]
        L.gespmm_csr_spmm_f32_ex.restype = ctypes.c_int
        L.gespmm_csr_spmm_f32_ex.argtypes = [i64, i64, i64, i64, p, p, p, p, i64, p, i64, ctypes.POINTER(Opts), p]
        L.gespmm_opts_init.restype = None
        L.gespmm_opts_init.argtypes = [ctypes.POINTER(Opts)]
        L.gespmm_max_row_nnz.restype = ctypes.c_int
        L.gespmm_max_row_nnz.argtypes = [i64, p, ctypes.POINTER(ctypes.c_int32), p]
        L.gespmm_reload_env.restype = None
        L.gespmm_reload_env.argtypes = []
        L.gespmm_pad_workspace_bytes.restype = sz
        L.gespmm_pad_workspace_bytes.argtypes = [i64, i64, i64, i64]
        L.gespmm_thread_cleanup.restype = None
        L.gespmm_thread_cleanup.argtypes = []
        L.gespmm_row_sum_is_sequential_ex.restype = ctypes.c_int
        L.gespmm_row_sum_is_sequential_ex.argtypes = [i64, i64, ctypes.POINTER(Opts)]
        L.gespmm_csr_spmm_f32_bparts.restype = ctypes.c_int
        L.gespmm_csr_spmm_f32_bparts.argtypes = [i64, i64, i64, i64, p, p, p, ctypes.c_int, ctypes.POINTER(p),
                                                 ctypes.POINTER(i64), i64, p, i64, p]
        L.gespmm_csr_spmm_max_f32.restype = ctypes.c_int
        L.gespmm_csr_spmm_max_f32.argtypes = [i64, i64, i64, i64, p, p, p, p, i64, p, i64, ctypes.c_float, p]
        L.gespmm_row_sum_is_sequential.restype = ctypes.c_int
        L.gespmm_row_sum_is_sequential.argtypes = [i64, i64]
        L.gespmm_enable_peer_access.restype = ctypes.c_int
        L.gespmm_enable_peer_access.argtypes = [ctypes.c_int]
        L.gespmm_ipc_open.restype = ctypes.c_int
        L.gespmm_ipc_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(p)]
        L.gespmm_ipc_close.restype = ctypes.c_int
        L.gespmm_ipc_close.argtypes = [p]
        L.gespmm_ipc_alloc.restype = ctypes.c_int
        L.gespmm_ipc_alloc.argtypes = [sz, ctypes.POINTER(p), ctypes.c_char_p]
        L.gespmm_ipc_free.restype = ctypes.c_int
        L.gespmm_ipc_free.argtypes = [p]
        L.gespmm_csr_spmm_f32_host.restype = ctypes.c_int
        L.gespmm_csr_spmm_f32_host.argtypes = [i64, i64, i64, i64, p, p, p, p, i64, p, i64, ctypes.c_int]
        L.gespmm_csr2csc_workspace_bytes.restype = sz
        L.gespmm_csr2csc_workspace_bytes.argtypes = [i64, i64, i64]
        L.gespmm_csr2csc_f32.restype = ctypes.c_int
        L.gespmm_csr2csc_f32.argtypes = [i64, i64, i64, p, p, p, p, p, p, p, sz, p]
        L.gespmm_read_mtx.restype = ctypes.c_int
        L.gespmm_read_mtx.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                      ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(p), ctypes.POINTER(p),
                                      ctypes.POINTER(p)]
        outs = [ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64),
                ctypes.POINTER(p), ctypes.POINTER(p), ctypes.POINTER(p)]
        L.gespmm_write_csr.restype = ctypes.c_int
        L.gespmm_write_csr.argtypes = [ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32, i64, p, p, p]
        L.gespmm_write_mtx.restype = ctypes.c_int
        L.gespmm_write_mtx.argtypes = [ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32, i64, p, p, p]
        L.gespmm_read_csr.restype = ctypes.c_int
        L.gespmm_read_csr.argtypes = [ctypes.c_char_p] + outs
        L.gespmm_read_mtx_cached.restype = ctypes.c_int
        L.gespmm_read_mtx_cached.argtypes = [ctypes.c_char_p, ctypes.c_char_p] + outs + [ctypes.POINTER(ctypes.c_int)]
        L.gespmm_free_host.restype = None
        L.gespmm_free_host.argtypes = [p]
        _lib = L
    return _lib


def version():
    return lib().gespmm_version()


def error_string(code):
    return lib().gespmm_error_string(int(code)).decode()


def csr_spmm_f32(M, N, K, nnz, rowptr, colind, val, B, ldb, C, ldc, stream=None):
    """Raw call with device pointers (ints).  ``val`` may be None/0 (all-ones A)."""
    rc = lib().gespmm_csr_spmm_f32(M, N, K, nnz, rowptr, colind, val or None, B, ldb, C, ldc, stream or None)
    if rc != OK:
        raise GespmmError(rc, "gespmm_csr_spmm_f32")


def opts(sequential=False, no_overlap=False, max_row_nnz=-1, row_scale=None, col_scale=None, bias=None, walker=0,
         task_keys=0, long_row=0, panel_v=0, l2_policy=0, l2_window_rows=0, hot_columns=None, workspace=None, workspace_bytes=0):
    """A gespmm_opts initialised by the library and filled in; the scale / bias arguments are device pointers (ints)."""
    o = Opts()
    lib().gespmm_opts_init(ctypes.byref(o))
    o.flags = (FLAG_SEQUENTIAL if sequential else 0) | (FLAG_NO_OVERLAP if no_overlap else 0)
    o.max_row_nnz = int(max_row_nnz)
    o.row_scale, o.col_scale, o.bias = row_scale or None, col_scale or None, bias or None
    o.walker, o.task_keys, o.long_row, o.panel_v = int(walker), int(task_keys), int(long_row), int(panel_v)
    o.l2_policy, o.l2_window_rows = int(l2_policy), int(l2_window_rows)
    o.hot_columns = hot_columns or None
    o.workspace, o.workspace_bytes = workspace or None, int(workspace_bytes)
    return o


def csr_spmm_f32_ex(M, N, K, nnz, rowptr, colind, val, B, ldb, C, ldc, options=None, stream=None):
    """gespmm_csr_spmm_f32 with per-call options (an Opts from opts(), or None)."""
    rc = lib().gespmm_csr_spmm_f32_ex(M, N, K, nnz, rowptr, colind, val or None, B, ldb, C, ldc,
                                      None if options is None else ctypes.byref(options), stream or None)
    if rc != OK:
        raise GespmmError(rc, "gespmm_csr_spmm_f32_ex")


def max_row_nnz(M, rowptr, stream=None):
    """Longest row of a device CSR (synchronises the stream)."""
    out = ctypes.c_int32(0)
    rc = lib().gespmm_max_row_nnz(int(M), rowptr, ctypes.byref(out), stream or None)
    if rc != OK:
        raise GespmmError(rc, "gespmm_max_row_nnz")
    return int(out.value)


def pad_workspace_bytes(M, N, K, nnz):
    """Bytes of Opts.workspace that let an odd width K > 16 run on the 16-byte-slice walkers (0: the library would not pad)."""
    return int(lib().gespmm_pad_workspace_bytes(int(M), int(N), int(K), int(nnz)))


def thread_cleanup():
    """Release the calling thread's helper streams / events (they are re-created on demand)."""
    lib().gespmm_thread_cleanup()


def reload_env():
    """Re-read the GESPMM_* tuning environment (it is read once, at the first call into the library)."""
    lib().gespmm_reload_env()


def csr_spmm_max_f32(M, N, K, nnz, rowptr, colind, val, B, ldb, C, ldc, init=-10000.0, stream=None):
    """Max-reduce variant (dgl-custom/binary_reduce_max.cu); init = -10000 is the reference's max_init()."""
    rc = lib().gespmm_csr_spmm_max_f32(M, N, K, nnz, rowptr, colind, val or None, B, ldb, C, ldc, float(init), stream or None)
    if rc != OK:
        raise GespmmError(rc, "gespmm_csr_spmm_max_f32")


def row_sum_is_sequential(K, row_nnz, options=None):
    """True if a row of ``row_nnz`` nonzeros at width K is summed in the reference's sequential order (bit-identical)."""
    if options is not None:
        return bool(lib().gespmm_row_sum_is_sequential_ex(int(K), int(row_nnz), ctypes.byref(options)))
    return bool(lib().gespmm_row_sum_is_sequential(int(K), int(row_nnz)))


def enable_peer_access(peer_device):
    rc = lib().gespmm_enable_peer_access(int(peer_device))
    if rc != OK:
        raise GespmmError(rc, "gespmm_enable_peer_access(%d)" % peer_device)


def ipc_open(handle):
    """64-byte CUDA IPC memory handle -> base address (int) of the mapping on the current device."""
    base = ctypes.c_void_p()
    rc = lib().gespmm_ipc_open(bytes(handle), ctypes.byref(base))
    if rc != OK:
        raise GespmmError(rc, "gespmm_ipc_open")
    return int(base.value)


def ipc_alloc(nbytes):
    """cudaMalloc on the current device -> (device address, 64-byte IPC handle)."""
    ptr = ctypes.c_void_p()
    handle = ctypes.create_string_buffer(64)
    rc = lib().gespmm_ipc_alloc(int(nbytes), ctypes.byref(ptr), handle)
    if rc != OK:
        raise GespmmError(rc, "gespmm_ipc_alloc")
    return int(ptr.value), handle.raw


def ipc_free(ptr):
    lib().gespmm_ipc_free(ctypes.c_void_p(int(ptr)))


def ipc_close(base):
    lib().gespmm_ipc_close(ctypes.c_void_p(int(base)))


def csr_spmm_f32_bparts(M, N, K, nnz, rowptr, colind, val, part_ptrs, part_begin, ldb, C, ldc, stream=None):
    """B as row blocks: ``part_ptrs`` device pointers (ints), ``part_begin`` row boundaries (len parts + 1)."""
    parts = len(part_ptrs)
    ptrs = (ctypes.c_void_p * parts)(*[ctypes.c_void_p(int(x)) for x in part_ptrs])
    begins = (ctypes.c_int64 * (parts + 1))(*[int(x) for x in part_begin])
    rc = lib().gespmm_csr_spmm_f32_bparts(M, N, K, nnz, rowptr, colind, val or None, parts, ptrs, begins, ldb, C, ldc,
                                          stream or None)
    if rc != OK:
        raise GespmmError(rc, "gespmm_csr_spmm_f32_bparts")


def csr_spmm_host(rowptr, colind, val, B, device=0):
    """Host numpy arrays in, host numpy C out, through gespmm_csr_spmm_f32_host."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colind = np.ascontiguousarray(colind, dtype=np.int32)
    B = np.ascontiguousarray(B, dtype=np.float32)
    if val is not None:
        val = np.ascontiguousarray(val, dtype=np.float32)
    M, (N, K), nnz = rowptr.shape[0] - 1, B.shape, colind.shape[0]
    C = np.empty((M, K), dtype=np.float32)
    rc = lib().gespmm_csr_spmm_f32_host(M, N, K, nnz, rowptr.ctypes.data, colind.ctypes.data,
                                        val.ctypes.data if val is not None else None, B.ctypes.data, K,
                                        C.ctypes.data, K, device)
    if rc != OK:
        raise GespmmError(rc, "gespmm_csr_spmm_f32_host")
    return C


def _take_csr(call, what):
    """Run a reader entry point and copy its malloc'ed arrays into numpy arrays."""
    L = lib()
    nr, nc, nnz = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64()
    rp, ci, vv = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    rc = call(ctypes.byref(nr), ctypes.byref(nc), ctypes.byref(nnz), ctypes.byref(rp), ctypes.byref(ci), ctypes.byref(vv))
    if rc != OK:
        raise GespmmError(rc, what)
    try:
        n = nnz.value
        rowptr = np.ctypeslib.as_array(ctypes.cast(rp, ctypes.POINTER(ctypes.c_int32)), (nr.value + 1,)).copy()
        if n:
            colind = np.ctypeslib.as_array(ctypes.cast(ci, ctypes.POINTER(ctypes.c_int32)), (n,)).copy()
            val = np.ctypeslib.as_array(ctypes.cast(vv, ctypes.POINTER(ctypes.c_float)), (n,)).copy()
        else:
            colind, val = np.empty(0, np.int32), np.empty(0, np.float32)
    finally:
        L.gespmm_free_host(rp); L.gespmm_free_host(ci); L.gespmm_free_host(vv)
    return nr.value, nc.value, rowptr, colind, val


def read_mtx(path, cache=None):
    """MatrixMarket file -> (nrows, ncols, rowptr, colind, val) as numpy arrays (readMtx post-conditions).
    ``cache``: None = always parse; True = keep / use the binary image "<path>.gespmm-csr"; a path = that image."""
    L = lib()
    if cache is None or cache is False:
        return _take_csr(lambda *o: L.gespmm_read_mtx(os.fsencode(path), *o), "gespmm_read_mtx(%s)" % path)
    image = None if cache is True else os.fsencode(cache)
    return _take_csr(lambda *o: L.gespmm_read_mtx_cached(os.fsencode(path), image, *o, None), "gespmm_read_mtx_cached(%s)" % path)


def read_mtx_cached(path, cache_path=None):
    """Like read_mtx(path, cache=...) but also reports whether the image was used: (..., cache_hit)."""
    L = lib()
    hit = ctypes.c_int(0)
    out = _take_csr(lambda *o: L.gespmm_read_mtx_cached(os.fsencode(path), None if cache_path is None else os.fsencode(cache_path),
                                                       *o, ctypes.byref(hit)), "gespmm_read_mtx_cached(%s)" % path)
    return out + (bool(hit.value),)


def write_csr(path, nrows, ncols, rowptr, colind, val):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colind = np.ascontiguousarray(colind, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float32)
    rc = lib().gespmm_write_csr(os.fsencode(path), int(nrows), int(ncols), int(colind.shape[0]), rowptr.ctypes.data,
                                colind.ctypes.data if colind.shape[0] else None, val.ctypes.data if val.shape[0] else None)
    if rc != OK:
        raise GespmmError(rc, "gespmm_write_csr(%s)" % path)


def write_mtx(path, nrows, ncols, rowptr, colind, val=None):
    """CSR -> `general` MatrixMarket coordinate file (pattern if val is None, else real)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colind = np.ascontiguousarray(colind, dtype=np.int32)
    if val is not None:
        val = np.ascontiguousarray(val, dtype=np.float32)
    rc = lib().gespmm_write_mtx(os.fsencode(path), int(nrows), int(ncols), int(colind.shape[0]), rowptr.ctypes.data,
                                colind.ctypes.data if colind.shape[0] else None, None if val is None else val.ctypes.data)
    if rc != OK:
        raise GespmmError(rc, "gespmm_write_mtx(%s)" % path)


def read_csr(path):
    L = lib()
    return _take_csr(lambda *o: L.gespmm_read_csr(os.fsencode(path), *o), "gespmm_read_csr(%s)" % path)
