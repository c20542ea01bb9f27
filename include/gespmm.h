/*
 * gespmm.h -- C ABI of the B200-native CSR x dense SpMM (libgespmm.so).
 *
 * This is the drop-in boundary for the one hot path of hgyhungry/ge-spmm:
 *     C[M,K] = A_csr[M,N] * B[N,K]      fp32 data, int32 CSR indices, row-major dense.
 * Every reference call site of that path passes the same raw-pointer tuple
 * (m, k, rowptr, colind, [val], B, C); the entry points below are what an FFI binding
 * for it would bind.  Plain pointers and sizes only -- no torch, no C++ types.
 *
 *   entry point                    replaces (reference file:line)
 *   -----------------------------  ---------------------------------------------------
 *   gespmm_csr_spmm_f32            spmm_cuda / spmm_cuda_no_edge_value launch blocks
 *                                  (pytorch-custom/spmm_kernel.cu:425-458, 175-207) and the
 *                                  kernels they pick (topoSimple/topoCache/topoCacheCoarsen
 *                                  SPMMKernel :31-173, spmm_test0/1/2 :210-379);
 *                                  spmmWrapper (spmm_test.cu:456-492) and spmm_test0..4<T>
 *                                  (spmm_test.cu:64-454);  XTopoCsrmm<float>
 *                                  (dgl-custom/binary_reduce_sum.cu:309-335) has the same shape.
 *   gespmm_csr_spmm_f32_ex         the same launch blocks plus the element-wise passes GCNConv.forward runs
 *                                  around them (pytorch-custom/op.py:142-147: x * out_deg_norm, * in_deg_norm, + bias),
 *                                  fused; per-call summation order instead of a process-wide switch
 *   gespmm_max_row_nnz             no reference counterpart (per-graph figure that lets a call skip the long-row kernel)
 *   gespmm_csr_spmm_max_f32        topo*SPMMMaxKernel / XTopoCsrmmmax<float>
 *                                  (dgl-custom/binary_reduce_max.cu:26-204)
 *   gespmm_csr2csc_f32             csr2cscKernel / csr2csc_cuda
 *                                  (pytorch-custom/spmm_kernel.cu:381-423, 460-477)
 *   gespmm_read_mtx / _free        readMtx<float> + COO->CSR of the CLI
 *                                  (util/util.hpp:286-333, spmm_test.cu:557-581)
 *   gespmm_read_mtx_cached,        no reference counterpart: the parsed CSR kept as a binary image next
 *   gespmm_write_csr / _read_csr   to the .mtx (the reference re-parses on every run, run_test.sh:5-10)
 *   gespmm_write_mtx               the .mtx rewriting loop of data/conv.c:149-158 (fprintf per entry)
 *   gespmm_row_sum_is_sequential   no reference counterpart (every reference kernel sums sequentially,
 *                                  spmm_kernel.cu:56-59, 165-168): tells which rows this library does
 *   gespmm_csr_spmm_f32_host       the CLI's cudaMalloc / cudaMemcpy block around the launch
 *                                  (spmm_test.cu:609-640)
 *   gespmm_csr_spmm_f32_bparts,    no reference counterpart (the reference is single-GPU): B left
 *   gespmm_ipc_*, _enable_peer_*   row-sharded across GPUs and gathered over NVLink by the kernel
 *
 * Conventions
 *   - All device pointers must live on the device that is current when the call is made.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, which is what
 *     the reference launches on: spmm_kernel.cu:189,196,203).  Calls are asynchronous on it.
 *   - gespmm_csr_spmm_f32 allocates no memory and takes no workspace; its only state is one helper
 *     stream per (host thread, device), created on first use, on which the long-row kernel overlaps
 *     the main kernel (forked from / joined into `stream` with events).  Re-entrant, graph-capturable.
 *   - Return value: GESPMM_OK or a negative GESPMM_ERR_* code.  Never exits the process
 *     (the reference's checkCudaError macros call exit(): spmm_kernel.cu:5-19).
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     GESPMM_ERR_CUDA.
 */
#ifndef GESPMM_H
#define GESPMM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GESPMM_OK                0
#define GESPMM_ERR_INVALID_ARG  -1   /* null pointer with non-zero size, negative size, ld < K   */
#define GESPMM_ERR_CUDA         -2   /* a CUDA runtime call or the launch failed                  */
#define GESPMM_ERR_TOO_LARGE    -3   /* M, N or nnz does not fit the int32 index contract         */
#define GESPMM_ERR_IO           -4   /* file missing / not a MatrixMarket coordinate file        */
#define GESPMM_ERR_WORKSPACE    -5   /* workspace too small                                       */
#define GESPMM_ERR_NOMEM        -6   /* host allocation failed                                    */

/* Library version, major*10000 + minor*100 + patch. */
int gespmm_version(void);

/* Static, human-readable text for a GESPMM_* code. */
const char *gespmm_error_string(int code);

/*
 * C[M,K] = A[M,N] * B[N,K].
 *   rowptr[M+1], colind[nnz]  int32, device.   rowptr[0] == 0, rowptr[M] == nnz.
 *   val[nnz]                  fp32, device, or NULL: A is then all-ones and the row result is the
 *                             plain sum of the gathered B rows (the "no_edge_value" kernels).
 *   B                         fp32, device, row-major, row stride ldb >= K (reference: ldb == K).
 *   C                         fp32, device, row-major, row stride ldc >= K.  Every element of
 *                             C[0:M, 0:K] is written (empty rows -> 0), like the reference.
 * Per output element the products are accumulated in CSR order into one fp32 accumulator
 * starting from 0 (FFMA for valued, FADD for unvalued) -- the reference kernels' order, hence
 * bit-identical results -- with two exceptions, both deterministic and differing from the reference
 * by fp32 re-association only (see gespmm_row_sum_is_sequential below):
 *   - rows longer than GESPMM_LONG_ROW nonzeros are summed in 8 contiguous segments (one per warp
 *     of a CTA; 64 segments across a thread-block cluster from 32768 nonzeros) combined in fixed order;
 *   - for K <= 64 (K % 4 == 0, aligned operands) and for any K <= 16 a warp gathers 2 / 4 / 8 B rows per
 *     instruction and keeps one partial sum per lane group (nonzeros p, p + NG, p + 2 NG, ... of the row), added
 *     in a fixed butterfly order at the row end.  GESPMM_FLAG_SEQUENTIAL in gespmm_opts (per call) or
 *     GESPMM_SEQUENTIAL=1 in the environment (process default) selects, for every K, the fastest walker
 *     that keeps the sequential order (lane groups own disjoint rows; 0.5-1.07x the default's speed on B200).
 * N is used for argument checking only (colind values are trusted, like the reference).
 */
int gespmm_csr_spmm_f32(int64_t M, int64_t N, int64_t K, int64_t nnz,
                        const int32_t *rowptr, const int32_t *colind, const float *val,
                        const float *B, int64_t ldb, float *C, int64_t ldc, void *stream);

/*
 * Same product with B given as `parts` (<= 8) row blocks that live in different allocations: block q
 * holds rows [part_begin[q], part_begin[q+1]) of B, row-major with stride ldb, at B_parts[q]
 * (part_begin[0] == 0, part_begin[parts] == N).  B_parts and part_begin are HOST arrays of `parts` and
 * `parts + 1` entries; the pointers in B_parts are device pointers readable from the current device --
 * in the multi-GPU use they are the other ranks' row blocks of B, mapped through CUDA IPC, and the
 * kernel gathers their rows over NVLink while it computes: there is no replication step and no
 * replicated copy of B.  (New: the reference is single-GPU; its B is one array.)
 * Requires K % 4 == 0, ldb % 4 == 0, ldc % 4 == 0 and 16-byte aligned blocks and C.
 */
int gespmm_csr_spmm_f32_bparts(int64_t M, int64_t N, int64_t K, int64_t nnz,
                               const int32_t *rowptr, const int32_t *colind, const float *val,
                               int parts, const float *const *B_parts, const int64_t *part_begin,
                               int64_t ldb, float *C, int64_t ldc, void *stream);

/*
 * C[r, :] = max over the nonzeros p of row r of (val[p] *) B[colind[p], :], starting from `init`, which is
 * also what an empty row yields.  Replaces topoSimple/topoCache/topoCacheCoarsenSPMMMaxKernel and
 * XTopoCsrmmmax<float> (dgl-custom/binary_reduce_max.cu:26-168, 170-204): pass val = NULL and
 * init = -10000.0f for that code's exact results (its max_init(), :22-24), or -INFINITY for a true maximum.
 * The comparison is the reference's `acc > x ? acc : x` (:18-20).
 */
int gespmm_csr_spmm_max_f32(int64_t M, int64_t N, int64_t K, int64_t nnz,
                            const int32_t *rowptr, const int32_t *colind, const float *val,
                            const float *B, int64_t ldb, float *C, int64_t ldc, float init, void *stream);

/*
 * Per-call options of the product (gespmm_csr_spmm_f32_ex).  Everything a caller could previously only choose
 * through process-wide GESPMM_* environment variables is a field here, so that a library embedded in a
 * trainer can pick the summation order (or fuse its normalisation) per call and per thread.
 * Initialise with gespmm_opts_init(), then set fields; unknown / zero fields mean "automatic".
 */
#define GESPMM_FLAG_SEQUENTIAL  0x1u  /* sum every row of at most GESPMM_LONG_ROW nonzeros in strictly sequential CSR
                                         order for every K (bit-identical to the reference kernels); without it the
                                         faster re-associating sub-warp walker is used for K <= 64 (K % 4 == 0) and
                                         for any K <= 16 */
#define GESPMM_FLAG_NO_OVERLAP  0x2u  /* run the long-row kernel on `stream` itself, not on the helper stream */

/* values of gespmm_opts.walker (tuning / comparisons; 0 = automatic).  A walker asked for at a width it does not cover
 * falls back to the walker of that width with the same summation order (e.g. any request but AUTO / SUBWARP at
 * K <= 16 on 4-byte slices runs the sequential row-group kernel). */
#define GESPMM_WALKER_AUTO      0
#define GESPMM_WALKER_RING      1     /* cp.async gather ring, one nonzero per warp-wide copy, sequential order */
#define GESPMM_WALKER_REGISTER  2     /* register-staged gathers (no shared memory), sequential order           */
#define GESPMM_WALKER_SUBWARP   3     /* K <= 64: 2 / 4 / 8 nonzeros per warp-wide copy, re-associated           */
#define GESPMM_WALKER_ROWS      4     /* K <= 64: lane groups own disjoint rows, sequential order                */
#define GESPMM_WALKER_BULK      5     /* K > 64: one TMA bulk copy (cp.async.bulk + mbarrier) per gathered row,
                                         sequential order; the only walker whose gathers can carry l2_policy        */

typedef struct gespmm_opts {
    uint32_t struct_size;    /* sizeof(gespmm_opts), set by gespmm_opts_init (ABI versioning)                    */
    uint32_t flags;          /* GESPMM_FLAG_*                                                                    */
    int64_t  max_row_nnz;    /* longest row of A if the caller knows it (gespmm_max_row_nnz), else -1: when it is
                                <= the long-row threshold the long-row kernel is not launched at all            */
    /* Fused epilogue / prologue of a GCN layer (pytorch-custom/op.py:142-147), all device pointers, all nullable:
     *   C[r, c] = (sum_p val[p] * (B[colind[p], c] * col_scale[colind[p]])) * row_scale[r] + bias[c]
     * Every product and sum is a separately rounded fp32 operation in exactly this order, so the result is
     * bit-identical to scaling B, running the plain product, scaling C and adding the bias in four passes. */
    const float *row_scale;  /* [M]  */
    const float *col_scale;  /* [N]  */
    const float *bias;       /* [K]  */
    int32_t  walker;         /* GESPMM_WALKER_*                                                                  */
    int32_t  task_keys;      /* keys (rows + nonzeros) per task of the main kernel, multiple of 32 in [32, 1024]  */
    int32_t  long_row;       /* long-row threshold override, [512, 2^20]                                         */
    int32_t  panel_v;        /* 128-column blocks per pass (1..4)                                                */
    int32_t  l2_policy;      /* experimental: near + 4 far + 16 store, each 0 normal / 1 evict_first / 2 evict_last /
                                3 unchanged: L2 eviction priorities of the C stores (any walker) and of the gathered
                                rows (GESPMM_WALKER_BULK only); same bits whatever the value.  DESIGN.md 3.6b     */
    int32_t  l2_window_rows; /* experimental: |col - row| up to which a gathered row counts as "near" (0 = default
                                131072, negative = no band at all)                                             */
    const uint32_t *hot_columns; /* experimental: device bitmap, ceil(N / 32) words, bit c set = row c of B is among
                                the most referenced ones ("hot": always near); NULL = none           */
    void    *workspace;      /* optional device scratch, 256-byte aligned, of at least                            */
    size_t   workspace_bytes;/* gespmm_pad_workspace_bytes(M, N, K, nnz) bytes: lets a width that is not a multiple of 4
                                (K > 16) run on the 16-byte-slice walkers through padded copies of B and C when the graph
                                is dense enough for that to pay (nnz >= 4 (M + N)): ~2x the 4-byte-slice path at K = 41,
                                47.  Same bits either way (sequential order).  Without it nothing is padded.            */
} gespmm_opts;

/* Bytes of gespmm_opts.workspace that the padded path needs for this product; 0 when the library would not pad
 * (K % 4 == 0, K <= 16, or a graph of fewer than 4 nonzeros per row of B and C: nnz < 4 (M + N)). */
size_t gespmm_pad_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t nnz);

void gespmm_opts_init(gespmm_opts *opts);

/* gespmm_csr_spmm_f32 with per-call options (opts == NULL: exactly gespmm_csr_spmm_f32). */
int gespmm_csr_spmm_f32_ex(int64_t M, int64_t N, int64_t K, int64_t nnz,
                           const int32_t *rowptr, const int32_t *colind, const float *val,
                           const float *B, int64_t ldb, float *C, int64_t ldc,
                           const gespmm_opts *opts, void *stream);

/*
 * Longest row of a device CSR: *out_host = max_r (rowptr[r+1] - rowptr[r]).  One small reduction kernel and a
 * 4-byte copy back; synchronises `stream`.  A per-graph quantity: compute it once, pass it in
 * gespmm_opts.max_row_nnz on every product over that graph (optional caller-owned plan).
 */
int gespmm_max_row_nnz(int64_t M, const int32_t *rowptr, int32_t *out_host, void *stream);

/*
 * The GESPMM_* tuning environment variables (GESPMM_TASK, GESPMM_LONG, GESPMM_VARIANT, GESPMM_SEQUENTIAL,
 * GESPMM_PANEL_V, GESPMM_OVERLAP, GESPMM_SUBWARP_MAX_K, GESPMM_L2_POLICY, GESPMM_L2_WINDOW) are read ONCE, at the first call
 * into the library; per-call choices belong in gespmm_opts.  gespmm_reload_env re-reads them (sweep
 * scripts and tests that change the environment of a running process; not thread-safe against concurrent calls).
 */
void gespmm_reload_env(void);

/*
 * The library's only hidden state is one helper stream + two events per (host thread, device), created on the first
 * product of a matrix that may hold long rows (conventions above).  A host thread that is about to exit -- or a host
 * that wants the device reset cleanly -- releases its own with this call; the next product re-creates them.
 */
void gespmm_thread_cleanup(void);

/* gespmm_row_sum_is_sequential for a call made with `opts` (NULL: same as the plain function). */
int gespmm_row_sum_is_sequential_ex(int64_t K, int64_t row_nnz, const gespmm_opts *opts);

/* Enable reads of `peer_device`'s memory from kernels on the current device (idempotent). */
int gespmm_enable_peer_access(int peer_device);

/*
 * Map a device allocation exported by another process (a 64-byte cudaIpcMemHandle_t from
 * gespmm_ipc_alloc / cudaIpcGetMemHandle) for kernels on the current device; peer access
 * to the owning GPU is enabled by the mapping.  *base receives the address of the START of the exported
 * allocation.  gespmm_ipc_close unmaps it.
 */
int gespmm_ipc_open(const unsigned char *handle64, void **base);
int gespmm_ipc_close(void *base);
/* cudaMalloc'ed block on the current device plus its IPC handle (what a rank exports of its B block). */
int gespmm_ipc_alloc(size_t bytes, void **dptr, unsigned char *handle64);
int gespmm_ipc_free(void *dptr);

/* Rows with more nonzeros than this take the segmented path described above. */
#define GESPMM_LONG_ROW 4096

/*
 * 1 if gespmm_csr_spmm_f32 sums a row of `row_nnz` nonzeros of a product of width K in the reference's
 * strictly sequential CSR order (its result is then bit-identical to the reference kernels'), 0 if the
 * row's sum is re-associated (deterministically): rows longer than GESPMM_LONG_ROW, and -- where the
 * sub-warp walker for narrow B (K <= 64 in 16-byte slices, any K <= 16 in 4-byte slices: several nonzeros per
 * warp-wide gather, one partial sum per lane group) is in use -- rows of more than one nonzero.  For K % 4 == 0 it
 * assumes aligned operands (ldb, ldc multiples of 4, 16-byte aligned B and C); unaligned ones take the sub-warp
 * walker only up to K = 16 and a sequential walker above.  Pure function of its arguments and of the GESPMM_*
 * tuning environment as read by the library (gespmm_reload_env).
 * (New: every reference kernel is sequential, pytorch-custom/spmm_kernel.cu:56-59, 165-168.)
 */
int gespmm_row_sum_is_sequential(int64_t K, int64_t row_nnz);

/*
 * Same product with HOST buffers: allocates device buffers on `device`, copies in, runs
 * gespmm_csr_spmm_f32, copies C back, frees, synchronises.  For callers without their own
 * device memory management (what the CLI does by hand, spmm_test.cu:609-640).
 */
int gespmm_csr_spmm_f32_host(int64_t M, int64_t N, int64_t K, int64_t nnz,
                             const int32_t *rowptr, const int32_t *colind, const float *val,
                             const float *B, int64_t ldb, float *C, int64_t ldc, int device);

/*
 * CSR -> CSC (i.e. the CSR of A^T), the format SPMMFunction.backward needs (op.py:20-36).
 *   in : rowptr[M+1], colind[nnz], val[nnz] (nullable)                    device
 *   out: colptr[N+1], rowind[nnz], csc_val[nnz] (NULL iff val is NULL)    device
 * Within a column, entries are ordered by increasing row (stable sort of the CSR order by
 * column), which is what cusparseCsr2cscEx2 and scipy's tocsc() produce.  Deterministic.
 * workspace: device scratch of at least gespmm_csr2csc_workspace_bytes(M, N, nnz) bytes.
 */
size_t gespmm_csr2csc_workspace_bytes(int64_t M, int64_t N, int64_t nnz);
int gespmm_csr2csc_f32(int64_t M, int64_t N, int64_t nnz,
                       const int32_t *rowptr, const int32_t *colind, const float *val,
                       int32_t *colptr, int32_t *rowind, float *csc_val,
                       void *workspace, size_t workspace_bytes, void *stream);

/*
 * MatrixMarket coordinate file -> host CSR with readMtx's post-conditions
 * (util/util.hpp:286-333): 0-based, sorted by (row, col); `symmetric` files are mirrored and
 * then self-loops and duplicate entries are dropped; `general` files keep both; `pattern`
 * values are 1; `complex` files yield no entries.  Values are the file's values (the CLI
 * then overwrites them with 1, spmm_test.cu:574 -- that is the caller's business).
 * The three arrays are malloc'ed by the library; release them with gespmm_free_host.
 */
int gespmm_read_mtx(const char *path, int32_t *nrows, int32_t *ncols, int64_t *nnz,
                    int32_t **rowptr, int32_t **colind, float **val);
void gespmm_free_host(void *p);

/*
 * Binary image of a host CSR (native endianness, tagged; 64-byte header, rowptr, colind, val) and
 * gespmm_read_mtx through such an image: the reference re-parses its .mtx files on every run of the
 * CLI (spmm_test.cu:537, run_test.sh:5-10); here the parsed arrays can be kept next to the file.
 *   gespmm_write_csr / gespmm_read_csr   plain save / load (load checks the CSR invariants: monotone
 *                                        rowptr, rowptr[nrows] == nnz, 0 <= colind < ncols).
 *   gespmm_read_mtx_cached               `cache_path` (NULL: "<path>.gespmm-csr") is used when it exists
 *                                        and records the .mtx's current size and modification time;
 *                                        otherwise the .mtx is parsed and the image (re)written, best
 *                                        effort -- a read-only directory does not fail the call.
 *                                        *cache_hit (nullable) tells which happened.
 * Arrays are malloc'ed by the library; release them with gespmm_free_host.
 */
/*
 * Host CSR -> MatrixMarket coordinate file, `general` symmetry, 1-based, one entry per line in CSR order; field
 * `pattern` when val is NULL, `real` (%.9g: fp32 round-trips) otherwise.  gespmm_read_mtx of the result gives the
 * arrays back.  (The reference's data/conv.c rewrites .mtx files entry by entry with fprintf.)
 */
int gespmm_write_mtx(const char *path, int32_t nrows, int32_t ncols, int64_t nnz,
                     const int32_t *rowptr, const int32_t *colind, const float *val);

int gespmm_write_csr(const char *path, int32_t nrows, int32_t ncols, int64_t nnz,
                     const int32_t *rowptr, const int32_t *colind, const float *val);
int gespmm_read_csr(const char *path, int32_t *nrows, int32_t *ncols, int64_t *nnz,
                    int32_t **rowptr, int32_t **colind, float **val);
int gespmm_read_mtx_cached(const char *path, const char *cache_path, int32_t *nrows, int32_t *ncols, int64_t *nnz,
                           int32_t **rowptr, int32_t **colind, float **val, int *cache_hit);

#ifdef __cplusplus
}
#endif
#endif /* GESPMM_H */
