// gespmm_spmm.cu -- host side of the fp32 CSR x dense SpMM: per-call options, the tuning environment, the choice of
// walker and launch shape, and the C-ABI entry points (include/gespmm.h).  The kernels and walkers are templates in
// gespmm_spmm_kernels.cuh, instantiated family by family in gespmm_spmm_*.cu.
#include "gespmm_spmm_kernels.cuh"

namespace gespmm_detail {

Side *thread_sides()
{
    thread_local Side sides[kMaxDevices];
    return sides;
}

}  // namespace gespmm_detail

using namespace gespmm_detail;

namespace {

int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

constexpr int kMaxLong = 1 << 20;  // largest accepted long-row threshold: a 32-row run then spans < 2^25 nonzeros
constexpr int kSubwarpMaxKDefault = 64;  // measured on B200 (profiles/r01_sweep_narrow.txt): 1.1-6x the ring walker at K <= 64
constexpr int kL2WindowDefault = 1 << 17;  // rows: 64 MB of 512-byte panel rows, about half of the L2

// The GESPMM_* tuning environment, read once (gespmm_reload_env re-reads it):
//   GESPMM_VARIANT unset / < 0 : automatic -- the sub-warp walker for K <= GESPMM_SUBWARP_MAX_K, else the ring walker
//     0 ring walker (sequential order for every K)   1 register-staged walker (comparisons)
//     2 sub-warp walker wherever it applies (K <= 64)   4 row-parallel narrow walker wherever it applies (sequential)
//     5 bulk walker (one TMA bulk copy per gathered row) for K > 64
//   GESPMM_SEQUENTIAL=1 : the fastest walker that keeps the reference's order for every K (= variant 4)
//   GESPMM_TASK / GESPMM_LONG / GESPMM_PANEL_V / GESPMM_OVERLAP / GESPMM_SMEM_PAD : launch-shape overrides
//   GESPMM_L2_POLICY = near + 4 far + 16 store (each 0 normal, 1 evict_first, 2 evict_last, 3 unchanged), GESPMM_L2_WINDOW rows
struct Tuning {
    int task, long_row, walker, panel_v, overlap, subwarp_max_k, l2_policy, l2_window, smem_pad;
};
int walker_of_variant(int variant)
{
    switch (variant) {
        case 0: return GESPMM_WALKER_RING;
        case 1: return GESPMM_WALKER_REGISTER;
        case 2: return GESPMM_WALKER_SUBWARP;
        case 4: return GESPMM_WALKER_ROWS;
        case 5: return GESPMM_WALKER_BULK;
        default: return GESPMM_WALKER_AUTO;
    }
}
Tuning read_env()
{
    Tuning t;
    t.task = env_int("GESPMM_TASK", 0);
    const int forced = env_int("GESPMM_LONG", 0);
    t.long_row = forced < kMinLong ? GESPMM_LONG_ROW : (forced > kMaxLong ? kMaxLong : forced);
    t.walker = env_int("GESPMM_SEQUENTIAL", 0) != 0 ? GESPMM_WALKER_ROWS : walker_of_variant(env_int("GESPMM_VARIANT", -1));
    t.panel_v = env_int("GESPMM_PANEL_V", 0);
    t.overlap = env_int("GESPMM_OVERLAP", 1);
    t.subwarp_max_k = env_int("GESPMM_SUBWARP_MAX_K", kSubwarpMaxKDefault);
    t.l2_policy = env_int("GESPMM_L2_POLICY", 0);
    t.l2_window = env_int("GESPMM_L2_WINDOW", kL2WindowDefault);
    t.smem_pad = env_int("GESPMM_SMEM_PAD", 0);
    return t;
}
Tuning g_tuning;
std::once_flag g_tuning_once;
const Tuning &tuning()
{
    std::call_once(g_tuning_once, [] { g_tuning = read_env(); });
    return g_tuning;
}

// One call's choices: the environment's, overridden field by field by the caller's gespmm_opts.
struct Choice {
    int walker, task, long_row, panel_v, l2_policy, l2_window, smem_pad;
    bool overlap;
    long long max_row_nnz;
    const float *row_scale, *col_scale, *bias;
    const unsigned *hot;
    void *workspace;
    size_t workspace_bytes;
    bool fuse() const { return row_scale || col_scale || bias; }
};
Choice choose(const gespmm_opts *o)
{
    const Tuning &t = tuning();
    Choice c;
    c.walker = t.walker; c.task = t.task; c.long_row = t.long_row; c.panel_v = t.panel_v; c.l2_policy = t.l2_policy;
    c.l2_window = t.l2_window; c.smem_pad = t.smem_pad; c.overlap = t.overlap != 0; c.max_row_nnz = -1;
    c.row_scale = c.col_scale = c.bias = nullptr;
    c.hot = nullptr;
    c.workspace = nullptr; c.workspace_bytes = 0;
    if (o) {
        c.hot = o->hot_columns;
        c.workspace = o->workspace; c.workspace_bytes = o->workspace_bytes;
        if (o->walker > 0) c.walker = o->walker;
        if (o->flags & GESPMM_FLAG_SEQUENTIAL) c.walker = GESPMM_WALKER_ROWS;
        if (o->flags & GESPMM_FLAG_NO_OVERLAP) c.overlap = false;
        if (o->task_keys > 0) c.task = o->task_keys;
        if (o->long_row >= kMinLong) c.long_row = o->long_row > kMaxLong ? kMaxLong : o->long_row;
        if (o->panel_v > 0) c.panel_v = o->panel_v;
        if (o->l2_policy > 0) c.l2_policy = o->l2_policy;
        if (o->l2_window_rows > 0) c.l2_window = o->l2_window_rows;
        else if (o->l2_window_rows < 0) c.l2_window = -1;  // no band: only hot columns count as near
        c.max_row_nnz = o->max_row_nnz;
        c.row_scale = o->row_scale; c.col_scale = o->col_scale; c.bias = o->bias;
    }
    return c;
}
bool use_rows(int64_t K, int walker) { return walker == GESPMM_WALKER_ROWS && K <= 64 && K % 4 == 0; }
bool use_subwarp(int64_t K, int walker)
{
    if (K > 64 || K % 4 != 0) return false;
    if (walker == GESPMM_WALKER_SUBWARP) return true;
    return walker == GESPMM_WALKER_AUTO && K <= tuning().subwarp_max_k;
}

// `mode`: 0 sum, 1 max, 2 fused sum with a gathered-row scale, 3 fused sum without one (row scale / bias only)
cudaError_t dispatch_all(bool valued, int mode, bool vec4, bool peer, int walker, bool hint, int V, bool masked, int K, const Args &a)
{
    if (!vec4 && K <= kRowGroupMaxK && walker != GESPMM_WALKER_REGISTER) {
        // tiny rows of any width: lane groups on 4-byte slices.  Default (and GESPMM_WALKER_SUBWARP): the groups share a
        // row's nonzeros (sub-warp walker, re-associated like its 16-byte form); any other request -- GESPMM_FLAG_SEQUENTIAL,
        // or a walker that sums in CSR order at the widths it was made for -- gets the sequential order: every group sums
        // its own row (row-group kernel).  sequential_for() below says the same.
        return run_lanegroup(mode, valued, !(walker == GESPMM_WALKER_AUTO || walker == GESPMM_WALKER_SUBWARP), K, a);
    }
    if (!vec4) return run_scalar(mode, valued, walker == GESPMM_WALKER_REGISTER, V, a);
    auto ring = valued ? run_ring_valued : run_ring_unvalued;
    if (peer) return ring(0, true, false, V, masked, a);
    if (use_rows(K, walker)) return run_narrow(mode, valued, true, K, a);
    if (use_subwarp(K, walker)) return run_narrow(mode, valued, false, K, a);
    if (mode != 0) return ring(mode, false, false, V, masked, a);  // mode 2: V == 1 (see run_spmm)
    if (walker == GESPMM_WALKER_REGISTER) return run_other(valued, false, false, V, masked, a);
    if (walker == GESPMM_WALKER_BULK && V == 1) return run_other(valued, true, hint, V, masked, a);
    return ring(0, false, hint && V == 1, V, masked, a);
}

// Row padding for widths that are not multiples of 4 (gespmm_opts.workspace): dst[r, 0:K4) = src[r, 0:K) followed by zeros,
// one float4 store per thread; and the way back, dst[r, 0:K) = src[r, 0:K) of a K4-strided source.
__global__ void pad_rows_kernel(long long rows, int K, int K4, const float *__restrict__ src, long long lds, float *__restrict__ dst)
{
    const int q4 = K4 >> 2;
    const long long total = rows * q4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / q4;
        const int c = (int)(i - r * q4) << 2;
        const float *s = src + r * lds + c;
        float4 v;
        v.x = c + 0 < K ? __ldg(s + 0) : 0.f; v.y = c + 1 < K ? __ldg(s + 1) : 0.f;
        v.z = c + 2 < K ? __ldg(s + 2) : 0.f; v.w = c + 3 < K ? __ldg(s + 3) : 0.f;
        reinterpret_cast<float4 *>(dst)[i] = v;
    }
}
__global__ void unpad_rows_kernel(long long rows, int K, int K4, const float *__restrict__ src, float *__restrict__ dst, long long ldd)
{
    const long long total = rows * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / K;
        const int c = (int)(i - r * K);
        __stcs(dst + r * ldd + c, __ldcs(src + r * K4 + c));
    }
}

__global__ void max_row_kernel(int M, const int *__restrict__ rowptr, int *__restrict__ out)
{
    int best = 0;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < M; r += (long long)gridDim.x * blockDim.x)
        best = max(best, __ldg(rowptr + r + 1) - __ldg(rowptr + r));
    best = __reduce_max_sync(kFull, best);
    if ((threadIdx.x & 31) == 0 && best > 0) atomicMax(out, best);
}

}  // namespace

namespace {

size_t align64(size_t floats) { return (floats + 63) & ~(size_t)63; }  // 256-byte aligned sub-buffers
size_t pad_workspace_bytes(int64_t M, int64_t N, int64_t K)
{
    const size_t K4 = (size_t)((K + 3) & ~3LL);
    return (align64((size_t)N * K4) + align64((size_t)M * K4) + align64(K4)) * sizeof(float);
}

// Shared by the entry points: argument checks, task window, dispatch.  `parts` == 0: B is one array.
int run_spmm(int64_t M, int64_t N, int64_t K, int64_t nnz, const int32_t *rowptr, const int32_t *colind, const float *val,
             const float *B, int parts, const float *const *B_parts, const int64_t *part_begin, int64_t ldb, float *C,
             int64_t ldc, void *stream, const gespmm_opts *opts, bool max_reduce = false, float init = 0.f)
{
    if (M < 0 || N < 0 || K < 0 || nnz < 0) return GESPMM_ERR_INVALID_ARG;
    // positions are summed in int inside the kernels (p + 64 + lane, a + warp * seg, ...): keep 64 K of headroom
    if (M > INT32_MAX - 64 || N > INT32_MAX || nnz > INT32_MAX - 65536 || K > INT32_MAX || ldb >= (1LL << 30) || ldc > INT32_MAX)
        return GESPMM_ERR_TOO_LARGE;
    if (opts && opts->struct_size != sizeof(gespmm_opts)) return GESPMM_ERR_INVALID_ARG;
    if (M == 0 || K == 0) return GESPMM_OK;
    if (ldb < K || ldc < K) return GESPMM_ERR_INVALID_ARG;
    if (!rowptr || !C) return GESPMM_ERR_INVALID_ARG;
    if (nnz > 0 && !colind) return GESPMM_ERR_INVALID_ARG;
    if (parts == 0 && nnz > 0 && !B) return GESPMM_ERR_INVALID_ARG;
    const Choice ch = choose(opts);
    if (ch.fuse() && (max_reduce || parts > 0)) return GESPMM_ERR_INVALID_ARG;  // the fused scaling belongs to the plain sum

    bool vec4 = (K % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    if (ch.bias && (reinterpret_cast<uintptr_t>(ch.bias) & 15)) vec4 = false;
    Args a;
    a.op.peer.parts = 0;
    if (parts > 0) {
        if (parts > kMaxParts || !B_parts || !part_begin) return GESPMM_ERR_INVALID_ARG;
        if (part_begin[0] != 0 || part_begin[parts] != N) return GESPMM_ERR_INVALID_ARG;
        for (int q = 0; q < parts; q++) {
            if (part_begin[q + 1] < part_begin[q]) return GESPMM_ERR_INVALID_ARG;
            if (part_begin[q + 1] > part_begin[q] && !B_parts[q]) return GESPMM_ERR_INVALID_ARG;
            if (reinterpret_cast<uintptr_t>(B_parts[q]) & 15) return GESPMM_ERR_INVALID_ARG;  // the sharded path is the aligned path
            a.op.peer.base[q] = B_parts[q];
            a.op.peer.lo[q] = (int)part_begin[q];
        }
        for (int q = parts; q < kMaxParts; q++) a.op.peer.base[q] = nullptr;
        for (int q = parts; q <= kMaxParts; q++) a.op.peer.lo[q] = (int)N;
        a.op.peer.parts = parts;
        if (!vec4) return GESPMM_ERR_INVALID_ARG;  // needs K % 4 == 0 and 16-byte aligned, 4-float-strided operands
    } else {
        vec4 = vec4 && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    }
    // A width that is not a multiple of 4, above the lane-group walkers' K <= 16, with a caller-given workspace: B is
    // copied into rows padded to K4 = 4 ceil(K / 4) floats, the product runs on the 16-byte-slice walkers IN SEQUENTIAL
    // ORDER (so the bits are those of the unpadded path and of the reference, whichever path a graph takes), and C comes
    // back through the same padding.  The 4-byte-slice ring walker is instruction-bound (ogbn-products shape, K = 41 / 47:
    // 5.5 / 5.8 ms against 3.1 / 2.7 padded); the two extra passes over B and C cost about as much as 2.5 nonzeros per row
    // of B and C, so graphs sparser than 4 nonzeros per (row of B + row of C) stay unpadded (cit-Patents shape: 0.98 vs 1.27).
    if (parts == 0 && K % 4 != 0 && K > kRowGroupMaxK && ch.workspace && ch.walker != GESPMM_WALKER_REGISTER &&
        nnz >= 4 * (M + N)) {
        const int64_t K4 = (K + 3) & ~3LL;
        const size_t need = pad_workspace_bytes(M, N, K);
        if (ch.workspace_bytes < need || (reinterpret_cast<uintptr_t>(ch.workspace) & 255)) return GESPMM_ERR_WORKSPACE;
        float *Bp = static_cast<float *>(ch.workspace);
        float *Cp = Bp + align64((size_t)N * K4);
        float *biasp = Cp + align64((size_t)M * K4);
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const int blocks = 148 * 8;
        if (nnz > 0) pad_rows_kernel<<<blocks, 256, 0, st>>>(N, (int)K, (int)K4, B, ldb, Bp);
        if (ch.bias) pad_rows_kernel<<<1, 64, 0, st>>>(1, (int)K, (int)K4, ch.bias, K, biasp);
        gespmm_opts inner;
        if (opts) inner = *opts; else gespmm_opts_init(&inner);
        inner.workspace = nullptr; inner.workspace_bytes = 0;
        inner.flags |= GESPMM_FLAG_SEQUENTIAL;
        if (ch.bias) inner.bias = biasp;
        const int rc = run_spmm(M, N, K4, nnz, rowptr, colind, val, Bp, 0, nullptr, nullptr, K4, Cp, K4, stream, &inner, max_reduce, init);
        if (rc != GESPMM_OK) return rc;
        unpad_rows_kernel<<<blocks, 256, 0, st>>>(M, (int)K, (int)K4, Cp, C, ldc);
        return cudaGetLastError() == cudaSuccess ? GESPMM_OK : GESPMM_ERR_CUDA;
    }
    const int W = vec4 ? 4 : 1;
    const int packs = (int)((K + 32 * W - 1) / (32 * W));
    // Packs per lane = panel width / (32 W) columns; blockIdx.y walks the panels one after the other, each
    // pass re-reading the CSR arrays.  With B in local memory a pass over a 128-column panel (V = 1) gathers
    // from an N x 512-byte slice of B, a quarter of the L2 footprint of a 512-column pass, and that outweighs
    // the extra colind reads on every shape measured (K = 256 / 512: Reddit 1.13x / 1.34x, R-MAT 1.14x / 1.19x,
    // cit-Patents 1.10x / 1.17x, ogbn-products 1.04x; profiles/r01_sweep_panel.txt).  Sharded B keeps wide
    // panels (one 2 KB request per remote row instead of four).  panel_v (1..4) overrides.
    const int max_v = packs >= 4 ? 4 : packs;
    int V = (vec4 && parts == 0) ? 1 : max_v;
    if (ch.panel_v >= 1 && ch.panel_v <= max_v && !(vec4 && ch.fuse())) V = ch.panel_v;
    const bool masked = (K % (32 * W * V)) != 0;  // some lanes' packs fall beyond K

    // Task window (keys per task).  A task's start-up (row search, first rowptr / colind fetch) is
    // amortised over its window, but its rows are walked 32 at a time, so the best window grows with
    // the average row: measured optima on B200 are ~96 keys at 5 keys/row (cit-Patents shape, at 2.5 M
    // to 20 M keys: 1.080 ms at 96, 1.096 at 64, 1.103 at 128, 1.157 at 192; profiles/r02_sweep_l2_bulk.txt), ~256 at 21
    // (R-MAT), 512-1024 at ~500 (Reddit shape), i.e. ~48*sqrt(keys per row);
    // capped so that the grid keeps at least ~8 waves of resident CTAs.
    const bool sub = vec4 && parts == 0 && use_subwarp(K, ch.walker);
    const bool rows_walker = vec4 && parts == 0 && use_rows(K, ch.walker);
    const long long total = nnz + M;
    const long long warps_per_wave = 148LL * 24;
    const double keys_per_row = (double)total / (double)M;
    long long tk = ((long long)(48.0 * sqrt(keys_per_row)) + 16) & ~31LL;  // nearest multiple of 32 (96 at 5.4 keys per row)
    // the sub-warp walker spends 1/NG of the instructions per nonzero, so a task's start-up weighs NG times more:
    // measured optima are 512 (cit-Patents shape) to 1024 (ogbn-products, R-MAT, Reddit shapes) at K = 16, 32
    if (sub || rows_walker) tk *= (K > 32 ? 2 : (K > 16 ? 4 : 8));
    const bool rowgroup = !vec4 && K <= kRowGroupMaxK && ch.walker != GESPMM_WALKER_REGISTER;
    if (rowgroup) tk *= (K > 8 ? 2 : (K > 4 ? 4 : 8));  // NG rows at a time: a task's start-up weighs NG times more
    const long long cap = (total / (8 * warps_per_wave)) & ~31LL;
    if (tk > cap) tk = cap;
    int task = (int)(tk < 32 ? 32 : (tk > kMaxTask ? kMaxTask : tk));
    if (ch.task >= 32 && ch.task <= kMaxTask) task = ch.task & ~31;

    a.M = (int)M; a.K = (int)K; a.task = task; a.long_row = ch.long_row; a.nnz = nnz; a.rowptr = rowptr;
    a.op.colind = colind; a.op.val = val; a.op.B = B; a.op.C = C; a.op.ldb = (int)ldb; a.op.ldc = (int)ldc;
    a.st = static_cast<cudaStream_t>(stream);
    a.overlap = ch.overlap;
    a.has_long = nnz > ch.long_row && (ch.max_row_nnz < 0 || ch.max_row_nnz > ch.long_row);
    a.smem_pad = ch.smem_pad;
    a.op.init = init;
    a.op.row_scale = ch.row_scale; a.op.col_scale = ch.col_scale; a.op.bias = ch.bias;
    a.op.l2_near = ch.l2_policy & 3; a.op.l2_far = (ch.l2_policy >> 2) & 3; a.op.l2_store = (ch.l2_policy >> 4) & 3;
    a.op.l2_window = ch.l2_window;
    a.op.l2_hot = ch.hot;
    const int mode = max_reduce ? 1 : (ch.fuse() ? (ch.col_scale ? 2 : 3) : 0);  // 3: row scale / bias only, nothing per nonzero
    const bool hint = ch.l2_policy > 0 && mode == 0;
    const cudaError_t err = dispatch_all(val != nullptr, mode, vec4, parts > 0, ch.walker, hint, V, masked, (int)K, a);
    return err == cudaSuccess ? GESPMM_OK : GESPMM_ERR_CUDA;
}

int sequential_for(int64_t K, int64_t row_nnz, const Choice &ch)
{
    if (row_nnz > ch.long_row) return 0;  // segmented (kernel B)
    // (a width that is not a multiple of 4 above 16 is summed in CSR order with or without a padding workspace)
    if (row_nnz > 1 && use_subwarp(K, ch.walker) && !use_rows(K, ch.walker)) return 0;  // per-group partial sums
    // widths that are not multiples of 4 up to 16: the same sub-warp walker on 4-byte slices unless a sequential
    // walker was asked for (row-group kernel / register walker)
    if (row_nnz > 1 && K % 4 != 0 && K <= kRowGroupMaxK &&
        (ch.walker == GESPMM_WALKER_AUTO || ch.walker == GESPMM_WALKER_SUBWARP)) return 0;
    return 1;
}

}  // namespace

extern "C" size_t gespmm_pad_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t nnz)
{
    if (M < 0 || N < 0 || K <= kRowGroupMaxK || K % 4 == 0 || nnz < 4 * (M + N)) return 0;
    return pad_workspace_bytes(M, N, K);
}

extern "C" void gespmm_opts_init(gespmm_opts *opts)
{
    if (!opts) return;
    memset(opts, 0, sizeof(*opts));
    opts->struct_size = (uint32_t)sizeof(*opts);
    opts->max_row_nnz = -1;
}

extern "C" void gespmm_thread_cleanup(void)
{
    Side *sides = thread_sides();
    int cur = 0;
    const bool have_cur = cudaGetDevice(&cur) == cudaSuccess;
    for (int d = 0; d < kMaxDevices; d++) {
        Side &sd = sides[d];
        if (!sd.ok) continue;
        if (cudaSetDevice(d) == cudaSuccess) {
            cudaStreamSynchronize(sd.stream);
            cudaEventDestroy(sd.join);
            cudaEventDestroy(sd.fork);
            cudaStreamDestroy(sd.stream);
        }
        sd = Side();
    }
    if (have_cur) cudaSetDevice(cur);
    cudaGetLastError();
}

extern "C" void gespmm_reload_env(void)
{
    tuning();  // the once-flag is spent first, so that a concurrent first call cannot overwrite the fresh values
    g_tuning = read_env();
}

extern "C" int gespmm_csr_spmm_f32(int64_t M, int64_t N, int64_t K, int64_t nnz, const int32_t *rowptr,
                                   const int32_t *colind, const float *val, const float *B, int64_t ldb,
                                   float *C, int64_t ldc, void *stream)
{
    return run_spmm(M, N, K, nnz, rowptr, colind, val, B, 0, nullptr, nullptr, ldb, C, ldc, stream, nullptr);
}

extern "C" int gespmm_csr_spmm_f32_ex(int64_t M, int64_t N, int64_t K, int64_t nnz, const int32_t *rowptr,
                                      const int32_t *colind, const float *val, const float *B, int64_t ldb,
                                      float *C, int64_t ldc, const gespmm_opts *opts, void *stream)
{
    return run_spmm(M, N, K, nnz, rowptr, colind, val, B, 0, nullptr, nullptr, ldb, C, ldc, stream, opts);
}

extern "C" int gespmm_max_row_nnz(int64_t M, const int32_t *rowptr, int32_t *out_host, void *stream)
{
    if (M < 0 || M > INT32_MAX - 64 || !out_host || (M > 0 && !rowptr)) return GESPMM_ERR_INVALID_ARG;
    *out_host = 0;
    if (M == 0) return GESPMM_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int *d = nullptr;
    if (cudaMalloc(&d, sizeof(int)) != cudaSuccess) { cudaGetLastError(); return GESPMM_ERR_CUDA; }
    int rc = GESPMM_ERR_CUDA;
    do {
        if (cudaMemsetAsync(d, 0, sizeof(int), st) != cudaSuccess) break;
        const int blocks = (int)((M + 255) / 256 < 148 * 8 ? (M + 255) / 256 : 148 * 8);
        max_row_kernel<<<blocks, 256, 0, st>>>((int)M, rowptr, d);
        if (cudaGetLastError() != cudaSuccess) break;
        if (cudaMemcpyAsync(out_host, d, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        rc = GESPMM_OK;
    } while (0);
    if (rc != GESPMM_OK) cudaGetLastError();
    cudaFree(d);
    return rc;
}

extern "C" int gespmm_row_sum_is_sequential(int64_t K, int64_t row_nnz) { return sequential_for(K, row_nnz, choose(nullptr)); }

extern "C" int gespmm_row_sum_is_sequential_ex(int64_t K, int64_t row_nnz, const gespmm_opts *opts)
{
    if (opts && opts->struct_size != sizeof(gespmm_opts)) return 0;
    return sequential_for(K, row_nnz, choose(opts));
}

extern "C" int gespmm_csr_spmm_f32_bparts(int64_t M, int64_t N, int64_t K, int64_t nnz, const int32_t *rowptr,
                                          const int32_t *colind, const float *val, int parts, const float *const *B_parts,
                                          const int64_t *part_begin, int64_t ldb, float *C, int64_t ldc, void *stream)
{
    if (parts < 1) return GESPMM_ERR_INVALID_ARG;
    return run_spmm(M, N, K, nnz, rowptr, colind, val, nullptr, parts, B_parts, part_begin, ldb, C, ldc, stream, nullptr);
}

extern "C" int gespmm_csr_spmm_max_f32(int64_t M, int64_t N, int64_t K, int64_t nnz, const int32_t *rowptr,
                                       const int32_t *colind, const float *val, const float *B, int64_t ldb,
                                       float *C, int64_t ldc, float init, void *stream)
{
    return run_spmm(M, N, K, nnz, rowptr, colind, val, B, 0, nullptr, nullptr, ldb, C, ldc, stream, nullptr, true, init);
}
