// gespmm_spmm.cu -- fp32 CSR x dense SpMM for sm_100a, and its C-ABI launcher.
//
// Replaces the reference kernels topoSimple/topoCache/topoCacheCoarsenSPMMKernel and
// spmm_test0..4 (pytorch-custom/spmm_kernel.cu:31-173, 210-379; spmm_test.cu:64-454) and
// their launch blocks (spmm_kernel.cu:175-207, 425-458; spmm_test.cu:456-492).  Not a
// port: the reference maps (row, 64 columns) to a warp and walks the row serially with
// 32-bit loads; here the unit of work is a fixed-size slice of the merged (rows + nonzeros)
// sequence, walked as one flat stream of nonzeros with 128-bit B-panel loads.
//
// Work decomposition
//   key(r) = rowptr[r] + r is strictly increasing, so the half-open key windows
//   [t*T, (t+1)*T) partition the rows: warp-task t owns the rows whose key falls in its
//   window.  Every task therefore costs at most T "row stores + nonzero gathers" plus the
//   tail of its last row, whatever the degree distribution (empty rows are work too: their
//   C rows must be zeroed).  A task finds its rows with a 16-ary search on rowptr (both
//   window ends at once, one per half-warp), so no per-graph preprocessing and no workspace.
//
// Flat stream
//   A warp walks its rows' nonzeros [rowptr[row_lo], rowptr[row_hi]) in CSR order.  Lanes
//   fetch 32 (colind, val) pairs with one coalesced load each (next chunk prefetched while
//   the current one is consumed), broadcast them with __shfl, and issue U independent
//   128-bit loads of B-row panels before the first FMA, so a warp keeps U*V*512 B of gathers
//   in flight regardless of how short the rows are.  Row ends are detected on the (warp-
//   uniform) nonzero counter: store C row, reset accumulators, continue.  Each output
//   element is accumulated in CSR order in one register, FFMA for valued / FADD for
//   unvalued, which is the reference's order -- results are bit-identical to it.
//
// Long rows
//   Rows with more than GESPMM_LONG_ROW nonzeros are not walked by their owner warp.  They
//   are queued in shared memory and, after a CTA barrier, summed by all warps of the CTA in
//   contiguous segments whose partials are combined in fixed order through shared memory
//   (deterministic; fp32 re-association only).
//
// Column mapping
//   Lane l owns, for v < V, the float4 at column ((v*32 + l) * 4) of the current panel
//   (panel = 128*V columns; blockIdx.y walks panels for K > 512).  One warp-wide LDG.128
//   therefore reads 512 contiguous bytes of a B row.  K % 4 != 0 or unaligned operands take
//   the scalar instantiation (lane owns column v*32 + l).

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "gespmm.h"

namespace {

constexpr int kWarps = 8;
constexpr int kCta = kWarps * 32;
constexpr int kLongRow = GESPMM_LONG_ROW;
constexpr int kMaxTask = 512;                               // largest task window (keys)
constexpr int kMaxLong = kWarps * kMaxTask / kLongRow + 2;  // long rows that can start in one CTA's windows
constexpr unsigned kFull = 0xffffffffu;

// ---- per-lane vector of owned columns: float4 (aligned fast path) or float (general) --------
template <bool VEC4> struct Pack;
template <> struct Pack<true> {
    using T = float4;
    static constexpr int kWidth = 4;
    static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ T ldg(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
    static __device__ __forceinline__ void stcs(float *p, const T &a) { __stcs(reinterpret_cast<float4 *>(p), a); }
    static __device__ __forceinline__ void fma(T &acc, float a, const T &b) {
        acc.x = fmaf(a, b.x, acc.x); acc.y = fmaf(a, b.y, acc.y);
        acc.z = fmaf(a, b.z, acc.z); acc.w = fmaf(a, b.w, acc.w);
    }
    static __device__ __forceinline__ void add(T &acc, const T &b) {
        acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
};
template <> struct Pack<false> {
    using T = float;
    static constexpr int kWidth = 1;
    static __device__ __forceinline__ T zero() { return 0.f; }
    static __device__ __forceinline__ T ldg(const float *p) { return __ldg(p); }
    static __device__ __forceinline__ void stcs(float *p, const T &a) { __stcs(p, a); }
    static __device__ __forceinline__ void fma(T &acc, float a, const T &b) { acc = fmaf(a, b, acc); }
    static __device__ __forceinline__ void add(T &acc, const T &b) { acc += b; }
};

// First r in [0, M) with rowptr[r] + r >= target, else M.  `half` lanes (16) cooperate; the two
// half-warps run independent searches in lock-step (uniform trip count = worst of the two).
__device__ __forceinline__ int search_key16(const int *__restrict__ rowptr, int M, long long target,
                                            int sub /* lane & 15 */, int shift /* 0 or 16 */)
{
    int lo = 0, hi = M;
    while (__any_sync(kFull, lo < hi)) {
        const int len = hi - lo;
        const int step = (len + 15) >> 4;
        const long long probe = (long long)lo + (long long)sub * step;
        bool below = false;
        if (len > 0 && probe < hi) below = ((long long)__ldg(rowptr + probe) + probe) < target;
        const unsigned bal = (__ballot_sync(kFull, below) >> shift) & 0xffffu;
        if (len > 0) {
            const int cnt = __popc(bal);  // monotone: the first `cnt` probes are below target
            if (cnt == 0) { hi = lo; }
            else {
                const long long nlo = (long long)lo + (long long)(cnt - 1) * step + 1;
                const long long nhi = (long long)lo + (long long)cnt * step;
                hi = (int)(nhi < hi ? nhi : hi);
                lo = (int)nlo;
                if (lo > hi) lo = hi;
            }
        }
    }
    return lo;
}

template <int V, bool VALUED, bool VEC4, int U>
struct Walker {
    using P = Pack<VEC4>;
    using T = typename P::T;

    const int *__restrict__ colind;
    const float *__restrict__ val;
    const float *__restrict__ Bl;   // B + this lane's first owned column
    float *__restrict__ Cl;         // C + this lane's first owned column
    long long ldb, ldc;
    unsigned vmask;                 // bit v set: this lane's v-th pack is inside K
    int lane;

    __device__ __forceinline__ void store_row(long long row, const T (&acc)[V]) const {
        float *c = Cl + row * ldc;
#pragma unroll
        for (int v = 0; v < V; v++)
            if (vmask & (1u << v)) P::stcs(c + v * 32 * P::kWidth, acc[v]);
    }

    // acc[] += sum over nonzeros [s, e) in CSR order; at every row end (taken from the lanes'
    // my_end registers, rows first..last of the current 32-row chunk based at row `rb`) the row is
    // stored and the accumulators reset.  With first > last no row is ever flushed (segment mode).
    __device__ __forceinline__ void stream(int s, int e, T (&acc)[V], int my_end, long long rb, int first,
                                           int last) const {
        int cur = first;
        int cur_end = (first <= last) ? __shfl_sync(kFull, my_end, first) : 0x7fffffff;
        int ncol = 0;
        float nval = 1.f;
        if (s + lane < e) {
            ncol = __ldcs(colind + s + lane);
            if (VALUED) nval = __ldcs(val + s + lane);
        }
        for (int p0 = s; p0 < e; p0 += 32) {
            const int mcol = ncol;
            const float mval = nval;
            const int pn = p0 + 32 + lane;
            if (pn < e) {
                ncol = __ldcs(colind + pn);
                if (VALUED) nval = __ldcs(val + pn);
            }
            const int n = min(32, e - p0);
#pragma unroll 1
            for (int j0 = 0; j0 < n; j0 += U) {
                T b[U][V];
                float a[U];
                if (j0 + U <= n) {
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int c = __shfl_sync(kFull, mcol, j0 + u);
                        if (VALUED) a[u] = __shfl_sync(kFull, mval, j0 + u);
                        const float *bp = Bl + (long long)c * ldb;
#pragma unroll
                        for (int v = 0; v < V; v++)
                            if (vmask & (1u << v)) b[u][v] = P::ldg(bp + v * 32 * P::kWidth);
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int p = p0 + j0 + u;
                        while (p >= cur_end) {
                            store_row(rb + cur, acc);
#pragma unroll
                            for (int v = 0; v < V; v++) acc[v] = P::zero();
                            ++cur;
                            cur_end = __shfl_sync(kFull, my_end, cur);
                        }
#pragma unroll
                        for (int v = 0; v < V; v++) {
                            if (VALUED) P::fma(acc[v], a[u], b[u][v]);
                            else P::add(acc[v], b[u][v]);
                        }
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int c = __shfl_sync(kFull, mcol, (j0 + u) & 31);
                        if (VALUED) a[u] = __shfl_sync(kFull, mval, (j0 + u) & 31);
                        if (j0 + u < n) {
                            const float *bp = Bl + (long long)c * ldb;
#pragma unroll
                            for (int v = 0; v < V; v++)
                                if (vmask & (1u << v)) b[u][v] = P::ldg(bp + v * 32 * P::kWidth);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        if (j0 + u < n) {
                            const int p = p0 + j0 + u;
                            while (p >= cur_end) {
                                store_row(rb + cur, acc);
#pragma unroll
                                for (int v = 0; v < V; v++) acc[v] = P::zero();
                                ++cur;
                                cur_end = __shfl_sync(kFull, my_end, cur);
                            }
#pragma unroll
                            for (int v = 0; v < V; v++) {
                                if (VALUED) P::fma(acc[v], a[u], b[u][v]);
                                else P::add(acc[v], b[u][v]);
                            }
                        }
                    }
                }
            }
        }
        for (; cur <= last; ++cur) {  // the last row with nonzeros, then trailing empty rows
            store_row(rb + cur, acc);
#pragma unroll
            for (int v = 0; v < V; v++) acc[v] = P::zero();
        }
    }
};

template <int V, bool VALUED, bool VEC4, int U>
__global__ void __launch_bounds__(kCta)
spmm_flat_kernel(int M, int K, long long total_keys, int task, const int *__restrict__ rowptr,
                 const int *__restrict__ colind, const float *__restrict__ val, const float *__restrict__ B,
                 long long ldb, float *__restrict__ C, long long ldc)
{
    using P = Pack<VEC4>;
    using T = typename P::T;
    constexpr int W = P::kWidth;

    __shared__ int s_long[kMaxLong];
    __shared__ int s_nlong;
    __shared__ T s_part[kWarps][V * 32];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_nlong = 0;
    __syncthreads();

    const int col0 = blockIdx.y * (32 * V * W) + lane * W;
    unsigned vmask = 0;
#pragma unroll
    for (int v = 0; v < V; v++)
        if (col0 + v * 32 * W < K) vmask |= 1u << v;

    Walker<V, VALUED, VEC4, U> wk;
    wk.colind = colind; wk.val = val;
    wk.Bl = B + col0; wk.Cl = C + col0;
    wk.ldb = ldb; wk.ldc = ldc; wk.vmask = vmask; wk.lane = lane;

    // ---- this warp's rows -------------------------------------------------------------------
    const long long t = (long long)blockIdx.x * kWarps + warp;
    const long long k0 = t * task;
    int row_lo = M, row_hi = M;
    {
        // all warps run the search (it contains warp collectives); out-of-range tasks get [M, M)
        const int shift = lane & 16;
        const long long target = k0 + (shift ? task : 0);
        const int r = search_key16(rowptr, M, target < total_keys ? target : total_keys + 1, lane & 15, shift);
        row_lo = __shfl_sync(kFull, r, 0);
        row_hi = __shfl_sync(kFull, r, 16);
        if (k0 >= total_keys) row_lo = row_hi = M;
    }

    for (int rb = row_lo; rb < row_hi; rb += 32) {
        const int nrows = min(32, row_hi - rb);
        int my_start = 0, my_end = 0;
        if (lane < nrows) {
            my_start = __ldg(rowptr + rb + lane);
            my_end = __ldg(rowptr + rb + lane + 1);
        }
        unsigned long_mask = __ballot_sync(kFull, my_end - my_start > kLongRow);
        int run = 0;
        while (true) {
            const int stop = long_mask ? (__ffs(long_mask) - 1) : nrows;  // next long row, or end of chunk
            if (stop > run) {
                const int s = __shfl_sync(kFull, my_start, run);
                const int e = __shfl_sync(kFull, my_end, stop - 1);
                T acc[V];
#pragma unroll
                for (int v = 0; v < V; v++) acc[v] = P::zero();
                wk.stream(s, e, acc, my_end, rb, run, stop - 1);
            }
            if (stop >= nrows) break;
            if (lane == 0) {
                const int slot = atomicAdd(&s_nlong, 1);
                if (slot < kMaxLong) s_long[slot] = rb + stop;
            }
            long_mask &= long_mask - 1;
            run = stop + 1;
        }
    }

    // ---- long rows: all warps of the CTA, contiguous segments, fixed-order combine ----------------
    __syncthreads();
    const int nlong = min(s_nlong, kMaxLong);
    for (int i = 0; i < nlong; i++) {
        const int r = s_long[i];
        const int a = __ldg(rowptr + r), b = __ldg(rowptr + r + 1);
        int seg = (b - a + kWarps - 1) / kWarps;
        seg = (seg + 31) & ~31;
        const int s = min(b, a + warp * seg), e = min(b, s + seg);
        T acc[V];
#pragma unroll
        for (int v = 0; v < V; v++) acc[v] = P::zero();
        wk.stream(s, e, acc, 0, 0, 1, 0);
#pragma unroll
        for (int v = 0; v < V; v++) s_part[warp][v * 32 + lane] = acc[v];
        __syncthreads();
        if (threadIdx.x < V * 32) {
            T sum = s_part[0][threadIdx.x];
#pragma unroll
            for (int w = 1; w < kWarps; w++) P::add(sum, s_part[w][threadIdx.x]);
            const int c = blockIdx.y * (32 * V * W) + threadIdx.x * W;
            if (c < K) P::stcs(C + (long long)r * ldc + c, sum);
        }
        __syncthreads();
    }
}

// K == 0 or M == 0 never reaches here.  nnz == 0 is handled by the same kernel (all rows empty).
template <int V, bool VALUED, bool VEC4>
cudaError_t launch(int M, int K, long long nnz, int task, const int *rowptr, const int *colind, const float *val,
                   const float *B, long long ldb, float *C, long long ldc, cudaStream_t st)
{
    constexpr int U = (V == 1) ? 8 : (V == 2 ? 4 : 2);
    constexpr int W = VEC4 ? 4 : 1;
    const long long total = nnz + M;
    const long long ntask = (total + task - 1) / task;
    dim3 grid((unsigned)((ntask + kWarps - 1) / kWarps), (unsigned)((K + 32 * V * W - 1) / (32 * V * W)), 1);
    spmm_flat_kernel<V, VALUED, VEC4, U><<<grid, kCta, 0, st>>>(M, K, total, task, rowptr, colind, val, B, ldb, C, ldc);
    return cudaGetLastError();
}

template <bool VALUED, bool VEC4>
cudaError_t launch_v(int V, int M, int K, long long nnz, int task, const int *rowptr, const int *colind,
                     const float *val, const float *B, long long ldb, float *C, long long ldc, cudaStream_t st)
{
    switch (V) {
        case 1: return launch<1, VALUED, VEC4>(M, K, nnz, task, rowptr, colind, val, B, ldb, C, ldc, st);
        case 2: return launch<2, VALUED, VEC4>(M, K, nnz, task, rowptr, colind, val, B, ldb, C, ldc, st);
        case 3: return launch<3, VALUED, VEC4>(M, K, nnz, task, rowptr, colind, val, B, ldb, C, ldc, st);
        default: return launch<4, VALUED, VEC4>(M, K, nnz, task, rowptr, colind, val, B, ldb, C, ldc, st);
    }
}

int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

}  // namespace

extern "C" int gespmm_csr_spmm_f32(int64_t M, int64_t N, int64_t K, int64_t nnz, const int32_t *rowptr,
                                   const int32_t *colind, const float *val, const float *B, int64_t ldb,
                                   float *C, int64_t ldc, void *stream)
{
    if (M < 0 || N < 0 || K < 0 || nnz < 0) return GESPMM_ERR_INVALID_ARG;
    if (M > INT32_MAX - 1 || N > INT32_MAX || nnz > INT32_MAX || K > INT32_MAX) return GESPMM_ERR_TOO_LARGE;
    if (M == 0 || K == 0) return GESPMM_OK;
    if (ldb < K || ldc < K) return GESPMM_ERR_INVALID_ARG;
    if (!rowptr || !C) return GESPMM_ERR_INVALID_ARG;
    if (nnz > 0 && (!colind || !B)) return GESPMM_ERR_INVALID_ARG;

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec4 = (K % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    const int W = vec4 ? 4 : 1;
    const int V = (int)((K + 32 * W - 1) / (32 * W)) >= 4 ? 4 : (int)((K + 32 * W - 1) / (32 * W));

    // Task window: kMaxTask keys, shrunk for small problems so the grid still covers the 148 SMs
    // several times over.
    static const int forced = env_int("GESPMM_TASK", 0);
    const long long total = nnz + M;
    int task = kMaxTask;
    const long long want = 148LL * 8 * kWarps;  // ~8 CTAs per SM
    if (total / task < want) {
        long long tk = total / want;
        tk = (tk + 31) & ~31LL;
        task = (int)(tk < 32 ? 32 : (tk > kMaxTask ? kMaxTask : tk));
    }
    if (forced >= 32 && forced <= kMaxTask) task = forced & ~31;

    cudaError_t err;
    if (val) {
        err = vec4 ? launch_v<true, true>(V, (int)M, (int)K, nnz, task, rowptr, colind, val, B, ldb, C, ldc, st)
                   : launch_v<true, false>(V, (int)M, (int)K, nnz, task, rowptr, colind, val, B, ldb, C, ldc, st);
    } else {
        err = vec4 ? launch_v<false, true>(V, (int)M, (int)K, nnz, task, rowptr, colind, val, B, ldb, C, ldc, st)
                   : launch_v<false, false>(V, (int)M, (int)K, nnz, task, rowptr, colind, val, B, ldb, C, ldc, st);
    }
    return err == cudaSuccess ? GESPMM_OK : GESPMM_ERR_CUDA;
}
