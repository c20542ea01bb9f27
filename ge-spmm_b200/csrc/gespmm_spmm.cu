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
//   A warp walks its rows' nonzeros [rowptr[row_lo], rowptr[row_hi]) in CSR order, 32 rows
//   (one rowptr register per lane) and 32 nonzeros (one colind/val register per lane, the next
//   32 prefetched) at a time.  (colind, val) are broadcast with __shfl and U independent
//   128-bit loads of B-row panels are issued before the first FMA, so a warp keeps U*V*512 B
//   of gathers in flight however short the rows are.  Row ends inside a 32-nonzero chunk are
//   one bit mask (__reduce_or_sync over the lanes' row ends): at a set bit the C row is stored
//   and the accumulators reset.  Each output element is accumulated in CSR order in one
//   register, FFMA for valued / FADD for unvalued, which is the reference's order -- results
//   are bit-identical to it.
//
// Gather ring (the default for aligned K >= 64)
//   In the register variant the number of B rows in flight per warp is bounded by the registers
//   that hold them.  The ring variant takes them out of the register file: every lane copies
//   its own 16-byte slice of each gathered B row into a per-warp shared-memory ring with
//   cp.async (LDGSTS, no register staging), G rows per commit group, NS groups deep, and reads
//   the same 16 bytes back with LDS.128 one group later -- lanes only ever read what they
//   copied themselves, so cp.async.wait_group is the only synchronisation.  In-flight gather
//   bytes per SM are then bounded by shared memory (up to ~190 KB), not by registers.
//
// Long rows
//   Rows with more than `long_row` (GESPMM_LONG_ROW) nonzeros are not walked by their owner
//   warp.  They are queued in shared memory and, after a CTA barrier, summed by all warps of
//   the CTA in contiguous segments whose partials are combined in fixed order through shared
//   memory (deterministic; fp32 re-association only).
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

constexpr int kMaxWarps = 8;    // warps per CTA is a template parameter NW <= kMaxWarps
constexpr int kMaxTask = 1024;  // largest task window (keys)
constexpr int kMinLong = 512;   // smallest accepted long-row threshold
constexpr int kMaxLong = kMaxWarps * kMaxTask / kMinLong + 2;  // long rows that can start in one CTA's windows
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

__device__ __forceinline__ unsigned low_bits(int n) { return n >= 32 ? kFull : ((1u << n) - 1u); }

// First r in [0, M) with rowptr[r] + r >= target, else M.  16 lanes cooperate; the two
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
    static constexpr int kStride = 32 * P::kWidth;  // floats between a lane's consecutive packs

    const int *__restrict__ colind;
    const float *__restrict__ val;
    const float *__restrict__ Bl;   // B + this lane's first owned column
    float *__restrict__ Cl;         // C + this lane's first owned column
    int ldb, ldc;
    unsigned vmask;                 // bit v set: this lane's v-th pack is inside K
    int lane;

    __device__ __forceinline__ void store_row(int row, const T (&acc)[V]) const {
        float *c = Cl + (long long)row * ldc;
#pragma unroll
        for (int v = 0; v < V; v++)
            if (vmask & (1u << v)) P::stcs(c + v * kStride, acc[v]);
    }

    // One batch of U nonzeros (chunk positions j0 .. j0+U-1): all loads first, then the FMAs in
    // CSR order with row flushes at the bits of `ends`.  FULL: every position is a real nonzero.
    template <bool FULL>
    __device__ __forceinline__ void batch(int mcol, float mval, int j0, unsigned live, unsigned ends, T (&acc)[V],
                                          unsigned &rows_left, int rb) const {
        T b[U][V];
        float a[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int c = __shfl_sync(kFull, mcol, j0 + u);
            if (VALUED) a[u] = __shfl_sync(kFull, mval, j0 + u);
            if (FULL || (live & (1u << u))) {
                const float *bp = Bl + (long long)c * ldb;
#pragma unroll
                for (int v = 0; v < V; v++)
                    if (vmask & (1u << v)) b[u][v] = P::ldg(bp + v * kStride);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (FULL || (live & (1u << u))) {
#pragma unroll
                for (int v = 0; v < V; v++) {
                    if (VALUED) P::fma(acc[v], a[u], b[u][v]);
                    else P::add(acc[v], b[u][v]);
                }
                if (ends & (1u << u)) {  // last nonzero of the current row
                    store_row(rb + __ffs(rows_left) - 1, acc);
                    rows_left &= rows_left - 1;
#pragma unroll
                    for (int v = 0; v < V; v++) acc[v] = P::zero();
                }
            }
        }
    }

    // Nonzeros [s, e) in CSR order.  `rows` = the non-empty rows (bits = row - rb) that END inside
    // [s, e), in order; `my_end` = lane's row end.  rows == 0: pure accumulation (segment mode).
    __device__ __forceinline__ void stream(int s, int e, T (&acc)[V], int my_end, unsigned rows, int rb) const {
        int ncol = 0;
        float nval = 1.f;
        if (s + lane < e) {
            ncol = __ldcs(colind + s + lane);
            if (VALUED) nval = __ldcs(val + s + lane);
        }
        const bool my_row = (rows >> lane) & 1u;
        unsigned rows_left = rows;
        for (int p0 = s; p0 < e; p0 += 32) {
            const int mcol = ncol;
            const float mval = nval;
            const int pn = p0 + 32 + lane;
            if (pn < e) {
                ncol = __ldcs(colind + pn);
                if (VALUED) nval = __ldcs(val + pn);
            }
            const unsigned rel = (unsigned)(my_end - 1 - p0);
            const unsigned endmask = __reduce_or_sync(kFull, (my_row && rel < 32u) ? (1u << rel) : 0u);
            const int n = min(32, e - p0);
            const unsigned livemask = n >= 32 ? kFull : ((1u << n) - 1u);
#pragma unroll 1
            for (int j0 = 0; j0 < n; j0 += U) {
                const unsigned live = livemask >> j0, ends = endmask >> j0;
                if ((live & ((1u << U) - 1u)) == ((1u << U) - 1u)) batch<true>(mcol, mval, j0, live, ends, acc, rows_left, rb);
                else batch<false>(mcol, mval, j0, live, ends, acc, rows_left, rb);
            }
        }
    }
};

// ---- cp.async helpers ---------------------------------------------------------------------------
template <int CP>
__device__ __forceinline__ void cp_async16(unsigned saddr, const float *g, unsigned long long policy)
{
    if (CP == 1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(g) : "memory");
    else if (CP == 2) asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(saddr), "l"(g), "l"(policy) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ unsigned long long l2_evict_last_policy()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds128(unsigned saddr)
{
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr) : "memory");
    return r;
}

// Ring variant of the walker (float4 packs only).  G rows per stage, NS stages, S32 = 32/G stages
// per 32-nonzero chunk; stage j of a chunk lives in ring slot j % NS (NS divides S32), and the
// copies for stage j + NS - 1 are issued right before stage j is consumed.
template <int V, bool VALUED, int G, int NS, int CP>
struct WalkerRing {
    using P = Pack<true>;
    using T = float4;
    static constexpr int kStride = 128;
    static constexpr int S32 = 32 / G;
    static constexpr int L = NS - 1;
    static constexpr int kStageBytes = G * V * 512;
    static constexpr int kRingBytes = NS * kStageBytes;  // per warp
    static_assert(32 % G == 0 && S32 % NS == 0 && (NS & (NS - 1)) == 0 && L >= 1 && L <= S32, "bad ring shape");

    const int *__restrict__ colind;
    const float *__restrict__ val;
    const float *__restrict__ Bl;
    float *__restrict__ Cl;
    int ldb, ldc;
    unsigned vmask;
    int lane;
    unsigned ring;  // shared-space address of this lane's 16 bytes in (slot 0, row 0, pack 0)
    unsigned long long policy;  // L2 eviction policy for the gathers (CP == 2)

    __device__ __forceinline__ void store_row(int row, const T (&acc)[V]) const {
        float *c = Cl + (long long)row * ldc;
#pragma unroll
        for (int v = 0; v < V; v++)
            if (vmask & (1u << v)) P::stcs(c + v * kStride, acc[v]);
    }

    // copies for the G nonzeros at chunk positions [pos0, pos0 + G) of a chunk holding n nonzeros,
    // into the stage at byte offset `slot` of the ring; always exactly one commit group
    __device__ __forceinline__ void issue(int cols, int pos0, int n, unsigned slot) const {
#pragma unroll
        for (int i = 0; i < G; i++) {
            const int c = __shfl_sync(kFull, cols, pos0 + i);
            if (pos0 + i < n) {
                const float *bp = Bl + (long long)c * ldb;
#pragma unroll
                for (int v = 0; v < V; v++)
                    if (vmask & (1u << v)) cp_async16<CP>(ring + slot + (i * V + v) * 512, bp + v * kStride, policy);
            }
        }
        cp_async_commit();
    }

    __device__ __forceinline__ void consume(float vals, int pos0, int n, unsigned endmask, T (&acc)[V], unsigned &rows_left,
                                            int rb, unsigned slot) const {
        constexpr int UB = G < 4 ? G : 4;
#pragma unroll
        for (int i0 = 0; i0 < G; i0 += UB) {
            T b[UB][V];
            float a[UB];
#pragma unroll
            for (int i = 0; i < UB; i++) {
                if (VALUED) a[i] = __shfl_sync(kFull, vals, pos0 + i0 + i);
                if (pos0 + i0 + i < n) {
#pragma unroll
                    for (int v = 0; v < V; v++)
                        if (vmask & (1u << v)) b[i][v] = lds128(ring + slot + ((i0 + i) * V + v) * 512);
                }
            }
#pragma unroll
            for (int i = 0; i < UB; i++) {
                const int pos = pos0 + i0 + i;
                if (pos < n) {
#pragma unroll
                    for (int v = 0; v < V; v++) {
                        if (VALUED) P::fma(acc[v], a[i], b[i][v]);
                        else P::add(acc[v], b[i][v]);
                    }
                    if ((endmask >> pos) & 1u) {
                        store_row(rb + __ffs(rows_left) - 1, acc);
                        rows_left &= rows_left - 1;
#pragma unroll
                        for (int v = 0; v < V; v++) acc[v] = P::zero();
                    }
                }
            }
        }
    }

    __device__ __forceinline__ void stream(int s, int e, T (&acc)[V], int my_end, unsigned rows, int rb) const {
        int ccol = 0, ncol = 0, fcol = 0;
        float cval = 1.f, nval = 1.f;
        if (s + lane < e) {
            ccol = __ldcs(colind + s + lane);
            if (VALUED) cval = __ldcs(val + s + lane);
        }
        if (s + 32 + lane < e) ncol = __ldcs(colind + s + 32 + lane);
        const bool my_row = (rows >> lane) & 1u;
        unsigned rows_left = rows;
#pragma unroll
        for (int j = 0; j < L; j++) issue(ccol, j * G, min(32, e - s), (j % NS) * kStageBytes);
#pragma unroll 1
        for (int p0 = s; p0 < e; p0 += 32) {
            if (p0 + 64 + lane < e) fcol = __ldcs(colind + p0 + 64 + lane);
            if (VALUED && p0 + 32 + lane < e) nval = __ldcs(val + p0 + 32 + lane);
            const unsigned rel = (unsigned)(my_end - 1 - p0);
            const unsigned endmask = __reduce_or_sync(kFull, (my_row && rel < 32u) ? (1u << rel) : 0u);
            const int n = min(32, e - p0);
            const int n_next = e - p0 - 32;
#pragma unroll 1
            for (int j = 0; j < S32; j++) {
                if (j * G >= n) break;  // only in the last chunk, where nothing is in flight past it
                const int jj = j + L;   // stage whose copies are issued now
                const bool nxt = jj >= S32;
                issue(nxt ? ncol : ccol, (jj & (S32 - 1)) * G, nxt ? n_next : n, (jj & (NS - 1)) * kStageBytes);
                cp_async_wait<L>();
                consume(cval, j * G, n, endmask, acc, rows_left, rb, (j & (NS - 1)) * kStageBytes);
            }
            ccol = ncol; ncol = fcol; cval = nval;
        }
        cp_async_wait<0>();
    }
};

template <class WK> struct RingBytes { static constexpr int value = 0; };
template <int V, bool VALUED, int G, int NS, int CP>
struct RingBytes<WalkerRing<V, VALUED, G, NS, CP>> { static constexpr int value = WalkerRing<V, VALUED, G, NS, CP>::kRingBytes; };

template <class WK, int V, bool VEC4, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
spmm_flat_kernel(int M, int K, long long total_keys, int task, int long_row, const int *__restrict__ rowptr,
                 const int *__restrict__ colind, const float *__restrict__ val, const float *__restrict__ B,
                 int ldb, float *__restrict__ C, int ldc)
{
    using P = Pack<VEC4>;
    using T = typename P::T;
    constexpr int W = P::kWidth;

    __shared__ int s_long[kMaxLong];
    __shared__ int s_nlong;
    __shared__ T s_part[NW][V * 32];
    extern __shared__ __align__(16) unsigned char s_dyn[];  // gather rings (ring variant only)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_nlong = 0;
    __syncthreads();

    const int col0 = blockIdx.y * (32 * V * W) + lane * W;
    unsigned vmask = 0;
#pragma unroll
    for (int v = 0; v < V; v++)
        if (col0 + v * 32 * W < K) vmask |= 1u << v;

    WK wk;
    wk.colind = colind; wk.val = val;
    wk.Bl = B + col0; wk.Cl = C + col0;
    wk.ldb = ldb; wk.ldc = ldc; wk.vmask = vmask; wk.lane = lane;
    if constexpr (RingBytes<WK>::value > 0) {
        wk.ring = (unsigned)__cvta_generic_to_shared(s_dyn) + warp * RingBytes<WK>::value + lane * 16;
        wk.policy = l2_evict_last_policy();
    }

    // ---- this warp's rows -------------------------------------------------------------------
    const long long t = (long long)blockIdx.x * NW + warp;
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
        const int len = my_end - my_start;
        unsigned long_mask = __ballot_sync(kFull, len > long_row);
        const unsigned nonempty = __ballot_sync(kFull, len > 0) & ~long_mask;
        // empty rows: zeros, one warp-wide store per row
        {
            unsigned em = ~(nonempty | long_mask) & low_bits(nrows);
            T z[V];
#pragma unroll
            for (int v = 0; v < V; v++) z[v] = P::zero();
            while (em) {
                wk.store_row(rb + __ffs(em) - 1, z);
                em &= em - 1;
            }
        }
        // runs of short rows between long rows, each walked as one flat stream
        int run = 0;
        while (true) {
            const int stop = long_mask ? (__ffs(long_mask) - 1) : nrows;  // next long row, or end of chunk
            const unsigned run_bits = low_bits(stop) & ~low_bits(run);
            const unsigned rows = nonempty & run_bits;
            if (rows) {
                const int s = __shfl_sync(kFull, my_start, __ffs(rows) - 1);
                const int e = __shfl_sync(kFull, my_end, 31 - __clz(rows));
                T acc[V];
#pragma unroll
                for (int v = 0; v < V; v++) acc[v] = P::zero();
                wk.stream(s, e, acc, my_end, rows, rb);
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
        int seg = (b - a + NW - 1) / NW;
        seg = (seg + 31) & ~31;
        const int s = min(b, a + warp * seg), e = min(b, s + seg);
        T acc[V];
#pragma unroll
        for (int v = 0; v < V; v++) acc[v] = P::zero();
        wk.stream(s, e, acc, 0, 0u, 0);
#pragma unroll
        for (int v = 0; v < V; v++) s_part[warp][v * 32 + lane] = acc[v];
        __syncthreads();
        for (int x = threadIdx.x; x < V * 32; x += NW * 32) {
            T sum = s_part[0][x];
#pragma unroll
            for (int w = 1; w < NW; w++) P::add(sum, s_part[w][x]);
            const int c = blockIdx.y * (32 * V * W) + x * W;
            if (c < K) P::stcs(C + (long long)r * ldc + c, sum);
        }
        __syncthreads();
    }
}

struct Args {
    int M, K, task, long_row, ldb, ldc;
    long long nnz;
    const int *rowptr, *colind;
    const float *val, *B;
    float *C;
    cudaStream_t st;
};

template <class WK, int V, bool VEC4, int NW, int MINB>
cudaError_t launch(const Args &a)
{
    constexpr int W = VEC4 ? 4 : 1;
    constexpr int dyn = RingBytes<WK>::value * NW;
    auto kern = spmm_flat_kernel<WK, V, VEC4, NW, MINB>;
    if (dyn > 0) {
        static cudaError_t attr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
        if (attr != cudaSuccess) return attr;
    }
    const long long total = a.nnz + a.M;
    const long long ntask = (total + a.task - 1) / a.task;
    dim3 grid((unsigned)((ntask + NW - 1) / NW), (unsigned)((a.K + 32 * V * W - 1) / (32 * V * W)), 1);
    kern<<<grid, NW * 32, dyn, a.st>>>(a.M, a.K, total, a.task, a.long_row, a.rowptr, a.colind, a.val, a.B, a.ldb, a.C, a.ldc);
    return cudaGetLastError();
}

template <int V, bool VALUED, bool VEC4, int U, int MINB>
cudaError_t launch_reg(const Args &a) { return launch<Walker<V, VALUED, VEC4, U>, V, VEC4, 8, MINB>(a); }

// CP: 0 = cp.async.cg, 1 = cp.async.ca (allocate in L1), 2 = cp.async.cg + L2 evict_last hint
template <int V, bool VALUED, int G, int NS, int CP, int NW, int MINB>
cudaError_t launch_ring(const Args &a) { return launch<WalkerRing<V, VALUED, G, NS, CP>, V, true, NW, MINB>(a); }

int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// (V, variant) -> instantiation.  Variants other than 0 exist for tuning (GESPMM_VARIANT).
template <bool VALUED, bool VEC4>
cudaError_t launch_v(int V, int variant, const Args &a)
{
    if constexpr (!VEC4) {  // scalar instantiations: correctness path for odd K / unaligned operands
        switch (V) {
            case 1: return launch_reg<1, VALUED, false, 8, 1>(a);
            case 2: return launch_reg<2, VALUED, false, 4, 1>(a);
            case 3: return launch_reg<3, VALUED, false, 4, 1>(a);
            default: return launch_reg<4, VALUED, false, 4, 1>(a);
        }
    } else {
        switch (V) {
            case 1:
                switch (variant) {
                    case 1: return launch_reg<1, VALUED, true, 8, 3>(a);
                    case 2: return launch_ring<1, VALUED, 8, 2, 0, 4, 6>(a);
                    case 3: return launch_ring<1, VALUED, 4, 2, 0, 8, 4>(a);
                    case 4: return launch_ring<1, VALUED, 4, 2, 0, 4, 8>(a);
                    case 5: return launch_ring<1, VALUED, 8, 2, 2, 8, 3>(a);
                    case 6: return launch_ring<1, VALUED, 8, 2, 2, 4, 6>(a);
                    case 7: return launch_ring<1, VALUED, 8, 2, 1, 8, 3>(a);
                    case 8: return launch_ring<1, VALUED, 16, 2, 0, 4, 3>(a);
                    default: return launch_ring<1, VALUED, 8, 2, 0, 8, 3>(a);
                }
            case 2:
                switch (variant) {
                    case 1: return launch_reg<2, VALUED, true, 4, 3>(a);
                    case 2: return launch_ring<2, VALUED, 4, 2, 0, 4, 6>(a);
                    case 3: return launch_ring<2, VALUED, 2, 2, 0, 8, 4>(a);
                    case 5: return launch_ring<2, VALUED, 4, 2, 2, 8, 3>(a);
                    default: return launch_ring<2, VALUED, 4, 2, 0, 8, 3>(a);
                }
            case 3:
                switch (variant) {
                    case 1: return launch_reg<3, VALUED, true, 2, 2>(a);
                    default: return launch_ring<3, VALUED, 2, 2, 0, 8, 3>(a);
                }
            default:
                switch (variant) {
                    case 1: return launch_reg<4, VALUED, true, 2, 2>(a);
                    case 2: return launch_ring<4, VALUED, 2, 2, 0, 4, 6>(a);
                    case 5: return launch_ring<4, VALUED, 2, 2, 2, 8, 3>(a);
                    default: return launch_ring<4, VALUED, 2, 2, 0, 8, 3>(a);
                }
        }
    }
}

}  // namespace

extern "C" int gespmm_csr_spmm_f32(int64_t M, int64_t N, int64_t K, int64_t nnz, const int32_t *rowptr,
                                   const int32_t *colind, const float *val, const float *B, int64_t ldb,
                                   float *C, int64_t ldc, void *stream)
{
    if (M < 0 || N < 0 || K < 0 || nnz < 0) return GESPMM_ERR_INVALID_ARG;
    if (M > INT32_MAX - 64 || N > INT32_MAX || nnz > INT32_MAX - 64 || K > INT32_MAX || ldb > INT32_MAX || ldc > INT32_MAX)
        return GESPMM_ERR_TOO_LARGE;
    if (M == 0 || K == 0) return GESPMM_OK;
    if (ldb < K || ldc < K) return GESPMM_ERR_INVALID_ARG;
    if (!rowptr || !C) return GESPMM_ERR_INVALID_ARG;
    if (nnz > 0 && (!colind || !B)) return GESPMM_ERR_INVALID_ARG;

    const bool vec4 = (K % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    const int W = vec4 ? 4 : 1;
    const int panels = (int)((K + 32 * W - 1) / (32 * W));
    const int V = panels >= 4 ? 4 : panels;

    // Task window (keys per warp): sized so the grid is ~40 waves of resident CTAs -- small enough
    // that the tail of the last wave is negligible, large enough (<= 512) that a task's start-up
    // (row search, first rowptr/colind fetch) is amortised.  Measured on B200: 128 is best for the
    // 20 M-key cit-Patents shape, 512 for the 100-200 M-key Reddit / products / R-MAT shapes.
    // GESPMM_TASK / GESPMM_LONG / GESPMM_VARIANT are tuning overrides (read per call).
    const int forced_task = env_int("GESPMM_TASK", 0);
    const int forced_long = env_int("GESPMM_LONG", 0);
    const int variant = env_int("GESPMM_VARIANT", 0);
    const long long total = nnz + M;
    const long long warps_per_wave = 148LL * 3 * kMaxWarps;
    long long tk = total / (40 * warps_per_wave);
    tk &= ~31LL;
    int task = (int)(tk < 32 ? 32 : (tk > 512 ? 512 : tk));
    if (forced_task >= 32 && forced_task <= kMaxTask) task = forced_task & ~31;
    int long_row = GESPMM_LONG_ROW;
    if (forced_long >= kMinLong) long_row = forced_long;

    Args a;
    a.M = (int)M; a.K = (int)K; a.task = task; a.long_row = long_row; a.ldb = (int)ldb; a.ldc = (int)ldc;
    a.nnz = nnz; a.rowptr = rowptr; a.colind = colind; a.val = val; a.B = B; a.C = C;
    a.st = static_cast<cudaStream_t>(stream);
    cudaError_t err;
    if (val) err = vec4 ? launch_v<true, true>(V, variant, a) : launch_v<true, false>(V, variant, a);
    else err = vec4 ? launch_v<false, true>(V, variant, a) : launch_v<false, false>(V, variant, a);
    return err == cudaSuccess ? GESPMM_OK : GESPMM_ERR_CUDA;
}
