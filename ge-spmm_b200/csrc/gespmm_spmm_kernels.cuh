// gespmm_spmm.cu -- fp32 CSR x dense SpMM for sm_100a, and its C-ABI launcher.
//
// Replaces the reference kernels topoSimple/topoCache/topoCacheCoarsenSPMMKernel and
// spmm_test0..4 (pytorch-custom/spmm_kernel.cu:31-173, 210-379; spmm_test.cu:64-454) and
// their launch blocks (spmm_kernel.cu:175-207, 425-458; spmm_test.cu:456-492).  Not a
// port: the reference maps (row, 64 columns) to a warp and walks the row serially with
// 32-bit loads; here the unit of work is a fixed-size slice of the merged (rows + nonzeros)
// sequence, walked as one flat stream of nonzeros whose B rows are gathered 512 bytes per
// warp instruction into a shared-memory ring.
//
// Work decomposition (kernel A, spmm_flat_kernel: one warp per CTA)
//   key(r) = rowptr[r] + r is strictly increasing, so the half-open key windows
//   [t*T, (t+1)*T) partition the rows: task t owns the rows whose key falls in its window.
//   Every task therefore costs at most T "row stores + nonzero gathers" plus the tail of its
//   last row, whatever the degree distribution (empty rows are work too: their C rows must
//   be zeroed).  A task finds its rows with a 16-ary search on rowptr (both window ends at
//   once, one per half-warp): no per-graph preprocessing, no workspace.  One task = one
//   32-thread CTA, so the hardware block scheduler is the dynamic load balancer and no warp
//   ever waits for another.
//
// Flat stream
//   A warp walks its rows' nonzeros [rowptr[row_lo], rowptr[row_hi]) in CSR order, 32 rows
//   (one rowptr register per lane) and 32 nonzeros (one colind/val register per lane, two
//   chunks prefetched) at a time.  Row ends inside a 32-nonzero chunk are one bit mask
//   (__reduce_or_sync over the lanes' row ends): at a set bit the C row is stored and the
//   accumulators reset.  Each output element is accumulated in CSR order in one register,
//   FFMA for valued / FADD for unvalued, which is the reference's order -- results are
//   bit-identical to it.
//
// Gather ring (aligned operands, K % 4 == 0)
//   Every lane copies its own 16-byte slice of each gathered B row into a per-warp
//   shared-memory ring with cp.async (LDGSTS: no register staging), G rows per commit group,
//   NS groups deep, and reads the same 16 bytes back with LDS.128 one group later -- lanes
//   only ever read what they copied themselves, so cp.async.wait_group is the only
//   synchronisation.  In-flight gather bytes per SM are bounded by shared memory (~100 KB
//   in flight out of ~200 KB of rings), not by registers.  Unaligned operands / K % 4 != 0
//   take the same walkers on 4-byte slices (cp.async.ca 4 + LDS.32; W = 1), or -- with scratch
//   memory from the caller -- padded copies of B and C on the 16-byte ones (run_spmm).
//   WalkerBulk is the same ring fed by one TMA bulk copy (cp.async.bulk + mbarrier) per row.
//
// Narrow B (K <= 64 in 16-byte slices, any K <= 16 in 4-byte slices)
//   A B row then fills only 32/NG lanes' 16-byte slices (NG = 2 / 4 / 8 for K <= 64 / 32 / 16), so the warp is
//   split into NG lane groups and every copy / read-back instruction serves NG nonzeros at once:
//   WalkerSub (default) deals consecutive nonzeros of the flat stream to the groups and adds the groups' partial
//   sums at each row end (re-associated, deterministic); WalkerRows (GESPMM_FLAG_SEQUENTIAL) deals whole rows to the
//   groups and keeps the sequential order.  1.1-4.8x / 1.2-2.1x the ring walker at these widths.
//   spmm_rowgroup_kernel: the sequential order for K <= 16 in 4-byte slices (every group sums its own row).
//
// Per-call options (gespmm_opts): summation order, the graph's longest row (skips kernel B), a per-gathered-row
//   scale, a per-stored-row scale and a bias fused into every walker (FUSE: GCNConv's element-wise passes, bit-identical
//   to running them separately), L2 eviction priorities (HINT), padding workspace.
//
// Reductions
//   sum (the hot path) or max (gespmm_csr_spmm_max_f32: the reference's DGL patch,
//   dgl-custom/binary_reduce_max.cu:18-168, `acc > x ? acc : x` from a caller-given start value).
//   Max is order-independent, so long rows are bit-identical to a sequential walk as well.
//
// Long rows (kernel B, spmm_long_kernel: 8 warps per CTA)
//   Rows with more than `long_row` nonzeros are skipped by kernel A.  Kernel B finds them
//   without a list: every thread probes one 256-aligned nonzero position, binary-searches
//   the row containing it, and claims that row if the row is long and the position is the
//   first aligned one inside it.  Claimed rows are summed by all 8 warps of the CTA in
//   contiguous segments whose partials are combined in fixed order through shared memory
//   (deterministic; differs from the reference by fp32 re-association only).  Rows of at least
//   32768 nonzeros (R-MAT hubs) are summed by the whole 8-CTA thread-block cluster -- 64
//   segments, CTA partials combined in rank order through distributed shared memory -- so the
//   longest row of the matrix does not become the tail of the launch.  Kernel B runs on a
//   helper stream, concurrently with kernel A.
//
// Sharded B (gespmm_csr_spmm_f32_bparts)
//   B may be given as up to 8 row blocks in different allocations -- the other GPUs' blocks mapped
//   through CUDA IPC.  The lane that loads a column resolves it to (block, local row) once and keeps
//   the row's byte address instead of the column; the gather then runs on the peer address over
//   NVLink.  Same walker, same order, same bits as with one contiguous B.
//
// Column mapping
//   Lane l owns, for v < V, the float4 at column ((v*32 + l) * 4) of the current panel
//   (panel = 128*V columns; blockIdx.y walks the panels one after the other).  One warp-wide
//   copy therefore moves 512 contiguous bytes of a B row.  V = 1 for B in local memory (a pass
//   over a 128-column panel gathers from an N x 512-byte slice of B: the smallest L2 footprint),
//   up to 4 for a sharded B.  K <= 64: see WalkerSub (several nonzeros per warp-wide copy).

#ifndef GESPMM_SPMM_KERNELS_CUH
#define GESPMM_SPMM_KERNELS_CUH

// Shared by the translation units of the product (gespmm_spmm*.cu): every walker, the three kernels and their launchers
// as templates; each .cu instantiates one family of them behind a plain function declared at the end of this header, so that
// the ~330 kernel instantiations compile in parallel (one translation unit took 4.5 minutes).

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <type_traits>

#include "gespmm.h"

namespace gespmm_detail {


constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxTask = 1024;     // largest task window (keys)
constexpr int kMinLong = 512;      // smallest accepted long-row threshold
constexpr int kLongWarps = 8;      // warps per CTA in the long-row kernel
constexpr int kProbeStride = 256;  // kernel B probes nonzero positions that are multiples of this (< kMinLong)
constexpr int kProbeWindow = kLongWarps * 32 * kProbeStride;  // nonzeros covered by one CTA of kernel B
constexpr int kMaxList = kProbeWindow / kMinLong + 2;         // long rows one CTA of kernel B can claim
constexpr int kClusterSize = 8;    // CTAs per cluster in kernel B (portable maximum)
constexpr int kHugeRow = 32768;    // rows at least this long are summed by a whole cluster (64 warps)
constexpr int kMaxHuge = kProbeWindow / kHugeRow + 2;         // huge rows one CTA of kernel B can claim

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
    static __device__ __forceinline__ T splat(float x) { return make_float4(x, x, x, x); }
    // the reference's max_reduce: acc > x ? acc : x (dgl-custom/binary_reduce_max.cu:18-20), NaN behaviour included
    static __device__ __forceinline__ void mx(T &acc, const T &b) {
        acc.x = acc.x > b.x ? acc.x : b.x; acc.y = acc.y > b.y ? acc.y : b.y;
        acc.z = acc.z > b.z ? acc.z : b.z; acc.w = acc.w > b.w ? acc.w : b.w;
    }
    static __device__ __forceinline__ T scaled(float a, const T &b) { return make_float4(a * b.x, a * b.y, a * b.z, a * b.w); }
    // separately rounded product / sum: never contracted into an FFMA (the fused GCN pro- and epilogue must give the
    // bits of the unfused element-wise passes, pytorch-custom/op.py:142-147)
    static __device__ __forceinline__ T mul_rn(const T &b, float s) {
        return make_float4(__fmul_rn(b.x, s), __fmul_rn(b.y, s), __fmul_rn(b.z, s), __fmul_rn(b.w, s));
    }
    static __device__ __forceinline__ void add_rn(T &acc, const T &b) {
        acc.x = __fadd_rn(acc.x, b.x); acc.y = __fadd_rn(acc.y, b.y); acc.z = __fadd_rn(acc.z, b.z); acc.w = __fadd_rn(acc.w, b.w);
    }
    static __device__ __forceinline__ T ldg_or_zero(const float *p, bool on) { return on ? ldg(p) : zero(); }
};
template <> struct Pack<false> {
    using T = float;
    static constexpr int kWidth = 1;
    static __device__ __forceinline__ T zero() { return 0.f; }
    static __device__ __forceinline__ T ldg(const float *p) { return __ldg(p); }
    static __device__ __forceinline__ void stcs(float *p, const T &a) { __stcs(p, a); }
    static __device__ __forceinline__ void fma(T &acc, float a, const T &b) { acc = fmaf(a, b, acc); }
    static __device__ __forceinline__ void add(T &acc, const T &b) { acc += b; }
    static __device__ __forceinline__ T splat(float x) { return x; }
    static __device__ __forceinline__ void mx(T &acc, const T &b) { acc = acc > b ? acc : b; }
    static __device__ __forceinline__ T scaled(float a, const T &b) { return a * b; }
    static __device__ __forceinline__ T mul_rn(const T &b, float s) { return __fmul_rn(b, s); }
    static __device__ __forceinline__ void add_rn(T &acc, const T &b) { acc = __fadd_rn(acc, b); }
    static __device__ __forceinline__ T ldg_or_zero(const float *p, bool on) { return on ? ldg(p) : zero(); }
};

// The reduction over a row: sum (FFMA / FADD, start 0) or max (start `init`, the reference's -10000 or -inf).
// FUSE (sum only): every gathered B row is first scaled by its column's col_scale (one rounded product), and a
// finished row is scaled by row_scale and offset by the bias on its way out (Epilogue below).
template <class P, bool VALUED, bool MAXR, bool FUSE = false>
struct Reduce {
    using T = typename P::T;
    static_assert(!(FUSE && MAXR), "the fused scaling belongs to the sum");
    static __device__ __forceinline__ T start(float init) { return MAXR ? P::splat(init) : P::zero(); }
    // SC: the gathered row is scaled (a fused walker whose call carries a col_scale); without it a fused walker's step
    // is the plain one -- the row scale and the bias cost nothing per nonzero
    template <bool SC = true>
    static __device__ __forceinline__ void step(T &acc, float a, const T &b, float s = 1.f) {
        if (FUSE && SC) {
            const T t = P::mul_rn(b, s);
            if (VALUED) P::fma(acc, a, t); else P::add_rn(acc, t);
        } else if (MAXR) { if (VALUED) P::mx(acc, P::scaled(a, b)); else P::mx(acc, b); }
        else { if (VALUED) P::fma(acc, a, b); else P::add(acc, b); }
    }
    static __device__ __forceinline__ void merge(T &acc, const T &other) {
        if (MAXR) P::mx(acc, other); else P::add(acc, other);
    }
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

constexpr int kMaxParts = 8;  // row blocks of a sharded B (one per GPU of an NVLink domain)

// B given as row blocks that live in different allocations -- typically one per GPU, mapped into this
// process through CUDA IPC and read over NVLink: block q holds rows [lo[q], lo[q+1]) at base[q].
struct PeerMap {
    const float *base[kMaxParts];
    int lo[kMaxParts + 1];
    int parts;  // 0: B is one array (Operands::B)
};

struct Operands {
    const int *colind;
    const float *val;
    const float *B;
    float *C;
    int ldb, ldc;
    float init;  // max-reduce: accumulator start and value of empty rows
    // fused GCN pro-/epilogue (FUSE walkers; each nullable): C[r,:] = (sum_p val[p] * (B[c_p,:] * col_scale[c_p])) * row_scale[r] + bias
    const float *row_scale, *col_scale, *bias;
    // L2 eviction-priority steering of the gathers (HINT walkers): priority code of "near" / "far" rows, the distance
    // |col - row| that separates them, and the priority of the C stores (codes: 0 normal, 1 evict_first, 2 evict_last, 3 unchanged)
    int l2_near, l2_far, l2_store, l2_window;
    const unsigned *l2_hot;  // bit c set: column c is "hot" (one of the most referenced rows of B): always near
    PeerMap peer;
};

// The row-end transformation of the fused walkers: acc * row_scale[row] + bias, two separately rounded operations.
template <class P, bool FUSE>
struct Epilogue {
    using T = typename P::T;
    static __device__ __forceinline__ T apply(const T &acc, float rs, const T &bias, bool has_bias) {
        if (!FUSE) return acc;
        T r = P::mul_rn(acc, rs);
        if (has_bias) P::add_rn(r, bias);
        return r;
    }
};

__device__ __forceinline__ unsigned long long l2_policy(int code)
{
    unsigned long long p;
    switch (code) {
        case 1: asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p)); break;
        case 2: asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p)); break;
        case 3: asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;\n" : "=l"(p)); break;
        default: asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;\n" : "=l"(p)); break;
    }
    return p;
}
__device__ __forceinline__ void st_hint_f4(float *p, const float4 &a, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;\n" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "l"(pol) : "memory");
}

// =================================================================================================
// Register walker: U B-row packs in flight per lane, in registers.  Any alignment, any K.
// =================================================================================================
template <int V, bool VALUED, bool VEC4, int U, bool MAXR = false, bool FUSE = false>
struct Walker {
    using P = Pack<VEC4>;
    using T = typename P::T;
    using R = Reduce<P, VALUED, MAXR, FUSE>;
    using E = Epilogue<P, FUSE>;
    static constexpr bool kFuse = FUSE;
    float init_v;
    __device__ __forceinline__ T start() const { return R::start(init_v); }
    static constexpr int kStride = 32 * P::kWidth;  // floats between a lane's consecutive packs
    static constexpr int kRingBytes = 0;

    const int *__restrict__ colind;
    const float *__restrict__ val;
    const float *__restrict__ Bl;   // B + this lane's first owned column
    float *__restrict__ Cl;         // C + this lane's first owned column
    int ldb, ldc;
    unsigned vmask;                 // bit v set: this lane's v-th pack is inside K
    int lane;
    // FUSE (see WalkerRing)
    const float *__restrict__ col_scale;
    const float *__restrict__ row_scale;
    float my_rs;
    T bias_v[V];
    bool has_bias;

    // panel = blockIdx.y: this warp owns columns [panel * kPanel, (panel + 1) * kPanel) of B and C
    static constexpr int kPanel = 32 * V * P::kWidth;
    __device__ __forceinline__ void init(const Operands &o, int panel, int K, int ln, unsigned /*ring*/) {
        const int col0 = panel * kPanel + ln * P::kWidth;
        vmask = 0;
#pragma unroll
        for (int v = 0; v < V; v++)
            if (col0 + v * kStride < K) vmask |= 1u << v;
        colind = o.colind; val = o.val; Bl = o.B + col0; Cl = o.C + col0; ldb = o.ldb; ldc = o.ldc; lane = ln;
        init_v = o.init;
        if constexpr (FUSE) {
            col_scale = o.col_scale; row_scale = o.row_scale; my_rs = 1.f;
            has_bias = o.bias != nullptr;
#pragma unroll
            for (int v = 0; v < V; v++) bias_v[v] = P::ldg_or_zero(o.bias + col0 + v * kStride, has_bias && (vmask & (1u << v)));
        }
    }
    __device__ __forceinline__ void finish(T (&)[V]) const {}  // a lane's accumulators are whole sums already
    __device__ __forceinline__ void load_rows(int rb, int nrows) {
        if constexpr (FUSE) my_rs = (row_scale && lane < nrows) ? __ldg(row_scale + rb + lane) : 1.f;
    }

    __device__ __forceinline__ void store_row(int rb, int rel, const T (&acc)[V]) const {
        float *c = Cl + (long long)(rb + rel) * ldc;
        float rs = 1.f;
        if constexpr (FUSE) rs = __shfl_sync(kFull, my_rs, rel);
#pragma unroll
        for (int v = 0; v < V; v++)
            if (vmask & (1u << v)) P::stcs(c + v * kStride, FUSE ? E::apply(acc[v], rs, bias_v[v], has_bias) : acc[v]);
    }

    template <bool FULL>
    __device__ __forceinline__ void batch(int mcol, float mval, float msc, int j0, unsigned live, unsigned ends, T (&acc)[V],
                                          unsigned &rows_left, int rb) const {
        T b[U][V];
        float a[U];
        [[maybe_unused]] float sc[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int c = __shfl_sync(kFull, mcol, j0 + u);
            if (VALUED) a[u] = __shfl_sync(kFull, mval, j0 + u);
            if (FUSE) sc[u] = __shfl_sync(kFull, msc, j0 + u);
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
                for (int v = 0; v < V; v++) R::step(acc[v], VALUED ? a[u] : 1.f, b[u][v], FUSE ? sc[u] : 1.f);
                if (ends & (1u << u)) {  // last nonzero of the current row
                    store_row(rb, __ffs(rows_left) - 1, acc);
                    rows_left &= rows_left - 1;
#pragma unroll
                    for (int v = 0; v < V; v++) acc[v] = start();
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
            float msc = 1.f;
            if constexpr (FUSE) msc = (col_scale && p0 + lane < e) ? __ldg(col_scale + mcol) : 1.f;
            const int pn = p0 + 32 + lane;
            if (pn < e) {
                ncol = __ldcs(colind + pn);
                if (VALUED) nval = __ldcs(val + pn);
            }
            const unsigned rel = (unsigned)(my_end - 1 - p0);
            const unsigned endmask = __reduce_or_sync(kFull, (my_row && rel < 32u) ? (1u << rel) : 0u);
            const int n = min(32, e - p0);
            const unsigned livemask = low_bits(n);
#pragma unroll 1
            for (int j0 = 0; j0 < n; j0 += U) {
                const unsigned live = livemask >> j0, ends = endmask >> j0;
                if ((live & ((1u << U) - 1u)) == ((1u << U) - 1u)) batch<true>(mcol, mval, msc, j0, live, ends, acc, rows_left, rb);
                else batch<false>(mcol, mval, msc, j0, live, ends, acc, rows_left, rb);
            }
        }
    }
};

// =================================================================================================
// Ring walker: B rows gathered into shared memory with cp.async.  float4 packs only.
// =================================================================================================
// CP: 0 = cp.async.cg (L2 only), 1 = cp.async.ca (allocate in L1).  (An L2 evict_last cache hint on
// the gathers was measured too: no gain on any shape, dropped.)
template <int CP>
__device__ __forceinline__ void cp_async16(unsigned saddr, const void *g)
{
    if (CP == 1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(g) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(g) : "memory");
}
// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier ---------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
                 "@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(unsigned dst, const void *src, unsigned bytes, unsigned bar, unsigned long long pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async4(unsigned saddr, const void *g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ float lds32(unsigned saddr)
{
    float r;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(r) : "r"(saddr) : "memory");
    return r;
}
__device__ __forceinline__ float4 lds128(unsigned saddr)
{
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr) : "memory");
    return r;
}

// G rows per stage, NS stages (power of two, divides S32 = 32/G).  Stage j of a 32-nonzero chunk
// lives in ring slot j % NS; the copies for stage j + NS - 1 are issued right before stage j is
// consumed.  MASKED: some lanes' packs lie beyond K (K is not a multiple of 128*V).
// FUSE: per-column scale on the gathered rows, per-row scale + bias on the stored rows (Operands).
// HINT: every gather carries an L2 eviction priority chosen by the lane that loaded the column from the distance
//       between the column and the rows being summed (bit 31 of the token = "far"); C stores carry one too.
// W = floats per lane and pack: 4 (16-byte slices: K % 4 == 0, aligned operands) or 1 (4-byte slices: any K, any 4-byte
//     alignment; cp.async.ca, the only 4-byte form) -- a pack is then 128 columns / 512 bytes or 32 columns / 128 bytes.
// SCALE (fused walkers): the call carries a col_scale; without it (row scale / bias only) nothing is added per nonzero.
template <int V, bool VALUED, int G, int NS, int CP, bool MASKED, bool PEER = false, bool MAXR = false, bool FUSE = false,
          bool HINT = false, int W = 4, bool SCALE = FUSE>
struct WalkerRing {
    static_assert(FUSE || !SCALE, "the gathered-row scale belongs to the fused walkers");
    static_assert(W == 4 || (W == 1 && !PEER && !HINT), "4-byte slices: local B, no L2 hints");
    using R = Reduce<Pack<W == 4>, VALUED, MAXR, FUSE>;
    using E = Epilogue<Pack<W == 4>, FUSE>;
    static constexpr bool kFuse = FUSE;
    static_assert(!(HINT && PEER), "the L2 hints are for local B");
    float init_v;
    __device__ __forceinline__ typename Pack<W == 4>::T start() const { return R::start(init_v); }
    // what a lane keeps per prefetched nonzero: its column (B is one array) or the byte address of its
    // B row (B is a set of row blocks; the owner lookup is done once, by the lane that loaded the column)
    using Tok = typename std::conditional<PEER, unsigned long long, int>::type;
    using P = Pack<W == 4>;
    using T = typename P::T;
    static constexpr int kStride = 32 * W;          // floats between a lane's consecutive packs
    static constexpr int kPackBytes = 32 * 4 * W;   // one warp-wide copy: 512 or 128 bytes of a B row
    static constexpr int S32 = 32 / G;
    static constexpr int L = NS - 1;
    static constexpr int UB = G < 4 ? G : 4;  // rows read back from the ring per LDS batch
    static constexpr int kStageBytes = G * V * kPackBytes;
    static constexpr int kRingBytes = NS * kStageBytes;  // per warp
    static_assert(32 % G == 0 && S32 % NS == 0 && (NS & (NS - 1)) == 0 && L >= 1 && L <= S32 && G % UB == 0, "bad ring shape");

    const int *__restrict__ colind;
    const float *__restrict__ val;
    const char *__restrict__ Bl;  // B + this lane's first owned column (byte pointer)
    float *__restrict__ Cl;
    unsigned ldb_bytes;
    int ldc;
    unsigned vmask;
    int lane;
    unsigned ring;              // shared-space address of this lane's 16 bytes in (slot 0, row 0, pack 0)
    const PeerMap *peer;        // PEER: the row blocks of B (in the kernel's parameter space)
    unsigned lane_off;          // PEER: byte offset of this lane's first owned column inside a B row
    // FUSE
    const float *__restrict__ col_scale;
    const float *__restrict__ row_scale;
    float my_rs;                // row_scale of row (rb + lane) of the current 32-row batch
    T bias_v[V];
    bool has_bias;
    // HINT
    unsigned long long pol_near, pol_far, pol_store;
    int window, cur_row;
    bool hint_store;            // store priority 0 = the plain streaming store

    static constexpr int kPanel = kStride * V;
    __device__ __forceinline__ void init(const Operands &o, int panel, int K, int ln, unsigned ring_base) {
        const int col0 = panel * kPanel + ln * W;
        vmask = 0;
#pragma unroll
        for (int v = 0; v < V; v++)
            if (col0 + v * kStride < K) vmask |= 1u << v;
        colind = o.colind; val = o.val; Bl = reinterpret_cast<const char *>(o.B + col0); Cl = o.C + col0;
        ldb_bytes = (unsigned)o.ldb * 4u; ldc = o.ldc; lane = ln;
        ring = ring_base + ln * (4 * W);
        peer = &o.peer; lane_off = (unsigned)col0 * 4u;
        init_v = o.init;
        if constexpr (FUSE) {
            col_scale = o.col_scale; row_scale = o.row_scale; my_rs = 1.f;
            has_bias = o.bias != nullptr;
#pragma unroll
            for (int v = 0; v < V; v++) bias_v[v] = P::ldg_or_zero(o.bias + col0 + v * kStride, has_bias && (vmask & (1u << v)));
        }
        if constexpr (HINT) {
            pol_near = l2_policy(o.l2_near); pol_far = l2_policy(o.l2_far); pol_store = l2_policy(o.l2_store);
            window = o.l2_window; cur_row = 0;
            hint_store = o.l2_store != 0;
        }
    }
    __device__ __forceinline__ void finish(T (&)[V]) const {}

    // FUSE: the row scales of the 32-row batch starting at rb (lane i: row rb + i), fetched next to its rowptr
    __device__ __forceinline__ void load_rows(int rb, int nrows) {
        if constexpr (FUSE) my_rs = (row_scale && lane < nrows) ? __ldg(row_scale + rb + lane) : 1.f;
    }

    __device__ __forceinline__ Tok load_tok(int p) const {
        const int c = __ldcs(colind + p);
        if constexpr (PEER) {
            int q = 0;
#pragma unroll
            for (int i = 1; i < kMaxParts; i++) q += (i < peer->parts && c >= peer->lo[i]) ? 1 : 0;
            return (unsigned long long)reinterpret_cast<uintptr_t>(peer->base[q]) +
                   (unsigned long long)(unsigned)(c - peer->lo[q]) * ldb_bytes;
        } else {
            return c;
        }
    }
    // FUSE: the scale of the column a token names (1 when there is no col_scale)
    __device__ __forceinline__ float load_scale(Tok t, bool on) const {
        if constexpr (SCALE && !PEER) return (on && col_scale) ? __ldg(col_scale + t) : 1.f;
        else return 1.f;
    }

    __device__ __forceinline__ bool pack_on(int v) const { return !MASKED || (vmask & (1u << v)); }

    // row (rb + rel) of the current batch leaves: the fused walkers scale it and add the bias on the way out
    __device__ __forceinline__ void store_row(int rb, int rel, const T (&acc)[V]) const {
        float *c = Cl + (long long)(rb + rel) * ldc;
        float rs = 1.f;
        if constexpr (FUSE) rs = __shfl_sync(kFull, my_rs, rel);
#pragma unroll
        for (int v = 0; v < V; v++) {
            if (pack_on(v)) {
                const T out = FUSE ? E::apply(acc[v], rs, bias_v[v], has_bias) : acc[v];
                if constexpr (HINT && W == 4) {
                    if (hint_store) st_hint_f4(c + v * kStride, out, pol_store);
                    else P::stcs(c + v * kStride, out);
                } else P::stcs(c + v * kStride, out);
            }
        }
    }

    // copies for the G nonzeros at chunk positions [pos0, pos0 + G) of a chunk holding n nonzeros,
    // into the stage at byte offset `slot` of the ring; always exactly one commit group
    template <bool FULL>
    __device__ __forceinline__ void issue_impl(Tok cols, int pos0, int n, unsigned slot) const {
        // addresses first, then the copies back to back (ptxas pads every LDGSTS that follows other
        // work with three dummy LDS; consecutive LDGSTS share one such pad)
#pragma unroll
        for (int i0 = 0; i0 < G; i0 += UB) {
            const char *bp[UB];
#pragma unroll
            for (int i = 0; i < UB; i++) {
                const Tok t = __shfl_sync(kFull, cols, pos0 + i0 + i);
                if constexpr (PEER) bp[i] = reinterpret_cast<const char *>((uintptr_t)(t + lane_off));
                else bp[i] = Bl + (unsigned long long)(unsigned)t * ldb_bytes;
            }
#pragma unroll
            for (int i = 0; i < UB; i++) {
                if (FULL || pos0 + i0 + i < n) {
#pragma unroll
                    for (int v = 0; v < V; v++) {
                        if (pack_on(v)) {
                            // (cp.async with an L2 cache-hint operand assembles -- LDGSTS with a policy descriptor -- but
                            // traps as an illegal instruction on sm_100a, profiles/r02_ldgsts_cache_hint_illegal_instruction.txt:
                            // per-gather priorities are the bulk walker's business)
                            if constexpr (W == 4) cp_async16<CP>(ring + slot + ((i0 + i) * V + v) * kPackBytes, bp[i] + v * kPackBytes);
                            else cp_async4(ring + slot + ((i0 + i) * V + v) * kPackBytes, bp[i] + v * kPackBytes);
                        }
                    }
                }
            }
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void issue(Tok cols, int pos0, int n, unsigned slot) const {
        if (pos0 + G <= n) issue_impl<true>(cols, pos0, n, slot);
        else issue_impl<false>(cols, pos0, n, slot);
    }

    __device__ __forceinline__ void flush(T (&acc)[V], unsigned &rows_left, int rb) const {
        store_row(rb, __ffs(rows_left) - 1, acc);
        rows_left &= rows_left - 1;
#pragma unroll
        for (int v = 0; v < V; v++) acc[v] = start();
    }

    template <bool FULL, bool SC>
    __device__ __forceinline__ void consume_impl(float vals, float scales, int pos0, int n, unsigned endmask, T (&acc)[V],
                                                 unsigned &rows_left, int rb, unsigned slot) const {
#pragma unroll
        for (int i0 = 0; i0 < G; i0 += UB) {
            T b[UB][V];
            float a[UB];
            [[maybe_unused]] float sc[UB];
#pragma unroll
            for (int i = 0; i < UB; i++) {
                if (VALUED) a[i] = __shfl_sync(kFull, vals, pos0 + i0 + i);
                if (SC) sc[i] = __shfl_sync(kFull, scales, pos0 + i0 + i);
                if (FULL || pos0 + i0 + i < n) {
#pragma unroll
                    for (int v = 0; v < V; v++)
                        if (pack_on(v)) {
                            if constexpr (W == 4) b[i][v] = lds128(ring + slot + ((i0 + i) * V + v) * kPackBytes);
                            else b[i][v] = lds32(ring + slot + ((i0 + i) * V + v) * kPackBytes);
                        }
                }
            }
            const unsigned ends = (endmask >> (pos0 + i0)) & ((1u << UB) - 1u);
            if (FULL && ends == 0u) {  // no row ends among these UB nonzeros: straight accumulation
#pragma unroll
                for (int i = 0; i < UB; i++) {
#pragma unroll
                    for (int v = 0; v < V; v++) R::template step<SC>(acc[v], VALUED ? a[i] : 1.f, b[i][v], SC ? sc[i] : 1.f);
                }
            } else {
#pragma unroll
                for (int i = 0; i < UB; i++) {
                    if (FULL || pos0 + i0 + i < n) {
#pragma unroll
                        for (int v = 0; v < V; v++) R::template step<SC>(acc[v], VALUED ? a[i] : 1.f, b[i][v], SC ? sc[i] : 1.f);
                        if (ends & (1u << i)) flush(acc, rows_left, rb);
                    }
                }
            }
        }
    }
    __device__ __forceinline__ void consume(float vals, float scales, int pos0, int n, unsigned endmask, T (&acc)[V],
                                            unsigned &rows_left, int rb, unsigned slot) const {
        if (pos0 + G <= n) consume_impl<true, SCALE>(vals, scales, pos0, n, endmask, acc, rows_left, rb, slot);
        else consume_impl<false, SCALE>(vals, scales, pos0, n, endmask, acc, rows_left, rb, slot);
    }

    __device__ __forceinline__ void stream(int s, int e, T (&acc)[V], int my_end, unsigned rows, int rb) {
        if constexpr (HINT) cur_row = rb;
        Tok ccol = 0, ncol = 0, fcol = 0;
        float cval = 1.f, nval = 1.f;
        [[maybe_unused]] float csc = 1.f, nsc = 1.f;
        if (s + lane < e) {
            ccol = load_tok(s + lane);
            if (VALUED) cval = __ldcs(val + s + lane);
        }
        if (s + 32 + lane < e) ncol = load_tok(s + 32 + lane);
        if constexpr (SCALE) csc = load_scale(ccol, s + lane < e);
        const bool my_row = (rows >> lane) & 1u;
        unsigned rows_left = rows;
#pragma unroll
        for (int j = 0; j < L; j++) issue(ccol, j * G, min(32, e - s), (j % NS) * kStageBytes);
#pragma unroll 1
        for (int p0 = s; p0 < e; p0 += 32) {
            if (p0 + 64 + lane < e) fcol = load_tok(p0 + 64 + lane);
            if (VALUED && p0 + 32 + lane < e) nval = __ldcs(val + p0 + 32 + lane);
            if constexpr (SCALE) nsc = load_scale(ncol, p0 + 32 + lane < e);  // ncol arrived one iteration ago
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
                consume(cval, csc, j * G, n, endmask, acc, rows_left, rb, (j & (NS - 1)) * kStageBytes);
            }
            ccol = ncol; ncol = fcol; cval = nval;
            if constexpr (SCALE) csc = nsc;
        }
        cp_async_wait<0>();
    }
};

// =================================================================================================
// Bulk walker: every gathered B row is ONE TMA bulk copy (cp.async.bulk, SASS UBLKCP) into the ring.
// =================================================================================================
// Same flat stream, same ring geometry (8 rows per stage, two stages, 512 bytes per row and 128-column panel) and same
// consume side as WalkerRing, but the copy side is the Blackwell/Hopper asynchronous proxy:
//   * the lane that LOADED a nonzero's column issues that row's copy itself -- one 512-byte cp.async.bulk per row, no
//     shuffle, no per-lane 16-byte LDGSTS through the LSU data pipe (every gathered byte used to cross shared memory's
//     store path once via LDGSTS and once via LDS; now only the LDS remains);
//   * completion is an mbarrier per stage slot: lane 0 arrives with the stage's byte count (expect_tx), the copies
//     complete_tx it, every lane waits on the slot's phase parity before its LDS.128;
//   * a bulk copy takes an L2 cache-policy operand (HINT): rows whose column lies within `window` rows of the rows being
//     summed are fetched evict_normal / evict_last, the others evict_first, so that the reuse window of a banded graph
//     is not pushed out of the L2 by rows that will not be referenced again (LDGSTS cannot carry the operand on sm_100a).
// Sum only, B in one array, K % 4 == 0 and aligned operands (a bulk copy needs 16-byte aligned addresses and sizes).
template <bool VALUED, bool MASKED, bool HINT>
struct WalkerBulk {
    using P = Pack<true>;
    using T = float4;
    using R = Reduce<P, VALUED, false, false>;
    static constexpr bool kFuse = false;
    static constexpr int G = 8, NS = 2, S32 = 32 / G, UB = 4;
    static constexpr int kStageBytes = G * 512;
    static constexpr int kRingBytes = NS * kStageBytes + 16;  // per warp: the ring + one mbarrier per slot
    static constexpr int kPanel = 128;
    static_assert((S32 & 1) == 0, "stage s of every chunk must land in slot s & 1");

    const int *__restrict__ colind;
    const float *__restrict__ val;
    const char *__restrict__ Bp;  // B + the panel's first column (byte pointer; the same for every lane)
    float *__restrict__ Cl;       // C + this lane's 4 columns
    unsigned ldb_bytes, row_bytes;
    int ldc, lane;
    bool on;                      // this lane's 4 columns lie inside K
    unsigned ring, ring_lane, bar;
    unsigned phase;               // bit s: the parity the next wait on slot s expects
    int slot_out;                 // slot of the one stage that is issued but not yet waited for
    unsigned long long pol_near, pol_far, pol_store;
    int window, cur_row;
    bool hint_store;
    const unsigned *__restrict__ hot;

    __device__ __forceinline__ T start() const { return P::zero(); }

    __device__ __forceinline__ void init(const Operands &o, int panel, int K, int ln, unsigned ring_base) {
        const int col0 = panel * kPanel + ln * 4;
        on = col0 < K;
        colind = o.colind; val = o.val;
        Bp = reinterpret_cast<const char *>(o.B + panel * kPanel);
        Cl = o.C + col0;
        ldb_bytes = (unsigned)o.ldb * 4u; ldc = o.ldc; lane = ln;
        row_bytes = (unsigned)min(kPanel, K - panel * kPanel) * 4u;
        ring = ring_base; ring_lane = ring_base + ln * 16; bar = ring_base + NS * kStageBytes;
        phase = 0; slot_out = 0;
        if (ln == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); }
        mbar_init_fence();
        __syncwarp();
        if constexpr (HINT) {
            pol_near = l2_policy(o.l2_near); pol_far = l2_policy(o.l2_far); pol_store = l2_policy(o.l2_store);
            window = o.l2_window; cur_row = 0; hint_store = o.l2_store != 0; hot = o.l2_hot;
        }
    }
    __device__ __forceinline__ void finish(T (&)[1]) const {}
    __device__ __forceinline__ void load_rows(int, int) {}

    // a lane's token for the nonzero it loaded: the column, bit 31 set when the row is "far" (HINT)
    __device__ __forceinline__ int load_tok(int p) const {
        const int c = __ldcs(colind + p);
        if constexpr (HINT) {
            // near = worth keeping in the L2: a hot column (the plan's bitmap of the most referenced rows of B), or a
            // column within `window` rows of the rows being summed (banded graphs)
            const int d = c - cur_row;
            bool near = (d < 0 ? -d : d) <= window;
            if (hot) near = near || ((__ldg(hot + ((unsigned)c >> 5)) >> (c & 31)) & 1u);
            return near ? c : (c | (int)0x80000000);
        } else {
            return c;
        }
    }

    __device__ __forceinline__ void store_row(int rb, int rel, const T (&acc)[1]) const {
        float *c = Cl + (long long)(rb + rel) * ldc;
        if (!MASKED || on) {
            if constexpr (HINT) {
                if (hint_store) st_hint_f4(c, acc[0], pol_store);
                else P::stcs(c, acc[0]);
            } else P::stcs(c, acc[0]);
        }
    }

    // the copies of the <= G nonzeros at chunk positions [pos0, pos0 + G) of a chunk holding n nonzeros (`toks`: lane i
    // holds nonzero i of that chunk) into ring slot `slot`; always exactly one arrival on the slot's mbarrier
    __device__ __forceinline__ void issue(int toks, int pos0, int n, int slot) {
        const int cnt = max(0, min(G, n - pos0));
        __syncwarp();  // every lane is done reading what the slot held (the copies below overwrite it)
        const unsigned b = bar + 8 * slot;
        if (lane == 0) mbar_arrive_tx(b, (unsigned)cnt * row_bytes);
        const int my = lane - pos0;
        const bool mine = my >= 0 && my < cnt;
        const char *src = Bp + (unsigned long long)((unsigned)toks & 0x7fffffffu) * ldb_bytes;
        const unsigned dst = ring + slot * kStageBytes + my * 512;
        if constexpr (HINT) {  // the policy operand is warp-uniform per instruction: one instruction per priority class
            if (mine && toks < 0) bulk_g2s_hint(dst, src, row_bytes, b, pol_far);
            if (mine && toks >= 0) bulk_g2s_hint(dst, src, row_bytes, b, pol_near);
        } else {
            if (mine) bulk_g2s(dst, src, row_bytes, b);
        }
        slot_out = slot;
    }
    __device__ __forceinline__ void wait(int slot) {
        mbar_wait(bar + 8 * slot, (phase >> slot) & 1u);
        phase ^= 1u << slot;
    }

    __device__ __forceinline__ void flush(T (&acc)[1], unsigned &rows_left, int rb) const {
        store_row(rb, __ffs(rows_left) - 1, acc);
        rows_left &= rows_left - 1;
        acc[0] = start();
    }

    template <bool FULL>
    __device__ __forceinline__ void consume_impl(float vals, int pos0, int n, unsigned endmask, T (&acc)[1], unsigned &rows_left,
                                                 int rb, int slot) const {
#pragma unroll
        for (int i0 = 0; i0 < G; i0 += UB) {
            T b[UB];
            float a[UB];
#pragma unroll
            for (int i = 0; i < UB; i++) {
                if (VALUED) a[i] = __shfl_sync(kFull, vals, pos0 + i0 + i);
                if ((FULL || pos0 + i0 + i < n) && (!MASKED || on)) b[i] = lds128(ring_lane + slot * kStageBytes + (i0 + i) * 512);
            }
            const unsigned ends = (endmask >> (pos0 + i0)) & ((1u << UB) - 1u);
            if (FULL && ends == 0u) {
#pragma unroll
                for (int i = 0; i < UB; i++) R::step(acc[0], VALUED ? a[i] : 1.f, b[i]);
            } else {
#pragma unroll
                for (int i = 0; i < UB; i++) {
                    if (FULL || pos0 + i0 + i < n) {
                        R::step(acc[0], VALUED ? a[i] : 1.f, b[i]);
                        if (ends & (1u << i)) flush(acc, rows_left, rb);
                    }
                }
            }
        }
    }
    __device__ __forceinline__ void consume(float vals, int pos0, int n, unsigned endmask, T (&acc)[1], unsigned &rows_left,
                                            int rb, int slot) const {
        if (pos0 + G <= n) consume_impl<true>(vals, pos0, n, endmask, acc, rows_left, rb, slot);
        else consume_impl<false>(vals, pos0, n, endmask, acc, rows_left, rb, slot);
    }

    // same contract as WalkerRing::stream
    __device__ __forceinline__ void stream(int s, int e, T (&acc)[1], int my_end, unsigned rows, int rb) {
        if constexpr (HINT) cur_row = rb;
        int ccol = 0, ncol = 0, fcol = 0;
        float cval = 1.f, nval = 1.f;
        if (s + lane < e) {
            ccol = load_tok(s + lane);
            if (VALUED) cval = __ldcs(val + s + lane);
        }
        if (s + 32 + lane < e) ncol = load_tok(s + 32 + lane);
        const bool my_row = (rows >> lane) & 1u;
        unsigned rows_left = rows;
        issue(ccol, 0, min(32, e - s), 0);
#pragma unroll 1
        for (int p0 = s; p0 < e; p0 += 32) {
            if (p0 + 64 + lane < e) fcol = load_tok(p0 + 64 + lane);
            if (VALUED && p0 + 32 + lane < e) nval = __ldcs(val + p0 + 32 + lane);
            const unsigned rel = (unsigned)(my_end - 1 - p0);
            const unsigned endmask = __reduce_or_sync(kFull, (my_row && rel < 32u) ? (1u << rel) : 0u);
            const int n = min(32, e - p0);
            const int n_next = e - p0 - 32;
#pragma unroll 1
            for (int j = 0; j < S32; j++) {
                if (j * G >= n) break;  // only in the last chunk; the stage issued last is empty
                const int jj = j + 1;   // stage whose copies are issued now
                const bool nxt = jj >= S32;
                issue(nxt ? ncol : ccol, (jj & (S32 - 1)) * G, nxt ? n_next : n, jj & 1);
                wait(j & 1);
                consume(cval, j * G, n, endmask, acc, rows_left, rb, j & 1);
            }
            ccol = ncol; ncol = fcol; cval = nval;
        }
        wait(slot_out);  // the (empty) stage issued last: every arrival has been waited for when the stream returns
    }
};

// =================================================================================================
// Sub-warp ring walker for narrow B (K <= 64): NG nonzeros per warp-wide copy.
// =================================================================================================
// A B row of K <= 64 floats fills only LPR = 32 / NG lanes' float4 slices, so the ring walker above
// spends a whole warp instruction (copy, read-back, shuffle, address) on 128-256 bytes.  Here the
// warp is NG groups of LPR lanes and one step handles a QUAD of NG consecutive nonzeros of the flat
// stream: group g copies / reads / accumulates nonzero (quad * NG + g), so every gather instruction
// moves 512 bytes again.  Each group keeps its own partial sum of the current row (the nonzeros whose
// stream position is g mod NG); at a row end the NG partials are added in a fixed butterfly order
// (group g with g ^ NG/2, then ^ NG/4, ...) and group 0 stores the row.  Deterministic, but NOT the
// reference's strictly sequential order: results differ from it by fp32 re-association (max-reduce:
// still bit-identical).  gespmm_row_sum_is_sequential() tells callers which rows that applies to.
// W = floats per lane: 4 (16-byte slices: K % 4 == 0, aligned operands) or 1 (4-byte slices: ANY K <= 32 / NG and any
// 4-byte alignment -- the class-count widths of a GCN's last layer, K = 3, 7, ...; cp.async.ca, the only 4-byte form).
template <int NG, bool VALUED, bool MAXR = false, bool FUSE = false, int W = 4, bool SCALE = FUSE>
struct WalkerSub {
    using P = Pack<W == 4>;
    using T = typename P::T;
    static_assert(W == 4 || W == 1, "16-byte or 4-byte slices");
    using R = Reduce<P, VALUED, MAXR, FUSE>;
    using E = Epilogue<P, FUSE>;
    static constexpr bool kFuse = FUSE;
    static_assert(NG == 2 || NG == 4 || NG == 8, "2, 4 or 8 B rows per warp-wide copy");
    static constexpr int LPR = 32 / NG;             // lanes per B row, one float4 each
    static constexpr int Q = 32 / NG;               // quads per 32-nonzero chunk
    static constexpr int QS = Q < 8 ? Q : 8;        // quads per ring stage
    static constexpr int SN = QS * NG;              // nonzeros per stage
    static constexpr int SPC = Q / QS;              // stages per chunk (1 or 2)
    static constexpr int UB = 4;                    // quads read back from the ring per LDS batch
    static constexpr int kRowBytes = 32 * 4 * W;    // one warp-wide copy (NG B rows) in the ring
    static constexpr int kStageBytes = QS * kRowBytes;
    static constexpr int kRingBytes = 2 * kStageBytes;  // per warp: one stage in flight, one being consumed
    static constexpr int kPanel = LPR * W;          // columns one warp covers (>= K: a single panel)
    static_assert(QS % UB == 0 && Q % QS == 0, "bad stage shape");

    float init_v;
    const int *__restrict__ colind;
    const float *__restrict__ val;
    const char *__restrict__ Bl;  // B + this lane's 4 columns (byte pointer)
    float *__restrict__ Cl;
    unsigned ldb_bytes;
    int ldc;
    int lane, g;                  // g = lane / LPR: which nonzero of a quad this lane works on
    bool active;                  // this lane's 4 columns lie inside K
    unsigned ring;                // shared-space address of this lane's 16 bytes in (stage 0, quad 0)
    // FUSE (see WalkerRing)
    const float *__restrict__ col_scale;
    const float *__restrict__ row_scale;
    float my_rs;
    T bias_v;
    bool has_bias;

    __device__ __forceinline__ T start() const { return R::start(init_v); }

    __device__ __forceinline__ void init(const Operands &o, int /*panel*/, int K, int ln, unsigned ring_base) {
        lane = ln; g = ln / LPR;
        const int col0 = (ln % LPR) * W;
        active = col0 < K;
        colind = o.colind; val = o.val; Bl = reinterpret_cast<const char *>(o.B + col0); Cl = o.C + col0;
        ldb_bytes = (unsigned)o.ldb * 4u; ldc = o.ldc;
        ring = ring_base + ln * (4 * W);
        init_v = o.init;
        if constexpr (FUSE) {
            col_scale = o.col_scale; row_scale = o.row_scale; my_rs = 1.f;
            has_bias = o.bias != nullptr;
            bias_v = P::ldg_or_zero(o.bias + col0, has_bias && active);
        }
    }
    __device__ __forceinline__ void load_rows(int rb, int nrows) {
        if constexpr (FUSE) my_rs = (row_scale && lane < nrows) ? __ldg(row_scale + rb + lane) : 1.f;
    }
    __device__ __forceinline__ float load_scale(int c, bool on) const {
        if constexpr (SCALE) return (on && col_scale) ? __ldg(col_scale + c) : 1.f;
        else return 1.f;
    }

    // whole-row values live in group 0 (after combine: in every group)
    __device__ __forceinline__ void store_row(int rb, int rel, const T (&acc)[1]) const {
        float rs = 1.f;
        if constexpr (FUSE) rs = __shfl_sync(kFull, my_rs, rel);
        if (g == 0 && active) P::stcs(Cl + (long long)(rb + rel) * ldc, FUSE ? E::apply(acc[0], rs, bias_v, has_bias) : acc[0]);
    }

    __device__ __forceinline__ void combine(T &t) const {
#pragma unroll
        for (int off = 16; off >= LPR; off >>= 1) {
            T x;
            if constexpr (W == 4) {
                x.x = __shfl_xor_sync(kFull, t.x, off); x.y = __shfl_xor_sync(kFull, t.y, off);
                x.z = __shfl_xor_sync(kFull, t.z, off); x.w = __shfl_xor_sync(kFull, t.w, off);
            } else {
                x = __shfl_xor_sync(kFull, t, off);
            }
            R::merge(t, x);
        }
    }
    // segment mode (kernel B): fold the groups' partials so that lanes [0, LPR) hold the segment's sums
    __device__ __forceinline__ void finish(T (&acc)[1]) const { combine(acc[0]); }

    __device__ __forceinline__ void flush(T &acc, unsigned &rows_left, int rb) const {
        T t = acc;
        combine(t);
        const int rel = __ffs(rows_left) - 1;
        float rs = 1.f;
        if constexpr (FUSE) rs = __shfl_sync(kFull, my_rs, rel);
        if (g == 0 && active) P::stcs(Cl + (long long)(rb + rel) * ldc, FUSE ? E::apply(t, rs, bias_v, has_bias) : t);
        rows_left &= rows_left - 1;
        acc = start();
    }

    // copies for the SN nonzeros at chunk positions [pos0, pos0 + SN) of a chunk holding n nonzeros into
    // the stage at byte offset `slot`; always exactly one commit group
    template <bool FULL>
    __device__ __forceinline__ void issue_impl(int cols, int pos0, int n, unsigned slot) const {
#pragma unroll
        for (int i0 = 0; i0 < QS; i0 += UB) {
            const char *bp[UB];
#pragma unroll
            for (int i = 0; i < UB; i++) {
                const int c = __shfl_sync(kFull, cols, pos0 + (i0 + i) * NG + g);
                bp[i] = Bl + (unsigned long long)(unsigned)c * ldb_bytes;
            }
#pragma unroll
            for (int i = 0; i < UB; i++)
                if (active && (FULL || pos0 + (i0 + i) * NG + g < n)) {
                    if constexpr (W == 4) cp_async16<0>(ring + slot + (i0 + i) * kRowBytes, bp[i]);
                    else cp_async4(ring + slot + (i0 + i) * kRowBytes, bp[i]);
                }
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void issue(int cols, int pos0, int n, unsigned slot) const {
        if (pos0 + SN <= n) issue_impl<true>(cols, pos0, n, slot);
        else issue_impl<false>(cols, pos0, n, slot);
    }

    template <bool FULL, bool SC>
    __device__ __forceinline__ void consume_impl(float vals, float scales, int pos0, int n, unsigned endmask, T &acc,
                                                 unsigned &rows_left, int rb, unsigned slot) const {
#pragma unroll
        for (int i0 = 0; i0 < QS; i0 += UB) {
            T b[UB];
            float a[UB];
            float sc[UB];
#pragma unroll
            for (int i = 0; i < UB; i++) {
                // read back unconditionally: a slice that was not copied this time (nonzero beyond n, columns
                // beyond K) holds stale ring bytes that are never added to anything that is stored
                a[i] = VALUED ? __shfl_sync(kFull, vals, pos0 + (i0 + i) * NG + g) : 1.f;
                sc[i] = SC ? __shfl_sync(kFull, scales, pos0 + (i0 + i) * NG + g) : 1.f;
                if constexpr (W == 4) b[i] = lds128(ring + slot + (i0 + i) * kRowBytes);
                else b[i] = lds32(ring + slot + (i0 + i) * kRowBytes);
            }
            if (FULL && ((endmask >> (pos0 + i0 * NG)) & low_bits(UB * NG)) == 0u) {  // no row ends in these UB quads
#pragma unroll
                for (int i = 0; i < UB; i++) R::template step<SC>(acc, a[i], b[i], sc[i]);
                continue;
            }
#pragma unroll
            for (int i = 0; i < UB; i++) {
                const bool live = FULL || (pos0 + (i0 + i) * NG + g < n);
                unsigned e4 = (endmask >> (pos0 + (i0 + i) * NG)) & ((1u << NG) - 1u);  // bit x: a row ends at group x's nonzero
                if (e4 == 0u) {  // warp-uniform: no row ends inside this quad (the common case for rows >> NG)
                    if (live) R::template step<SC>(acc, a[i], b[i], sc[i]);
                } else {
                    int lo = 0;  // groups below `lo` already added their nonzero of this quad (to an earlier row)
                    do {
                        const int hi = __ffs(e4) - 1;
                        if (live && g >= lo && g <= hi) R::template step<SC>(acc, a[i], b[i], sc[i]);
                        flush(acc, rows_left, rb);
                        lo = hi + 1;
                        e4 &= e4 - 1;
                    } while (e4);
                    if (live && g >= lo) R::template step<SC>(acc, a[i], b[i], sc[i]);
                }
            }
        }
    }
    __device__ __forceinline__ void consume(float vals, float scales, int pos0, int n, unsigned endmask, T &acc,
                                            unsigned &rows_left, int rb, unsigned slot) const {
        if (pos0 + SN <= n) consume_impl<true, SCALE>(vals, scales, pos0, n, endmask, acc, rows_left, rb, slot);
        else consume_impl<false, SCALE>(vals, scales, pos0, n, endmask, acc, rows_left, rb, slot);
    }

    // same contract as WalkerRing::stream; with rows == 0 the groups' partials are left in acc (see finish)
    __device__ __forceinline__ void stream(int s, int e, T (&acc)[1], int my_end, unsigned rows, int rb) const {
        int ccol = 0, ncol = 0, fcol = 0;
        float cval = 1.f, nval = 1.f;
        [[maybe_unused]] float csc = 1.f, nsc = 1.f;
        if (s + lane < e) {
            ccol = __ldcs(colind + s + lane);
            if (VALUED) cval = __ldcs(val + s + lane);
        }
        if (s + 32 + lane < e) ncol = __ldcs(colind + s + 32 + lane);
        if constexpr (SCALE) csc = load_scale(ccol, s + lane < e);
        const bool my_row = (rows >> lane) & 1u;
        unsigned rows_left = rows;
        unsigned slot = 0;  // stage being consumed; the other one is being filled
        issue(ccol, 0, min(32, e - s), 0);
#pragma unroll 1
        for (int p0 = s; p0 < e; p0 += 32) {
            if (p0 + 64 + lane < e) fcol = __ldcs(colind + p0 + 64 + lane);
            if (VALUED && p0 + 32 + lane < e) nval = __ldcs(val + p0 + 32 + lane);
            if constexpr (SCALE) nsc = load_scale(ncol, p0 + 32 + lane < e);
            const unsigned rel = (unsigned)(my_end - 1 - p0);
            const unsigned endmask = __reduce_or_sync(kFull, (my_row && rel < 32u) ? (1u << rel) : 0u);
            const int n = min(32, e - p0);
            const int n_next = e - p0 - 32;
#pragma unroll
            for (int j = 0; j < SPC; j++) {
                if (j * SN >= n) break;  // only in the last chunk, where nothing is in flight past it
                const bool nxt = j + 1 >= SPC;  // the stage to fill next opens the next chunk
                issue(nxt ? ncol : ccol, nxt ? 0 : (j + 1) * SN, nxt ? n_next : n, slot ^ kStageBytes);
                cp_async_wait<1>();
                consume(cval, csc, j * SN, n, endmask, acc[0], rows_left, rb, slot);
                slot ^= kStageBytes;
            }
            ccol = ncol; ncol = fcol; cval = nval;
            if constexpr (SCALE) csc = nsc;
        }
        cp_async_wait<0>();
    }
};

// =================================================================================================
// Row-parallel narrow walker (K <= 64): the lane groups own DISJOINT ROWS, sequential order kept.
// =================================================================================================
// Same lane layout and ring as WalkerSub, but instead of dealing consecutive nonzeros of one row to
// the NG groups (which re-associates the row's sum), the rows of the current run (<= 32 rows whose
// nonzeros are contiguous) are dealt to the groups as NG contiguous blocks, balanced on nonzeros by
// the midpoint of each row's range.  Every group then walks its own rows' nonzeros in CSR order,
// LPR = 32 / NG nonzeros per chunk, all groups in lock-step: step k of a chunk copies / reads back /
// accumulates the k-th nonzero of each group's chunk.  One accumulator per output element, products
// added in CSR order -- bit-identical to the reference like the ring walker -- and a row end is a
// predicated store by the group that owns the row: no cross-group combine.  The price is balance: a
// run takes as many chunks as its longest group needs (measured on the generators' degree
// distributions: 95-97 % of ideal at NG = 2, 82-92 % at NG = 4, 53-73 % at NG = 8).
// Rows only; kernel B (long rows, re-associated anyway) pairs it with WalkerSub.
template <int NG, bool VALUED, bool MAXR = false, bool FUSE = false, bool SCALE = FUSE>
struct WalkerRows {
    using P = Pack<true>;
    using T = float4;
    using R = Reduce<P, VALUED, MAXR, FUSE>;
    using E = Epilogue<P, FUSE>;
    static_assert(NG == 2 || NG == 4 || NG == 8, "2, 4 or 8 row blocks per warp");
    static constexpr int LPR = 32 / NG;             // lanes per group = nonzeros per group per chunk
    static constexpr int QS = LPR < 8 ? LPR : 8;    // steps per ring stage
    static constexpr int SPC = LPR / QS;            // stages per chunk (1 or 2)
    static constexpr int UB = 4;                    // steps read back from the ring per LDS batch
    static constexpr int kStageBytes = QS * 512;
    static constexpr int kRingBytes = 2 * kStageBytes;
    static constexpr int kPanel = LPR * 4;
    static constexpr unsigned kGroupOnes = LPR == 4 ? 0x11111111u : (LPR == 8 ? 0x01010101u : 0x00010001u);  // bit 0 of every group's field
    static_assert(QS % UB == 0 && LPR % QS == 0, "bad stage shape");

    float init_v;
    const int *__restrict__ colind;
    const float *__restrict__ val;
    const char *__restrict__ Bl;
    float *__restrict__ Cl;
    unsigned ldb_bytes;
    int ldc;
    int lane, g, sl;              // group and position inside the group
    bool active;                  // this lane's 4 columns lie inside K
    unsigned ring;
    // FUSE (see WalkerRing)
    const float *__restrict__ col_scale;
    const float *__restrict__ row_scale;
    float my_rs;
    T bias_v;
    bool has_bias;

    __device__ __forceinline__ T start() const { return R::start(init_v); }

    __device__ __forceinline__ void init(const Operands &o, int /*panel*/, int K, int ln, unsigned ring_base) {
        lane = ln; g = ln / LPR; sl = ln % LPR;
        const int col0 = sl * 4;
        active = col0 < K;
        colind = o.colind; val = o.val; Bl = reinterpret_cast<const char *>(o.B + col0); Cl = o.C + col0;
        ldb_bytes = (unsigned)o.ldb * 4u; ldc = o.ldc;
        ring = ring_base + ln * 16;
        init_v = o.init;
        if constexpr (FUSE) {
            col_scale = o.col_scale; row_scale = o.row_scale; my_rs = 1.f;
            has_bias = o.bias != nullptr;
            bias_v = P::ldg_or_zero(o.bias + col0, has_bias && active);
        }
    }
    __device__ __forceinline__ void load_rows(int rb, int nrows) {
        if constexpr (FUSE) my_rs = (row_scale && lane < nrows) ? __ldg(row_scale + rb + lane) : 1.f;
    }
    __device__ __forceinline__ float load_scale(int c, bool on) const {
        if constexpr (SCALE) return (on && col_scale) ? __ldg(col_scale + c) : 1.f;
        else return 1.f;
    }

    // kernel A's empty rows: one group writes the row
    __device__ __forceinline__ void store_row(int rb, int rel, const T (&acc)[1]) const {
        float rs = 1.f;
        if constexpr (FUSE) rs = __shfl_sync(kFull, my_rs, rel);
        if (g == 0 && active) P::stcs(Cl + (long long)(rb + rel) * ldc, FUSE ? E::apply(acc[0], rs, bias_v, has_bias) : acc[0]);
    }

    // copies for steps [k0, k0 + QS) of a chunk in which my group still has n nonzeros; one commit group
    __device__ __forceinline__ void issue(int cols, int k0, int n, unsigned slot) const {
#pragma unroll
        for (int i0 = 0; i0 < QS; i0 += UB) {
            const char *bp[UB];
#pragma unroll
            for (int i = 0; i < UB; i++) {
                const int c = __shfl_sync(kFull, cols, k0 + i0 + i, LPR);  // the (k0+i0+i)-th lane of my group
                bp[i] = Bl + (unsigned long long)(unsigned)c * ldb_bytes;
            }
#pragma unroll
            for (int i = 0; i < UB; i++)
                if (active && k0 + i0 + i < n) cp_async16<0>(ring + slot + (i0 + i) * 512, bp[i]);
        }
        cp_async_commit();
    }

    // steps [k0, k0 + QS) of the chunk: n = my group's nonzeros left, nmin = the least over the groups
    __device__ __forceinline__ void consume(float vals, float scales, int k0, int n, int nmin, unsigned endmask, T &acc,
                                            unsigned &left, int rb, unsigned slot) const {
        consume_impl<SCALE>(vals, scales, k0, n, nmin, endmask, acc, left, rb, slot);
    }
    template <bool SC>
    __device__ __forceinline__ void consume_impl(float vals, float scales, int k0, int n, int nmin, unsigned endmask, T &acc,
                                                 unsigned &left, int rb, unsigned slot) const {
#pragma unroll
        for (int i0 = 0; i0 < QS; i0 += UB) {
            T b[UB];
            float a[UB];
            float sc[UB];
#pragma unroll
            for (int i = 0; i < UB; i++) {
                a[i] = VALUED ? __shfl_sync(kFull, vals, k0 + i0 + i, LPR) : 1.f;
                sc[i] = SC ? __shfl_sync(kFull, scales, k0 + i0 + i, LPR) : 1.f;
                b[i] = lds128(ring + slot + (i0 + i) * 512);  // stale bytes where nothing was copied: never added
            }
            const unsigned ends = endmask & ((kGroupOnes * ((1u << UB) - 1u)) << (k0 + i0));  // any group, these UB steps
            if (ends == 0u && nmin >= k0 + i0 + UB) {  // warp-uniform: every group is live and no row ends
#pragma unroll
                for (int i = 0; i < UB; i++) R::template step<SC>(acc, a[i], b[i], sc[i]);
            } else {
#pragma unroll
                for (int i = 0; i < UB; i++) {
                    const int k = k0 + i0 + i;
                    if (k < n) R::template step<SC>(acc, a[i], b[i], sc[i]);
                    const int rel = (__ffs(left) - 1) & 31;  // my group's current row (any lane when the group is done)
                    float rs = 1.f;
                    if constexpr (FUSE) rs = __shfl_sync(kFull, my_rs, rel);  // every lane takes part: the groups diverge below
                    if ((endmask >> (g * LPR + k)) & 1u) {  // my group's current row ends with this nonzero
                        if (active) P::stcs(Cl + (long long)(rb + rel) * ldc, FUSE ? E::apply(acc, rs, bias_v, has_bias) : acc);
                        left &= left - 1;
                        acc = start();
                    }
                }
            }
        }
    }

    // Rows `rows` (bits = row - rb; non-empty, their nonzeros are exactly [s, e), lane r holds row r's end).
    __device__ __forceinline__ void stream(int s, int e, T (&accv)[1], int my_end, unsigned rows, int rb) const {
        T acc = accv[0];
        const bool my_row = (rows >> lane) & 1u;
        // the row I hold starts where the previous row of the run ends
        const unsigned below = rows & low_bits(lane);
        const int prev_end = __shfl_sync(kFull, my_end, below ? 31 - __clz(below) : 0);
        const int my_start = below ? prev_end : s;
        // its block: by the midpoint of its nonzero range inside [s, e)
        int grp = 0;
        if (my_row) grp = min(NG - 1, (((my_start - s) + (my_end - s)) * NG) / (2 * (e - s)));  // offsets < 32 * long_row: no overflow
        unsigned mine = 0;  // the rows of my group
#pragma unroll
        for (int q = 0; q < NG; q++) {
            const unsigned m = __ballot_sync(kFull, my_row && grp == q);
            if (q == g) mine = m;
        }
        int gs = __shfl_sync(kFull, my_start, mine ? __ffs(mine) - 1 : 0);
        int ge = __shfl_sync(kFull, my_end, mine ? 31 - __clz(mine) : 0);
        if (!mine) gs = ge = 0;
        const int row_gs = __shfl_sync(kFull, gs, grp * LPR);  // where the stream of my ROW's group starts
        const int maxlen = __reduce_max_sync(kFull, ge - gs);   // the run takes ceil(maxlen / LPR) chunks

        unsigned left = mine;
        int p = gs;
        int ccol = 0, ncol = 0, fcol = 0;
        float cval = 1.f, nval = 1.f;
        [[maybe_unused]] float csc = 1.f, nsc = 1.f;
        if (p + sl < ge) {
            ccol = __ldcs(colind + p + sl);
            if (VALUED) cval = __ldcs(val + p + sl);
        }
        if (p + LPR + sl < ge) ncol = __ldcs(colind + p + LPR + sl);
        if constexpr (SCALE) csc = load_scale(ccol, p + sl < ge);
        unsigned slot = 0;
        issue(ccol, 0, ge - p, 0);
#pragma unroll 1
        for (int c0 = 0; c0 < maxlen; c0 += LPR, p += LPR) {
            if (p + 2 * LPR + sl < ge) fcol = __ldcs(colind + p + 2 * LPR + sl);
            if (VALUED && p + LPR + sl < ge) nval = __ldcs(val + p + LPR + sl);
            if constexpr (SCALE) nsc = load_scale(ncol, p + LPR + sl < ge);
            // row ends of every group inside this chunk: one LPR-bit field per group
            const unsigned rel = (unsigned)(my_end - 1 - (row_gs + c0));
            const unsigned endmask = __reduce_or_sync(kFull, (my_row && rel < (unsigned)LPR) ? (1u << (grp * LPR + rel)) : 0u);
            const int n = ge - p;  // my group's nonzeros from this chunk on (<= 0: done)
            const int nmin = __reduce_min_sync(kFull, n);
#pragma unroll
            for (int j = 0; j < SPC; j++) {
                if (j * QS >= maxlen - c0) break;  // only in the last chunk, where nothing is in flight past it
                const bool nxt = j + 1 >= SPC;
                issue(nxt ? ncol : ccol, nxt ? 0 : (j + 1) * QS, nxt ? n - LPR : n, slot ^ kStageBytes);
                cp_async_wait<1>();
                consume(cval, csc, j * QS, n, nmin, endmask, acc, left, rb, slot);
                slot ^= kStageBytes;
            }
            ccol = ncol; ncol = fcol; cval = nval;
            if constexpr (SCALE) csc = nsc;
        }
        cp_async_wait<0>();
        accv[0] = acc;
    }
};

// =================================================================================================
// Kernel A: short rows.  One warp per CTA, one task per CTA.
// =================================================================================================
template <class WK, int V, bool VEC4, int MINB>
__global__ void __launch_bounds__(32, MINB)
spmm_flat_kernel(int M, int K, long long total_keys, int task, int long_row, const int *__restrict__ rowptr, Operands op)
{
    using P = Pack<VEC4>;
    using T = typename P::T;
    extern __shared__ __align__(16) unsigned char s_dyn[];  // gather ring (ring walkers only)

    const int lane = threadIdx.x;
    WK wk;
    wk.init(op, blockIdx.y, K, lane, (unsigned)__cvta_generic_to_shared(s_dyn));

    // ---- this task's rows ---------------------------------------------------------------------
    const long long k0 = (long long)blockIdx.x * task;
    int row_lo, row_hi;
    {
        const int shift = lane & 16;
        const long long target = k0 + (shift ? task : 0);
        const int r = search_key16(rowptr, M, target < total_keys ? target : total_keys + 1, lane & 15, shift);
        row_lo = __shfl_sync(kFull, r, 0);
        row_hi = __shfl_sync(kFull, r, 16);
    }

    for (int rb = row_lo; rb < row_hi; rb += 32) {
        const int nrows = min(32, row_hi - rb);
        int my_start = 0, my_end = 0;
        if (lane < nrows) {
            my_start = __ldg(rowptr + rb + lane);
            my_end = __ldg(rowptr + rb + lane + 1);
        }
        wk.load_rows(rb, nrows);
        const int len = my_end - my_start;
        unsigned long_mask = __ballot_sync(kFull, len > long_row);  // left to kernel B
        const unsigned nonempty = __ballot_sync(kFull, len > 0) & ~long_mask;
        {   // empty rows: zeros, one warp-wide store per row
            unsigned em = ~(nonempty | long_mask) & low_bits(nrows);
            T z[V];
#pragma unroll
            for (int v = 0; v < V; v++) z[v] = wk.start();
            while (em) {
                wk.store_row(rb, __ffs(em) - 1, z);
                em &= em - 1;
            }
        }
        // runs of short rows between long rows, each walked as one flat stream
        int run = 0;
        while (true) {
            const int stop = long_mask ? (__ffs(long_mask) - 1) : nrows;  // next long row, or end of chunk
            const unsigned rows = nonempty & low_bits(stop) & ~low_bits(run);
            if (rows) {
                const int s = __shfl_sync(kFull, my_start, __ffs(rows) - 1);
                const int e = __shfl_sync(kFull, my_end, 31 - __clz(rows));
                T acc[V];
#pragma unroll
                for (int v = 0; v < V; v++) acc[v] = wk.start();
                wk.stream(s, e, acc, my_end, rows, rb);
            }
            if (stop >= nrows) break;
            long_mask &= long_mask - 1;
            run = stop + 1;
        }
    }
}

// =================================================================================================
// Kernel A for tiny B rows that are not 16-byte multiples (K <= 16 with K % 4 != 0, or unaligned operands)
// =================================================================================================
// The class-count widths of a GCN's last layer (K = 3, 7, ...).  A B row is then 12-60 bytes at a 4-byte aligned
// address: no 16-byte cp.async, and one row per warp instruction would keep 3-15 of 32 lanes busy.  Here a warp is
// NG = 32 / LPR lane groups (LPR = 4 / 8 / 16 lanes >= K), lane `sl` of a group owns column `sl`, and every group sums
// its OWN row: nonzeros in CSR order into one accumulator per element -- the reference's order, bit-identical -- LPR
// nonzeros per step (the group's lanes hold the next LPR column indices, the following LPR are prefetched), each
// lane issuing up to 8 independent 4-byte gathers before the first add.  Rows are dealt dynamically: the lanes hold
// the row bounds of a 32-row batch, and a group that finishes its row takes the next unassigned row of the batch
// (ranked by ballot among the groups finishing in the same step), so a long row delays only its own group and the next
// batch is fetched as soon as the current one has been handed out.  No shared memory; B this narrow lives in the L2.
// Same task windows as spmm_flat_kernel (one task per warp); rows above `long_row` are left to kernel B.
constexpr int kRgWarps = 4;        // warps (= tasks) per CTA
constexpr int kRowGroupMaxK = 16;  // widest B row the row-group kernel takes

template <int LPR, bool VALUED, bool MAXR, bool FUSE>
__global__ void __launch_bounds__(kRgWarps * 32)
spmm_rowgroup_kernel(int M, int K, long long total_keys, int task, int long_row, const int *__restrict__ rowptr, Operands op)
{
    using P = Pack<false>;
    using R = Reduce<P, VALUED, MAXR, FUSE>;
    using E = Epilogue<P, FUSE>;
    static_assert(LPR == 4 || LPR == 8 || LPR == 16, "4, 8 or 16 lanes per row");
    constexpr int NG = 32 / LPR;
    constexpr int UB = LPR < 8 ? LPR : 8;  // gathers in flight per lane
    constexpr unsigned kLeaders = LPR == 4 ? 0x11111111u : (LPR == 8 ? 0x01010101u : 0x00010001u);  // lane 0 of every group

    const int lane = threadIdx.x & 31;
    const int g = lane / LPR, sl = lane % LPR, lead = g * LPR;
    const bool active = sl < K;
    const long long k0 = ((long long)blockIdx.x * kRgWarps + (threadIdx.x >> 5)) * task;
    if (k0 >= total_keys) return;  // whole warps leave; the kernel has no CTA-wide synchronisation

    int row_lo, row_hi;
    {
        const int shift = lane & 16;
        const long long target = k0 + (shift ? task : 0);
        const int r = search_key16(rowptr, M, target < total_keys ? target : total_keys + 1, lane & 15, shift);
        row_lo = __shfl_sync(kFull, r, 0);
        row_hi = __shfl_sync(kFull, r, 16);
    }

    const int *__restrict__ colind = op.colind;
    const float *__restrict__ val = op.val;
    const float *__restrict__ Bl = op.B + sl;
    float *__restrict__ Cl = op.C + sl;
    const int ldb = op.ldb, ldc = op.ldc;
    const float init_v = op.init;
    [[maybe_unused]] const bool has_bias = FUSE && op.bias != nullptr;
    [[maybe_unused]] const float bias = (has_bias && active) ? __ldg(op.bias + sl) : 0.f;
    const unsigned below = kLeaders & ((1u << lead) - 1u);  // the leaders of the groups before mine

    // the batch of 32 rows being handed out: lane i holds the bounds (and scale) of row rb + i
    int rb = row_lo, next_rb = row_lo;
    unsigned rows_left = 0;  // rows of the batch (bits = row - rb) that wait for a group: non-empty, not long
    int my_start = 0, my_end = 0;
    [[maybe_unused]] float my_rs = 1.f;
    // my group's row
    int row = -1, p = 0, e = 0;
    int ccol = 0, ncol = 0;
    float cval = 1.f, nval = 1.f;
    [[maybe_unused]] float csc = 1.f, rs = 1.f;
    float acc = R::start(init_v);

    while (true) {
        const bool need = row < 0;
        const unsigned needm = __ballot_sync(kFull, need) & kLeaders;
        // ---- next batch, as soon as some group is idle and the current batch has been handed out --------------
        while (needm != 0u && rows_left == 0u && next_rb < row_hi) {
            rb = next_rb;
            next_rb += 32;
            const int nrows = min(32, row_hi - rb);
            my_start = my_end = 0;
            if (lane < nrows) {
                my_start = __ldg(rowptr + rb + lane);
                my_end = __ldg(rowptr + rb + lane + 1);
            }
            if constexpr (FUSE) my_rs = (op.row_scale && lane < nrows) ? __ldg(op.row_scale + rb + lane) : 1.f;
            const int len = my_end - my_start;
            const unsigned longm = __ballot_sync(kFull, len > long_row);  // left to kernel B
            rows_left = __ballot_sync(kFull, len > 0) & ~longm;
            unsigned em = ~(rows_left | longm) & low_bits(nrows);
            while (em) {  // empty rows: the reduction's start value, NG rows per store instruction
                const unsigned bit = __fns(em, 0, g + 1);
                [[maybe_unused]] float rsv = 1.f;
                if constexpr (FUSE) rsv = __shfl_sync(kFull, my_rs, bit & 31u);
                if (bit != 0xffffffffu && active)
                    __stcs(Cl + (long long)(rb + (int)bit) * ldc, E::apply(R::start(init_v), rsv, bias, has_bias));
#pragma unroll
                for (int i = 0; i < NG; i++) em &= em - 1;  // (0 & -1 stays 0)
            }
        }
        // ---- idle groups take the next rows of the batch, in group order ----------------------------------------
        if (needm != 0u && rows_left != 0u) {
            const unsigned bit = need ? __fns(rows_left, 0, __popc(needm & below) + 1) : 0xffffffffu;
            const bool got = bit != 0xffffffffu;
            const int src = got ? (int)bit : 0;
            const int ns = __shfl_sync(kFull, my_start, src), ne = __shfl_sync(kFull, my_end, src);
            [[maybe_unused]] float nrs = 1.f;
            if constexpr (FUSE) nrs = __shfl_sync(kFull, my_rs, src);
            if (got) {
                row = rb + src; p = ns; e = ne;
                acc = R::start(init_v);
                if constexpr (FUSE) rs = nrs;
                ccol = ncol = 0;
                if (p + sl < e) {
                    ccol = __ldcs(colind + p + sl);
                    if (VALUED) cval = __ldcs(val + p + sl);
                }
                if (p + LPR + sl < e) {
                    ncol = __ldcs(colind + p + LPR + sl);
                    if (VALUED) nval = __ldcs(val + p + LPR + sl);
                }
                if constexpr (FUSE) csc = (op.col_scale && p + sl < e) ? __ldg(op.col_scale + ccol) : 1.f;
            }
            const int taken = min(__popc(needm), __popc(rows_left));
            for (int i = 0; i < taken; i++) rows_left &= rows_left - 1;
        }
        if (__ballot_sync(kFull, row >= 0) == 0u) {
            if (rows_left == 0u && next_rb >= row_hi) break;  // every row of the task is done
            continue;                                          // (a batch of empty / long rows only)
        }
        // ---- one step: up to LPR nonzeros of my group's row -------------------------------------------------------
        const int n = row >= 0 ? min(LPR, e - p) : 0;
#pragma unroll
        for (int u0 = 0; u0 < LPR; u0 += UB) {
            float b[UB], a[UB];
            [[maybe_unused]] float sc[UB];
#pragma unroll
            for (int u = 0; u < UB; u++) {
                const int c = __shfl_sync(kFull, ccol, lead + u0 + u);
                a[u] = VALUED ? __shfl_sync(kFull, cval, lead + u0 + u) : 1.f;
                if constexpr (FUSE) sc[u] = __shfl_sync(kFull, csc, lead + u0 + u);
                b[u] = 0.f;
                if (u0 + u < n && active) b[u] = __ldg(Bl + (long long)c * ldb);
            }
#pragma unroll
            for (int u = 0; u < UB; u++)
                if (u0 + u < n) R::step(acc, a[u], b[u], FUSE ? sc[u] : 1.f);
        }
        p += n;
        if (row >= 0) {
            if (p >= e) {  // the row is complete
                if (active) __stcs(Cl + (long long)row * ldc, E::apply(acc, rs, bias, has_bias));
                row = -1;
            } else {       // the prefetched indices become current; fetch the ones after them
                ccol = ncol; cval = nval;
                if constexpr (FUSE) csc = (op.col_scale && p + sl < e) ? __ldg(op.col_scale + ccol) : 1.f;
                ncol = 0;
                if (p + LPR + sl < e) {
                    ncol = __ldcs(colind + p + LPR + sl);
                    if (VALUED) nval = __ldcs(val + p + LPR + sl);
                }
            }
        }
    }
}

// =================================================================================================
// Kernel B: long rows.  8 warps per CTA, clusters of 8 CTAs.  Claim by probing, then segmented
// cooperative sums: one CTA per long row, the whole cluster (64 warps, partials combined through
// distributed shared memory) per huge row.
// =================================================================================================
template <class WK, int V, bool VEC4>
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kLongWarps * 32)
spmm_long_kernel(int M, int K, int nnz, int long_row, const int *__restrict__ rowptr, Operands op)
{
    namespace cg = cooperative_groups;
    using P = Pack<VEC4>;
    using T = typename P::T;
    constexpr int W = P::kWidth;
    __shared__ int s_rows[kMaxList];
    __shared__ int s_huge[kMaxHuge];
    __shared__ int s_n, s_nhuge;
    __shared__ T s_part[2][kLongWarps][V * 32];
    __shared__ T s_cpart[V * 32];  // this CTA's partial of a huge row, read by cluster rank 0
    extern __shared__ __align__(16) unsigned char s_dyn[];

    cg::cluster_group cluster = cg::this_cluster();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { s_n = 0; s_nhuge = 0; }
    __syncthreads();

    // ---- claim: the row containing my probe position, if long and this is its first probe --------
    {
        const long long p = ((long long)blockIdx.x * (kLongWarps * 32) + threadIdx.x) * kProbeStride;
        if (p < nnz) {
            int lo = 0, hi = M - 1;  // first r with rowptr[r + 1] > p
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(rowptr + mid + 1) > p) hi = mid;
                else lo = mid + 1;
            }
            const int a = __ldg(rowptr + lo), b = __ldg(rowptr + lo + 1);
            if (b - a > long_row && p - a < kProbeStride) {
                if (b - a >= kHugeRow) {
                    const int slot = atomicAdd(&s_nhuge, 1);
                    if (slot < kMaxHuge) s_huge[slot] = lo;
                } else {
                    const int slot = atomicAdd(&s_n, 1);
                    if (slot < kMaxList) s_rows[slot] = lo;
                }
            }
        }
    }
    __syncthreads();

    WK wk;
    wk.init(op, blockIdx.y, K, lane, (unsigned)__cvta_generic_to_shared(s_dyn) + warp * WK::kRingBytes);

    // ---- long rows: this CTA's 8 warps, contiguous segments, fixed-order combine ------------------
    const int nlist = min(s_n, kMaxList);
    for (int i = 0; i < nlist; i++) {
        const int r = s_rows[i];
        const int a = __ldg(rowptr + r), b = __ldg(rowptr + r + 1);
        int seg = (b - a + kLongWarps - 1) / kLongWarps;
        seg = (seg + 31) & ~31;
        const int s = min(b, a + warp * seg), e = min(b, s + seg);
        T acc[V];
#pragma unroll
        for (int v = 0; v < V; v++) acc[v] = wk.start();
        wk.stream(s, e, acc, 0, 0u, 0);
        wk.finish(acc);
        T(*part)[V * 32] = s_part[i & 1];  // double-buffered: one barrier per row
#pragma unroll
        for (int v = 0; v < V; v++) part[warp][v * 32 + lane] = acc[v];
        __syncthreads();
        for (int x = threadIdx.x; x < V * 32; x += kLongWarps * 32) {
            T sum = part[0][x];
#pragma unroll
            for (int w = 1; w < kLongWarps; w++) WK::R::merge(sum, part[w][x]);
            const int c = blockIdx.y * WK::kPanel + x * W;
            if (c < K) {
                if constexpr (WK::kFuse)
                    sum = Epilogue<P, true>::apply(sum, op.row_scale ? __ldg(op.row_scale + r) : 1.f,
                                                   P::ldg_or_zero(op.bias + c, op.bias != nullptr), op.bias != nullptr);
                P::stcs(op.C + (long long)r * op.ldc + c, sum);
            }
        }
    }

    // ---- huge rows: all 64 warps of the cluster; CTA partials meet in cluster rank 0 via DSMEM -----
    cluster.sync();  // every CTA's huge list is complete and visible cluster-wide
    const unsigned crank = cluster.block_rank(), csize = cluster.num_blocks();
    for (unsigned q = 0; q < csize; q++) {
        const int cnt = min(*cluster.map_shared_rank(&s_nhuge, q), kMaxHuge);
        const int *list = cluster.map_shared_rank(s_huge, q);
        for (int i = 0; i < cnt; i++) {
            const int r = list[i];
            const int a = __ldg(rowptr + r), b = __ldg(rowptr + r + 1);
            const int nseg = (int)csize * kLongWarps;
            int seg = (b - a + nseg - 1) / nseg;
            seg = (seg + 31) & ~31;
            const int s = min(b, a + ((int)crank * kLongWarps + warp) * seg), e = min(b, s + seg);
            T acc[V];
#pragma unroll
            for (int v = 0; v < V; v++) acc[v] = wk.start();
            wk.stream(s, e, acc, 0, 0u, 0);
            wk.finish(acc);
#pragma unroll
            for (int v = 0; v < V; v++) s_part[0][warp][v * 32 + lane] = acc[v];
            __syncthreads();
            for (int x = threadIdx.x; x < V * 32; x += kLongWarps * 32) {
                T sum = s_part[0][0][x];
#pragma unroll
                for (int w = 1; w < kLongWarps; w++) WK::R::merge(sum, s_part[0][w][x]);
                s_cpart[x] = sum;
            }
            cluster.sync();  // all CTA partials written
            if (crank == 0) {
                for (int x = threadIdx.x; x < V * 32; x += kLongWarps * 32) {
                    T sum = s_cpart[x];
                    for (unsigned c2 = 1; c2 < csize; c2++) WK::R::merge(sum, cluster.map_shared_rank(s_cpart, c2)[x]);
                    const int c = blockIdx.y * WK::kPanel + x * W;
                    if (c < K) {
                        if constexpr (WK::kFuse)
                            sum = Epilogue<P, true>::apply(sum, op.row_scale ? __ldg(op.row_scale + r) : 1.f,
                                                           P::ldg_or_zero(op.bias + c, op.bias != nullptr), op.bias != nullptr);
                        P::stcs(op.C + (long long)r * op.ldc + c, sum);
                    }
                }
            }
            cluster.sync();  // partials consumed: s_part / s_cpart may be rewritten
        }
    }
    cluster.sync();  // nobody exits while its shared memory may still be read by a peer
}

// =================================================================================================
// Host side
// =================================================================================================
struct Args {
    int M, K, task, long_row;
    bool overlap;   // run kernel B concurrently with kernel A (helper stream)
    bool has_long;  // some row may be longer than long_row: kernel B is needed
    int smem_pad;   // extra dynamic shared memory per CTA of kernel A (tuning: caps the resident CTAs per SM)
    long long nnz;
    const int *rowptr;
    Operands op;
    cudaStream_t st;
};

// Kernel B runs on a helper stream, forked from and joined back into the caller's stream with
// events, so that its long-running CTAs overlap kernel A instead of leaving the GPU idle behind
// their tail.  One helper stream + two events per (host thread, device), created on first use;
// the fork/join pattern is legal under stream capture, so the call stays graph-capturable.
struct Side {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    bool ok = false;
};
constexpr int kMaxDevices = 64;

Side *thread_sides();  // gespmm_spmm.cu: the calling thread's helper streams, one slot per device

inline Side *side_for_current_device()
{
    Side *sides = thread_sides();
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    Side &sd = sides[dev];
    if (!sd.ok) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = greatest priority: the long rows should be placed first
        const bool made = cudaStreamCreateWithPriority(&sd.stream, cudaStreamNonBlocking, hi) == cudaSuccess &&
                          cudaEventCreateWithFlags(&sd.fork, cudaEventDisableTiming) == cudaSuccess &&
                          cudaEventCreateWithFlags(&sd.join, cudaEventDisableTiming) == cudaSuccess;
        if (!made) {  // nothing half-made is kept
            if (sd.join) cudaEventDestroy(sd.join);
            if (sd.fork) cudaEventDestroy(sd.fork);
            if (sd.stream) cudaStreamDestroy(sd.stream);
            sd = Side();
            cudaGetLastError();
            return nullptr;
        }
        sd.ok = true;
    }
    return &sd;
}

// Opt a kernel into more than 48 KB of dynamic shared memory, once per (kernel, device).
template <class KernelT>
cudaError_t allow_smem(KernelT kern, int bytes, std::atomic<bool> (&done)[kMaxDevices])
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
    if (!done[dev].load(std::memory_order_acquire)) {
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        done[dev].store(true, std::memory_order_release);
    }
    return cudaSuccess;
}

// Kernel B (when some row may be long), forked onto the helper stream; *sd_out is the helper to join afterwards.
template <class WKB, int V, bool VEC4>
cudaError_t launch_long(const Args &a, unsigned panels, Side **sd_out)
{
    *sd_out = nullptr;
    if (!a.has_long) return cudaSuccess;
    Side *sd = a.overlap ? side_for_current_device() : nullptr;
    constexpr int dynB = WKB::kRingBytes * kLongWarps;
    auto kernB = spmm_long_kernel<WKB, V, VEC4>;
    if (dynB > 0) {  // static + dynamic shared memory exceeds the 48 KB default
        static std::atomic<bool> done[kMaxDevices];
        const cudaError_t e = allow_smem(kernB, dynB, done);
        if (e != cudaSuccess) return e;
    }
    cudaStream_t sb = a.st;
    if (sd) {
        if (cudaEventRecord(sd->fork, a.st) != cudaSuccess || cudaStreamWaitEvent(sd->stream, sd->fork, 0) != cudaSuccess)
            return cudaGetLastError();
        sb = sd->stream;
    }
    const unsigned ctas = (unsigned)((a.nnz + kProbeWindow - 1) / kProbeWindow);
    dim3 grid((ctas + kClusterSize - 1) / kClusterSize * kClusterSize, panels, 1);  // whole clusters
    kernB<<<grid, kLongWarps * 32, dynB, sb>>>(a.M, a.K, (int)a.nnz, a.long_row, a.rowptr, a.op);
    if (sd && cudaEventRecord(sd->join, sd->stream) != cudaSuccess) return cudaGetLastError();
    *sd_out = sd;
    return cudaGetLastError();
}

template <class WK, int V, bool VEC4, int MINB, class WKB = WK>
cudaError_t launch(const Args &a)
{
    const unsigned panels = (unsigned)((a.K + WK::kPanel - 1) / WK::kPanel);
    Side *sd = nullptr;
    const cudaError_t eb = launch_long<WKB, V, VEC4>(a, panels, &sd);
    if (eb != cudaSuccess) return eb;
    constexpr int dynA = WK::kRingBytes;
    auto kernA = spmm_flat_kernel<WK, V, VEC4, MINB>;
    const long long total = a.nnz + a.M;
    const long long ntask = (total + a.task - 1) / a.task;
    dim3 grid((unsigned)ntask, panels, 1);
    int pad = a.smem_pad;
    if (pad < 0 || dynA + pad > 48 * 1024) pad = 0;
    kernA<<<grid, 32, dynA + pad, a.st>>>(a.M, a.K, total, a.task, a.long_row, a.rowptr, a.op);
    if (sd && cudaStreamWaitEvent(a.st, sd->join, 0) != cudaSuccess) return cudaGetLastError();
    return cudaGetLastError();
}

// Tiny B rows that are not 16-byte multiples: spmm_rowgroup_kernel, long rows through the scalar register walker.
template <int LPR, bool VALUED, bool MAXR, bool FUSE>
cudaError_t launch_rowgroup(const Args &a)
{
    Side *sd = nullptr;
    const cudaError_t eb = launch_long<Walker<1, VALUED, false, 8, MAXR, FUSE>, 1, false>(a, 1u, &sd);
    if (eb != cudaSuccess) return eb;
    const long long total = a.nnz + a.M;
    const long long ntask = (total + a.task - 1) / a.task;
    const unsigned blocks = (unsigned)((ntask + kRgWarps - 1) / kRgWarps);
    spmm_rowgroup_kernel<LPR, VALUED, MAXR, FUSE><<<blocks, kRgWarps * 32, 0, a.st>>>(a.M, a.K, total, a.task, a.long_row, a.rowptr, a.op);
    if (sd && cudaStreamWaitEvent(a.st, sd->join, 0) != cudaSuccess) return cudaGetLastError();
    return cudaGetLastError();
}
template <bool VALUED, bool MAXR, bool FUSE>
cudaError_t dispatch_rowgroup(int K, const Args &a)
{
    if (K <= 4) return launch_rowgroup<4, VALUED, MAXR, FUSE>(a);
    if (K <= 8) return launch_rowgroup<8, VALUED, MAXR, FUSE>(a);
    return launch_rowgroup<16, VALUED, MAXR, FUSE>(a);
}

template <int V, bool VALUED, bool VEC4, int U, int MINB, bool MAXR = false, bool FUSE = false>
cudaError_t launch_reg(const Args &a) { return launch<Walker<V, VALUED, VEC4, U, MAXR, FUSE>, V, VEC4, MINB>(a); }

// Default ring shape per V: (G rows per stage, MINB CTAs per SM); two stages.
template <int V> struct Shape;
template <> struct Shape<1> { static constexpr int G = 8, MINB = 24; };
template <> struct Shape<2> { static constexpr int G = 4, MINB = 20; };
template <> struct Shape<3> { static constexpr int G = 2, MINB = 24; };
template <> struct Shape<4> { static constexpr int G = 2, MINB = 16; };

template <int V, bool VALUED, bool PEER, bool MAXR, bool FUSE = false, bool HINT = false, bool SCALE = FUSE>
cudaError_t launch_ring(const Args &a, bool masked)
{
    constexpr int G = Shape<V>::G, MINB = Shape<V>::MINB;
    return masked ? launch<WalkerRing<V, VALUED, G, 2, 0, true, PEER, MAXR, FUSE, HINT, 4, SCALE>, V, true, MINB>(a)
                  : launch<WalkerRing<V, VALUED, G, 2, 0, false, PEER, MAXR, FUSE, HINT, 4, SCALE>, V, true, MINB>(a);
}
template <bool VALUED, bool PEER, bool MAXR>
cudaError_t dispatch_ring(int V, const Args &a, bool masked)
{
    switch (V) {
        case 1: return launch_ring<1, VALUED, PEER, MAXR>(a, masked);
        case 2: return launch_ring<2, VALUED, PEER, MAXR>(a, masked);
        case 3: return launch_ring<3, VALUED, PEER, MAXR>(a, masked);
        default: return launch_ring<4, VALUED, PEER, MAXR>(a, masked);
    }
}

// Narrow B (K <= 64, aligned operands): NG = 2 / 4 / 8 nonzeros per warp-wide copy for K <= 64 / 32 / 16.
template <bool VALUED, bool MAXR, bool FUSE, bool SCALE = FUSE>
cudaError_t dispatch_sub(int K, const Args &a)
{
    if (K > 32) return launch<WalkerSub<2, VALUED, MAXR, FUSE, 4, SCALE>, 1, true, 24>(a);
    if (K > 16) return launch<WalkerSub<4, VALUED, MAXR, FUSE, 4, SCALE>, 1, true, 24>(a);
    return launch<WalkerSub<8, VALUED, MAXR, FUSE, 4, SCALE>, 1, true, 24>(a);
}

// Narrow B that is not made of 16-byte slices (any K <= 16, any 4-byte alignment): the same walker on 4-byte slices,
// NG = 2 / 4 / 8 nonzeros per warp-wide copy for K <= 16 / 8 / 4.
template <bool VALUED, bool MAXR, bool FUSE, bool SCALE = FUSE>
cudaError_t dispatch_sub1(int K, const Args &a)
{
    if (K > 8) return launch<WalkerSub<2, VALUED, MAXR, FUSE, 1, SCALE>, 1, false, 24>(a);
    if (K > 4) return launch<WalkerSub<4, VALUED, MAXR, FUSE, 1, SCALE>, 1, false, 24>(a);
    return launch<WalkerSub<8, VALUED, MAXR, FUSE, 1, SCALE>, 1, false, 24>(a);
}

// The row-parallel narrow walker (sequential order); long rows go to kernel B with the sub-warp walker.
template <bool VALUED, bool MAXR, bool FUSE, bool SCALE = FUSE>
cudaError_t dispatch_rows(int K, const Args &a)
{
    if (K > 32) return launch<WalkerRows<2, VALUED, MAXR, FUSE, SCALE>, 1, true, 24, WalkerSub<2, VALUED, MAXR, FUSE, 4, SCALE>>(a);
    if (K > 16) return launch<WalkerRows<4, VALUED, MAXR, FUSE, SCALE>, 1, true, 24, WalkerSub<4, VALUED, MAXR, FUSE, 4, SCALE>>(a);
    return launch<WalkerRows<8, VALUED, MAXR, FUSE, SCALE>, 1, true, 24, WalkerSub<8, VALUED, MAXR, FUSE, 4, SCALE>>(a);
}

// Any K, any 4-byte alignment, above the lane-group walkers' K <= 16: the ring walker on 4-byte slices -- 32 columns per
// pack, up to 4 packs per lane (128-column panels), 16 (V <= 2) or 8 rows per stage.
template <int V, bool VALUED, bool MAXR, bool FUSE, bool SCALE>
cudaError_t launch_ring1(const Args &a)
{
    constexpr int G = V <= 2 ? 16 : 8;  // 4-8 KB of ring per warp either way
    return launch<WalkerRing<V, VALUED, G, 2, 0, true, false, MAXR, FUSE, false, 1, SCALE>, V, false, 24>(a);
}
template <bool VALUED, bool MAXR, bool FUSE, bool SCALE = FUSE>
cudaError_t dispatch_ring1(int V, const Args &a)
{
    switch (V) {
        case 1: return launch_ring1<1, VALUED, MAXR, FUSE, SCALE>(a);
        case 2: return launch_ring1<2, VALUED, MAXR, FUSE, SCALE>(a);
        case 3: return launch_ring1<3, VALUED, MAXR, FUSE, SCALE>(a);
        default: return launch_ring1<4, VALUED, MAXR, FUSE, SCALE>(a);
    }
}

// Scalar instantiations of the register walker: any K, any alignment (GESPMM_WALKER_REGISTER; comparisons).
template <bool VALUED, bool MAXR, bool FUSE>
cudaError_t dispatch_scalar(int V, const Args &a)
{
    switch (V) {
        case 1: return launch_reg<1, VALUED, false, 8, 16, MAXR, FUSE>(a);
        case 2: return launch_reg<2, VALUED, false, 4, 16, MAXR, FUSE>(a);
        case 3: return launch_reg<3, VALUED, false, 4, 16, MAXR, FUSE>(a);
        default: return launch_reg<4, VALUED, false, 4, 16, MAXR, FUSE>(a);
    }
}

// The register-staged walker on aligned operands (GESPMM_WALKER_REGISTER: comparisons, and B that lives in the L2)
template <bool VALUED>
cudaError_t dispatch_reg4(int V, const Args &a)
{
    switch (V) {
        case 1: return launch_reg<1, VALUED, true, 8, 24>(a);
        case 2: return launch_reg<2, VALUED, true, 4, 24>(a);
        case 3: return launch_reg<3, VALUED, true, 2, 16>(a);
        default: return launch_reg<4, VALUED, true, 2, 16>(a);
    }
}

// ---- the families, one translation unit each ---------------------------------------------------------------------------
// mode: 0 sum, 1 max, 2 fused sum with a gathered-row scale (col_scale), 3 fused sum without one (row scale / bias only)
cudaError_t run_ring_valued(int mode, bool peer, bool hint, int V, bool masked, const Args &a);    // gespmm_spmm_ring_valued.cu
cudaError_t run_ring_unvalued(int mode, bool peer, bool hint, int V, bool masked, const Args &a);  // gespmm_spmm_ring_unvalued.cu
cudaError_t run_narrow(int mode, bool valued, bool rows, int K, const Args &a);                    // gespmm_spmm_narrow.cu: K <= 64, 16-byte slices
cudaError_t run_lanegroup(int mode, bool valued, bool rowgroup, int K, const Args &a);             // gespmm_spmm_lanegroup.cu: K <= 16, 4-byte slices
cudaError_t run_scalar(int mode, bool valued, bool reg, int V, const Args &a);                     // gespmm_spmm_scalar.cu: any K, 4-byte slices / registers
cudaError_t run_other(bool valued, bool bulk, bool hint, int V, bool masked, const Args &a);       // gespmm_spmm_other.cu: TMA bulk walker, register walker (vec4)

}  // namespace gespmm_detail

#endif  // GESPMM_SPMM_KERNELS_CUH
