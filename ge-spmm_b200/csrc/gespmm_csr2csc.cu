// gespmm_csr2csc.cu -- CSR -> CSC (CSR of the transpose) on the device, without cuSPARSE.
//
// Replaces csr2cscKernel / csr2csc_cuda (pytorch-custom/spmm_kernel.cu:381-423, 460-477),
// which call cusparseCsr2cscEx2 through a handle that is never created.  The consumer is
// SPMMFunction.backward (op.py:20-36): grad_feat = spmm(colptr, rowind, [csc_val], grad_out).
//
// Method: entries enumerated in CSR order are already sorted by row, so a STABLE sort by
// column yields (column, row) order -- the order cusparseCsr2cscEx2 / scipy tocsc() give.
// The stable sort is an LSD radix sort on the column index, 8 bits per pass, over
// (key, original position) pairs; only ceil(log2(N)/8) passes are run.  Each warp owns a
// contiguous chunk of kChunk entries, so ranking needs only __match_any_sync and a per-warp
// digit table -- no atomics anywhere, the result is deterministic.
//   pass:   hist    per-chunk digit counts            -> table[digit][chunk]
//           scan    exclusive scan of the table       (3-kernel scan, below)
//           scatter re-read chunk, rank, write (key, pos) to the other buffer
//   finish: colptr from the boundaries of the sorted keys; rowind / csc_val gathered through pos.
// Not on the timed path (once per graph), so written for clarity and determinism, not speed.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gespmm.h"

namespace {

constexpr int kChunk = 2048;     // entries per warp
constexpr int kWarpsPerCta = 8;
constexpr int kScanTile = 2048;  // 256 threads x 8
constexpr unsigned kFull = 0xffffffffu;

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- exclusive scan of int32, n up to 2^31 -----------------------------------------------------
__global__ void scan_tile_sums(const int *__restrict__ in, int *__restrict__ sums, long long n)
{
    __shared__ int red[8];
    const long long base = (long long)blockIdx.x * kScanTile;
    int s = 0;
    for (int i = threadIdx.x; i < kScanTile; i += 256) {
        const long long j = base + i;
        if (j < n) s += in[j];
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; w++) t += red[w];
        sums[blockIdx.x] = t;
    }
}

// one block: in-place exclusive scan of `sums` (n small: number of tiles)
__global__ void scan_sums(int *__restrict__ sums, long long n)
{
    __shared__ int wsum[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 1024) {
        const long long j = base + threadIdx.x;
        const int x = j < n ? sums[j] : 0;
        int incl = x;
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(kFull, incl, o);
            if ((threadIdx.x & 31) >= o) incl += y;
        }
        if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = wsum[threadIdx.x], wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(kFull, wi, o);
                if (threadIdx.x >= o) wi += y;
            }
            wsum[threadIdx.x] = wi - w;  // exclusive prefix of warp sums
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + wsum[threadIdx.x >> 5] + incl - x;
        if (j < n) sums[j] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + x;
        __syncthreads();
    }
}

__global__ void scan_tiles(const int *in, int *out, const int *__restrict__ sums, long long n)  // in may alias out
{
    __shared__ int wsum[8];
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * 8;
    int x[8], t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = (base + i < n) ? in[base + i] : 0; t += x[i]; }
    int incl = t;
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(kFull, incl, o);
        if ((threadIdx.x & 31) >= o) incl += y;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    int off = sums[blockIdx.x] + incl - t;
    for (int w = 0; w < (threadIdx.x >> 5); w++) off += wsum[w];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (base + i < n) out[base + i] = off;
        off += x[i];
    }
}

cudaError_t exclusive_scan(const int *in, int *out, long long n, int *tile_sums, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const long long tiles = (n + kScanTile - 1) / kScanTile;
    scan_tile_sums<<<(unsigned)tiles, 256, 0, st>>>(in, tile_sums, n);
    scan_sums<<<1, 1024, 0, st>>>(tile_sums, tiles);
    scan_tiles<<<(unsigned)tiles, 256, 0, st>>>(in, out, tile_sums, n);
    return cudaGetLastError();
}

// ---- radix pass ----------------------------------------------------------------------------------
// FIRST: keys come from colind and the payload is the position itself.
template <bool FIRST>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
radix_hist(const int *__restrict__ keys, long long nnz, int shift, long long nchunks, int *__restrict__ table)
{
    __shared__ int cnt[kWarpsPerCta][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long chunk = (long long)blockIdx.x * kWarpsPerCta + warp;
    for (int d = lane; d < 256; d += 32) cnt[warp][d] = 0;
    __syncwarp();
    if (chunk >= nchunks) return;
    const long long b = chunk * kChunk;
    const long long e = (b + kChunk < nnz) ? b + kChunk : nnz;
    for (long long p0 = b; p0 < e; p0 += 32) {
        const long long p = p0 + lane;
        const bool ok = p < e;
        const int digit = ok ? ((keys[p] >> shift) & 255) : 256 + lane;  // inactive lanes match nobody
        const unsigned peers = __match_any_sync(kFull, digit);
        if (ok && (__ffs(peers) - 1) == lane) cnt[warp][digit] += __popc(peers);
        __syncwarp();
    }
    for (int d = lane; d < 256; d += 32) table[(long long)d * nchunks + chunk] = cnt[warp][d];
}

template <bool FIRST>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
radix_scatter(const int *__restrict__ keys_in, const int *__restrict__ pos_in, int *__restrict__ keys_out,
              int *__restrict__ pos_out, long long nnz, int shift, long long nchunks, const int *__restrict__ table)
{
    __shared__ int base[kWarpsPerCta][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long chunk = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (chunk >= nchunks) return;
    for (int d = lane; d < 256; d += 32) base[warp][d] = table[(long long)d * nchunks + chunk];
    __syncwarp();
    const long long b = chunk * kChunk;
    const long long e = (b + kChunk < nnz) ? b + kChunk : nnz;
    for (long long p0 = b; p0 < e; p0 += 32) {
        const long long p = p0 + lane;
        const bool ok = p < e;
        const int key = ok ? keys_in[p] : 0;
        const int digit = ok ? ((key >> shift) & 255) : 256 + lane;
        const unsigned peers = __match_any_sync(kFull, digit);
        int dst = 0;
        if (ok) dst = base[warp][digit] + __popc(peers & ((1u << lane) - 1));
        __syncwarp();
        if (ok && (__ffs(peers) - 1) == lane) base[warp][digit] += __popc(peers);
        __syncwarp();
        if (ok) {
            keys_out[dst] = key;
            pos_out[dst] = FIRST ? (int)p : pos_in[p];
        }
    }
}

// rowid[p] = row of CSR position p
__global__ void expand_rows(const int *__restrict__ rowptr, int M, int *__restrict__ rowid)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long r0 = warp * 32;
    for (int i = 0; i < 32; i++) {
        const long long r = r0 + i;
        if (r >= M) return;
        const int s = rowptr[r], e = rowptr[r + 1];
        for (int p = s + lane; p < e; p += 32) rowid[p] = (int)r;
    }
}

// sorted keys -> colptr; sorted positions -> rowind, csc_val
__global__ void finish(const int *__restrict__ keys, const int *__restrict__ pos, const int *__restrict__ rowid,
                       const float *__restrict__ val, long long nnz, int N, int *__restrict__ colptr,
                       int *__restrict__ rowind, float *__restrict__ csc_val)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nnz) return;
    // colptr[c] = first sorted index whose key >= c
    const int prev = (i == 0) ? -1 : keys[i - 1];
    const int cur = (i == nnz) ? N : keys[i];
    for (int c = prev + 1; c <= cur; c++) colptr[c] = (int)i;
    if (i < nnz) {
        const int p = pos[i];
        rowind[i] = rowid[p];
        if (csc_val) csc_val[i] = val[p];
    }
}

struct Layout {
    size_t keyA, keyB, posA, posB, rowid, table, sums, total;
    long long nchunks;
};

Layout make_layout(int64_t nnz)
{
    Layout L;
    const size_t n4 = align_up((size_t)(nnz > 0 ? nnz : 1) * 4, 256);
    L.nchunks = (nnz + kChunk - 1) / kChunk;
    const size_t tab = (size_t)256 * (size_t)(L.nchunks > 0 ? L.nchunks : 1);
    size_t off = 0;
    L.keyA = off; off += n4;
    L.keyB = off; off += n4;
    L.posA = off; off += n4;
    L.posB = off; off += n4;
    L.rowid = off; off += n4;
    L.table = off; off += align_up(tab * 4, 256);
    L.sums = off; off += align_up(((tab + kScanTile - 1) / kScanTile + 1) * 4, 256);
    L.total = off;
    return L;
}

}  // namespace

extern "C" size_t gespmm_csr2csc_workspace_bytes(int64_t M, int64_t N, int64_t nnz)
{
    (void)M; (void)N;
    if (nnz < 0) return 0;
    return make_layout(nnz).total;
}

extern "C" int gespmm_csr2csc_f32(int64_t M, int64_t N, int64_t nnz, const int32_t *rowptr, const int32_t *colind,
                                  const float *val, int32_t *colptr, int32_t *rowind, float *csc_val,
                                  void *workspace, size_t workspace_bytes, void *stream)
{
    if (M < 0 || N < 0 || nnz < 0) return GESPMM_ERR_INVALID_ARG;
    if (M > INT32_MAX - 1 || N > INT32_MAX - 1 || nnz > INT32_MAX - 64) return GESPMM_ERR_TOO_LARGE;
    if (!rowptr || !colptr) return GESPMM_ERR_INVALID_ARG;
    if (nnz > 0 && (!colind || !rowind)) return GESPMM_ERR_INVALID_ARG;
    if ((val == nullptr) != (csc_val == nullptr)) return GESPMM_ERR_INVALID_ARG;
    const Layout L = make_layout(nnz);
    if (!workspace || workspace_bytes < L.total) return GESPMM_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char *ws = static_cast<char *>(workspace);
    int *keyA = (int *)(ws + L.keyA), *keyB = (int *)(ws + L.keyB);
    int *posA = (int *)(ws + L.posA), *posB = (int *)(ws + L.posB);
    int *rowid = (int *)(ws + L.rowid), *table = (int *)(ws + L.table), *sums = (int *)(ws + L.sums);

    if (nnz > 0) {
        const long long warps = (M + 31) / 32;
        expand_rows<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(rowptr, (int)M, rowid);
        int bits = 1;
        while (bits < 31 && ((int64_t)1 << bits) < N) bits++;
        const int passes = (bits + 7) / 8;
        const unsigned grid = (unsigned)((L.nchunks + kWarpsPerCta - 1) / kWarpsPerCta);
        const long long tab = 256LL * L.nchunks;
        const int *kin = colind;
        const int *pin = nullptr;
        int *kout = keyA, *pout = posA;
        for (int ps = 0; ps < passes; ps++) {
            const int shift = 8 * ps;
            if (ps == 0) radix_hist<true><<<grid, kWarpsPerCta * 32, 0, st>>>(kin, nnz, shift, L.nchunks, table);
            else radix_hist<false><<<grid, kWarpsPerCta * 32, 0, st>>>(kin, nnz, shift, L.nchunks, table);
            if (exclusive_scan(table, table, tab, sums, st) != cudaSuccess) return GESPMM_ERR_CUDA;
            if (ps == 0) radix_scatter<true><<<grid, kWarpsPerCta * 32, 0, st>>>(kin, pin, kout, pout, nnz, shift, L.nchunks, table);
            else radix_scatter<false><<<grid, kWarpsPerCta * 32, 0, st>>>(kin, pin, kout, pout, nnz, shift, L.nchunks, table);
            kin = kout; pin = pout;
            kout = (kout == keyA) ? keyB : keyA;
            pout = (pout == posA) ? posB : posA;
        }
        finish<<<(unsigned)((nnz + 1 + 255) / 256), 256, 0, st>>>(kin, pin, rowid, val, nnz, (int)N, colptr, rowind, csc_val);
    } else {
        if (cudaMemsetAsync(colptr, 0, (size_t)(N + 1) * 4, st) != cudaSuccess) return GESPMM_ERR_CUDA;
    }
    return cudaGetLastError() == cudaSuccess ? GESPMM_OK : GESPMM_ERR_CUDA;
}
