// membench -- measured ceilings for the SpMM roofline on this GPU (not part of the product path):
//   copy      : float4 stream copy, read+write bytes / time              (HBM ceiling)
//   gather S  : every warp reads random 512-byte rows (one LDG.128 per lane, 8 independent rows in
//               flight per warp) out of a table of S bytes and accumulates them; bytes gathered / time.
//               S below the L2 size gives the L2->SM gather ceiling, S far above it the DRAM
//               random-row ceiling.  These are the denominators for the high-degree (L2-resident)
//               and low-degree (DRAM-resident) regimes of SURVEY.md section 8d.
// Usage: membench [device]      prints one JSON object per line.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void copy_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// rows_per_warp random rows per warp, UNROLL independent loads in flight
template <int UNROLL>
__global__ void gather_kernel(const float4 *__restrict__ table, uint32_t nrows, int rows_per_warp, float4 *__restrict__ sink)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int i = 0; i < rows_per_warp; i += UNROLL) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const uint32_t r = hash32(warp * 0x9e3779b9U + (uint32_t)(i + u)) % nrows;
            v[u] = __ldg(table + (size_t)r * 32 + lane);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    if (acc.x == 123.456f) sink[warp * 32 + lane] = acc;  // never true: keeps the loads alive
}

int main(int argc, char **argv)
{
    const int dev = argc > 1 ? atoi(argv[1]) : 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, dev));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    {
        const size_t n = (size_t)1 << 27;  // 2 GiB per buffer
        float4 *a, *b;
        CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
        CK(cudaMemset(a, 1, n * 16));
        double best = 0;
        for (int it = 0; it < 6; it++) {
            CK(cudaEventRecord(e0));
            copy_kernel<<<p.multiProcessorCount * 16, 512>>>(a, b, n);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            const double gbs = 2.0 * n * 16 / ms / 1e6;
            if (it && gbs > best) best = gbs;
        }
        printf("{\"bench\": \"copy\", \"bytes\": %zu, \"gbs\": %.1f, \"sms\": %d, \"l2_bytes\": %d}\n", 2 * n * 16, best, p.multiProcessorCount, p.l2CacheSize);
        CK(cudaFree(a)); CK(cudaFree(b));
    }
    const size_t sizes[] = {(size_t)16 << 20, (size_t)32 << 20, (size_t)64 << 20, (size_t)96 << 20, (size_t)128 << 20,
                            (size_t)256 << 20, (size_t)1 << 30, (size_t)4 << 30};
    float4 *sink;
    CK(cudaMalloc(&sink, 16));
    for (size_t S : sizes) {
        float4 *t;
        CK(cudaMalloc(&t, S));
        CK(cudaMemset(t, 0, S));
        const uint32_t nrows = (uint32_t)(S / 512);
        const int warps_per_cta = 8, ctas = p.multiProcessorCount * 8 * 4, rows_per_warp = 2048;
        double best = 0;
        for (int it = 0; it < 4; it++) {
            CK(cudaEventRecord(e0));
            gather_kernel<8><<<ctas, warps_per_cta * 32>>>(t, nrows, rows_per_warp, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            const double gbs = (double)ctas * warps_per_cta * rows_per_warp * 512 / ms / 1e6;
            if (it && gbs > best) best = gbs;
        }
        printf("{\"bench\": \"gather512\", \"table_bytes\": %zu, \"gbs\": %.1f}\n", S, best);
        CK(cudaFree(t));
    }
    return 0;
}
