// oracle/_ref/libref_cli_kernels.so -- the reference CLI's templated kernels
// spmm_test0..4<float> and spmmWrapper (/root/reference/spmm_test.cu:64-492), compiled
// from where they lie: the whole translation unit is included with its main() renamed,
// and spmmWrapper is exposed through a C entry point.  Pointers are DEVICE pointers.
// TEST INFRASTRUCTURE ONLY (GPU-side oracle + "reference kernel on the same B200" timing).
#define main ref_cli_main
#include "spmm_test.cu"
#undef main

// method 0..4 = spmm_test0..4, tile_row = rows per block (the CLI times method 2, tile_row 8:
// spmm_test.cu:724,756).  Launches on the legacy default stream like the reference; returns
// cudaGetLastError() so the caller can see launch failures the reference never checks.
extern "C" int ref_spmm_wrapper(int method, int tile_row, int M, int K, int *rowptr, int *colind,
                                float *val, float *B, float *C)
{
    spmmWrapper(method, tile_row, M, K, rowptr, colind, val, B, C);
    return (int)cudaGetLastError();
}

// Time `iters` back-to-back launches the way the CLI does (spmm_test.cu:754-760); ms per launch.
extern "C" float ref_spmm_time_ms(int method, int tile_row, int M, int K, int *rowptr, int *colind,
                                  float *val, float *B, float *C, int warmup, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < warmup; i++) spmmWrapper(method, tile_row, M, K, rowptr, colind, val, B, C);
    cudaEventRecord(e0, 0);
    for (int i = 0; i < iters; i++) spmmWrapper(method, tile_row, M, K, rowptr, colind, val, B, C);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (cudaGetLastError() != cudaSuccess) return -1.f;
    return ms / iters;
}
