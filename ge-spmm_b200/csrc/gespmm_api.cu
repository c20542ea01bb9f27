// gespmm_api.cu -- the small non-kernel entry points of the C ABI (include/gespmm.h).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gespmm.h"

extern "C" int gespmm_version(void) { return 200; /* 0.2.0 */ }

extern "C" const char *gespmm_error_string(int code)
{
    switch (code) {
        case GESPMM_OK: return "success";
        case GESPMM_ERR_INVALID_ARG: return "invalid argument (null pointer, negative size or leading dimension < K)";
        case GESPMM_ERR_CUDA: return "CUDA runtime call or kernel launch failed (no usable CUDA device?)";
        case GESPMM_ERR_TOO_LARGE: return "M, N or nnz exceeds the int32 index contract";
        case GESPMM_ERR_IO: return "cannot read file / not a MatrixMarket coordinate file";
        case GESPMM_ERR_WORKSPACE: return "workspace too small";
        case GESPMM_ERR_NOMEM: return "host allocation failed";
        default: return "unknown gespmm error code";
    }
}

extern "C" void gespmm_free_host(void *p) { free(p); }

// Let kernels on the current device read memory of `peer_device` (NVLink / PCIe P2P).  Needed before
// gespmm_csr_spmm_f32_bparts is given blocks of B that live on other GPUs.
extern "C" int gespmm_enable_peer_access(int peer_device)
{
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) return GESPMM_ERR_CUDA;
    if (peer_device == cur) return GESPMM_OK;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, cur, peer_device) != cudaSuccess || !can) { cudaGetLastError(); return GESPMM_ERR_CUDA; }
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return GESPMM_OK; }
    return e == cudaSuccess ? GESPMM_OK : GESPMM_ERR_CUDA;
}

// Map another process's device allocation (cudaIpcMemHandle_t, 64 bytes) into this process for kernels
// on the CURRENT device; peer access to the owning GPU is enabled as part of the mapping.
extern "C" int gespmm_ipc_open(const unsigned char *handle64, void **base)
{
    if (!handle64 || !base) return GESPMM_ERR_INVALID_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    *base = nullptr;
    if (cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return GESPMM_ERR_CUDA; }
    return GESPMM_OK;
}

// A device allocation that other processes can map: plain cudaMalloc (framework caching allocators may
// hand out virtual-memory segments that cudaIpc* cannot export) plus its 64-byte IPC handle.
extern "C" int gespmm_ipc_alloc(size_t bytes, void **dptr, unsigned char *handle64)
{
    if (!dptr || !handle64) return GESPMM_ERR_INVALID_ARG;
    *dptr = nullptr;
    if (cudaMalloc(dptr, bytes ? bytes : 16) != cudaSuccess) { cudaGetLastError(); return GESPMM_ERR_CUDA; }
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, *dptr) != cudaSuccess) { cudaGetLastError(); cudaFree(*dptr); *dptr = nullptr; return GESPMM_ERR_CUDA; }
    memcpy(handle64, &h, sizeof(h));
    return GESPMM_OK;
}

extern "C" int gespmm_ipc_free(void *dptr)
{
    if (dptr && cudaFree(dptr) != cudaSuccess) { cudaGetLastError(); return GESPMM_ERR_CUDA; }
    return GESPMM_OK;
}

extern "C" int gespmm_ipc_close(void *base)
{
    if (!base) return GESPMM_OK;
    if (cudaIpcCloseMemHandle(base) != cudaSuccess) { cudaGetLastError(); return GESPMM_ERR_CUDA; }
    return GESPMM_OK;
}

// Host-buffer convenience call: what the reference CLI does by hand around its launches
// (cudaMalloc x5 + cudaMemcpy H2D, spmm_test.cu:609-640; D2H under VALIDATE, :689).
extern "C" int gespmm_csr_spmm_f32_host(int64_t M, int64_t N, int64_t K, int64_t nnz, const int32_t *rowptr,
                                        const int32_t *colind, const float *val, const float *B, int64_t ldb,
                                        float *C, int64_t ldc, int device)
{
    if (M < 0 || N < 0 || K < 0 || nnz < 0 || ldb < K || ldc < K) return GESPMM_ERR_INVALID_ARG;
    if (M == 0 || K == 0) return GESPMM_OK;
    if (!rowptr || !C || (nnz > 0 && (!colind || !B))) return GESPMM_ERR_INVALID_ARG;
    int caller_device = -1;
    if (cudaGetDevice(&caller_device) != cudaSuccess) { cudaGetLastError(); return GESPMM_ERR_CUDA; }
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return GESPMM_ERR_CUDA; }

    int32_t *d_rowptr = nullptr, *d_colind = nullptr;
    float *d_val = nullptr, *d_B = nullptr, *d_C = nullptr;
    cudaStream_t st = nullptr;
    int rc = GESPMM_ERR_CUDA;
    const size_t nB = (size_t)N * (size_t)K, nC = (size_t)M * (size_t)K;
    do {
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) break;
        if (cudaMalloc(&d_rowptr, (size_t)(M + 1) * 4) != cudaSuccess) break;
        if (cudaMalloc(&d_colind, (size_t)(nnz > 0 ? nnz : 1) * 4) != cudaSuccess) break;
        if (val && cudaMalloc(&d_val, (size_t)(nnz > 0 ? nnz : 1) * 4) != cudaSuccess) break;
        if (cudaMalloc(&d_B, (nB > 0 ? nB : 1) * 4) != cudaSuccess) break;
        if (cudaMalloc(&d_C, nC * 4) != cudaSuccess) break;
        if (cudaMemcpyAsync(d_rowptr, rowptr, (size_t)(M + 1) * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
        if (nnz > 0 && cudaMemcpyAsync(d_colind, colind, (size_t)nnz * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
        if (nnz > 0 && val && cudaMemcpyAsync(d_val, val, (size_t)nnz * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) break;
        // dense operands are packed to stride K on the device
        if (nB > 0 && B && cudaMemcpy2DAsync(d_B, (size_t)K * 4, B, (size_t)ldb * 4, (size_t)K * 4, (size_t)N,
                                        cudaMemcpyHostToDevice, st) != cudaSuccess) break;
        rc = gespmm_csr_spmm_f32(M, N, K, nnz, d_rowptr, d_colind, d_val, d_B, K, d_C, K, st);
        if (rc != GESPMM_OK) break;
        rc = GESPMM_ERR_CUDA;
        if (cudaMemcpy2DAsync(C, (size_t)ldc * 4, d_C, (size_t)K * 4, (size_t)K * 4, (size_t)M,
                              cudaMemcpyDeviceToHost, st) != cudaSuccess) break;
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        rc = GESPMM_OK;
    } while (0);
    if (rc != GESPMM_OK) cudaGetLastError();
    cudaFree(d_rowptr); cudaFree(d_colind); cudaFree(d_val); cudaFree(d_B); cudaFree(d_C);
    if (st) cudaStreamDestroy(st);
    cudaSetDevice(caller_device);  // the caller's current device is left as it was found
    return rc;
}
