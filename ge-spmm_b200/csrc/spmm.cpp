// spmm.cpp -- PyTorch extension module `spmm`, the operator surface of the reference
// (pytorch-custom/spmm.cpp:96-101): csr_spmm, csr_spmm_no_edge_value, csr2csc with the same
// names, arity and argument meaning.  Thin shim over the C ABI in include/gespmm.h.
//
// Differences from the reference shim, all deliberate:
//   * argument predicates (CUDA device, contiguous, int32 / float32) are the reference's
//     (spmm.cpp:30-41, 50-58, 77-91) but raise a Python exception (TORCH_CHECK) instead of a
//     C assert() that aborts the process; shapes are checked too;
//   * launches go to the current PyTorch stream under a device guard (the reference uses the
//     legacy default stream and no guard: spmm_kernel.cu:189,196,203);
//   * csr2csc works (the reference's uses an uninitialised cuSPARSE handle: spmm_kernel.cu:386).
//   * summation order: rows of at most 4096 nonzeros are summed in the reference's sequential CSR order -- bit-identical
//     results -- for K > 64 and for K % 4 != 0 above 16; for K <= 64 (and K <= 16 of any parity) csr_spmm /
//     csr_spmm_no_edge_value use the faster walkers that re-associate a row's sum (deterministic, within ~2e-6 of max|out|;
//     the 1e-4 bar of BASELINE.json holds with two orders of magnitude to spare).  csr_spmm_ex(..., sequential=True) asks for
//     the reference's order at every K, per call; row_sum_is_sequential(K, row_nnz, sequential) tells which rows get it.
// There is no CPU path: CPU tensors are rejected.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "gespmm.h"

namespace {

void check_index(const torch::Tensor &t, const char *name)
{
    TORCH_CHECK(t.device().type() == torch::kCUDA, name, " must be a CUDA tensor (this build has no CPU path)");
    TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
    TORCH_CHECK(t.dtype() == torch::kInt32, name, " must be int32");
    TORCH_CHECK(t.dim() == 1, name, " must be 1-D");
}

void check_float(const torch::Tensor &t, const char *name, int dim)
{
    TORCH_CHECK(t.device().type() == torch::kCUDA, name, " must be a CUDA tensor (this build has no CPU path)");
    TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
    TORCH_CHECK(t.dtype() == torch::kFloat32, name, " must be float32");
    TORCH_CHECK(t.dim() == dim, name, " must be ", dim, "-D");
}

torch::Tensor run(const torch::Tensor &rowptr, const torch::Tensor &colind, const float *val, const torch::Tensor &B,
                  const gespmm_opts *opts = nullptr)
{
    TORCH_CHECK(rowptr.size(0) >= 1, "A_rowptr must have at least one element");
    TORCH_CHECK(rowptr.device() == B.device() && colind.device() == B.device(), "all tensors must be on B's device");
    const int64_t M = rowptr.size(0) - 1, N = B.size(0), K = B.size(1), nnz = colind.size(0);
    c10::cuda::CUDAGuard guard(B.device());
    auto out = torch::empty({M, K}, B.options());  // spmm_kernel.cu:182-184
    cudaStream_t stream = at::cuda::getCurrentCUDAStream();
    // widths that are not multiples of 4 (K > 16): the library runs them on its 16-byte-slice walkers through padded
    // copies of B and C when it is handed scratch memory; the caching allocator makes that free after the first call
    gespmm_opts with_ws;
    torch::Tensor ws;
    const size_t ws_bytes = gespmm_pad_workspace_bytes(M, N, K, nnz);
    if (ws_bytes > 0) {
        if (opts) with_ws = *opts; else gespmm_opts_init(&with_ws);
        ws = torch::empty({(int64_t)ws_bytes}, B.options().dtype(torch::kUInt8));
        with_ws.workspace = ws.data_ptr();
        with_ws.workspace_bytes = ws_bytes;
        opts = &with_ws;
    }
    const int rc = gespmm_csr_spmm_f32_ex(M, N, K, nnz, rowptr.data_ptr<int>(), colind.data_ptr<int>(), val,
                                          B.data_ptr<float>(), K, out.data_ptr<float>(), K, opts, stream);
    TORCH_CHECK(rc == GESPMM_OK, "gespmm_csr_spmm_f32 failed: ", gespmm_error_string(rc));
    return out;
}

const float *optional_vector(const c10::optional<torch::Tensor> &t, const char *name, int64_t len, const torch::Tensor &B)
{
    if (!t.has_value() || !t->defined()) return nullptr;
    check_float(*t, name, t->dim());
    TORCH_CHECK(t->numel() == len, name, " must have ", len, " elements");
    TORCH_CHECK(t->device() == B.device(), "all tensors must be on B's device");
    return t->data_ptr<float>();
}

}  // namespace

// pytorch-custom/spmm.cpp:24-43
torch::Tensor csr_spmm(torch::Tensor A_rowptr, torch::Tensor A_colind, torch::Tensor A_csrVal, torch::Tensor B)
{
    check_index(A_rowptr, "A_rowptr");
    check_index(A_colind, "A_colind");
    check_float(A_csrVal, "A_csrVal", 1);
    check_float(B, "B", 2);
    TORCH_CHECK(A_csrVal.size(0) == A_colind.size(0), "A_csrVal and A_colind must have the same length");
    TORCH_CHECK(A_csrVal.device() == B.device(), "all tensors must be on B's device");
    return run(A_rowptr, A_colind, A_csrVal.data_ptr<float>(), B);
}

// pytorch-custom/spmm.cpp:45-60
torch::Tensor csr_spmm_no_edge_value(torch::Tensor A_rowptr, torch::Tensor A_colind, torch::Tensor B)
{
    check_index(A_rowptr, "A_rowptr");
    check_index(A_colind, "A_colind");
    check_float(B, "B", 2);
    return run(A_rowptr, A_colind, nullptr, B);
}

// pytorch-custom/spmm.cpp:70-93: fills colptr / rowind in place, returns the CSC values.
torch::Tensor csr2csc(torch::Tensor rowptr, torch::Tensor colind, torch::Tensor colptr, torch::Tensor rowind,
                      torch::Tensor csr_data)
{
    check_index(rowptr, "rowptr");
    check_index(colind, "colind");
    check_index(colptr, "colptr");
    check_index(rowind, "rowind");
    check_float(csr_data, "csr_data", 1);
    const int64_t M = rowptr.size(0) - 1, N = colptr.size(0) - 1, nnz = colind.size(0);
    TORCH_CHECK(M >= 0 && N >= 0, "rowptr and colptr must be non-empty");
    TORCH_CHECK(rowind.size(0) == nnz && csr_data.size(0) == nnz, "rowind / csr_data must have nnz elements");
    c10::cuda::CUDAGuard guard(rowptr.device());
    auto csc_val = torch::empty({nnz}, csr_data.options());  // spmm_kernel.cu:471-473
    const size_t ws_bytes = gespmm_csr2csc_workspace_bytes(M, N, nnz);
    auto ws = torch::empty({(int64_t)ws_bytes}, rowptr.options().dtype(torch::kUInt8));
    cudaStream_t stream = at::cuda::getCurrentCUDAStream();
    const int rc = gespmm_csr2csc_f32(M, N, nnz, rowptr.data_ptr<int>(), colind.data_ptr<int>(),
                                      csr_data.data_ptr<float>(), colptr.data_ptr<int>(), rowind.data_ptr<int>(),
                                      csc_val.data_ptr<float>(), ws.data_ptr(), ws_bytes, stream);
    TORCH_CHECK(rc == GESPMM_OK, "gespmm_csr2csc_f32 failed: ", gespmm_error_string(rc));
    return csc_val;
}

// New (no reference counterpart): the product with per-call options -- summation order, the graph's longest row
// (lets the call skip the long-row kernel), and GCNConv's normalisation / bias passes (pytorch-custom/op.py:142-147)
// fused into the kernel:  out = (A @ (B * col_scale[:, None])) * row_scale[:, None] + bias, bit-identical to the
// four separate passes.  A_csrVal / row_scale / col_scale / bias may be None.
torch::Tensor csr_spmm_ex(torch::Tensor A_rowptr, torch::Tensor A_colind, c10::optional<torch::Tensor> A_csrVal,
                          torch::Tensor B, bool sequential, int64_t max_row_nnz, c10::optional<torch::Tensor> row_scale,
                          c10::optional<torch::Tensor> col_scale, c10::optional<torch::Tensor> bias)
{
    check_index(A_rowptr, "A_rowptr");
    check_index(A_colind, "A_colind");
    check_float(B, "B", 2);
    const float *val = nullptr;
    if (A_csrVal.has_value() && A_csrVal->defined()) {
        check_float(*A_csrVal, "A_csrVal", 1);
        TORCH_CHECK(A_csrVal->size(0) == A_colind.size(0), "A_csrVal and A_colind must have the same length");
        TORCH_CHECK(A_csrVal->device() == B.device(), "all tensors must be on B's device");
        val = A_csrVal->data_ptr<float>();
    }
    gespmm_opts o;
    gespmm_opts_init(&o);
    if (sequential) o.flags |= GESPMM_FLAG_SEQUENTIAL;
    o.max_row_nnz = max_row_nnz;
    o.row_scale = optional_vector(row_scale, "row_scale", A_rowptr.size(0) - 1, B);
    o.col_scale = optional_vector(col_scale, "col_scale", B.size(0), B);
    o.bias = optional_vector(bias, "bias", B.size(1), B);
    return run(A_rowptr, A_colind, val, B, &o);
}

// Whether this operator sums a row of `row_nnz` nonzeros at width K in the reference's sequential order
// (gespmm_row_sum_is_sequential_ex for the options csr_spmm / csr_spmm_ex pass).
bool row_sum_is_sequential(int64_t K, int64_t row_nnz, bool sequential)
{
    gespmm_opts o;
    gespmm_opts_init(&o);
    if (sequential) o.flags |= GESPMM_FLAG_SEQUENTIAL;
    return gespmm_row_sum_is_sequential_ex(K, row_nnz, &o) != 0;
}

// Longest row of a CSR on the device (one small kernel + a 4-byte copy back; synchronises the current stream).
int64_t max_row_nnz(torch::Tensor rowptr)
{
    check_index(rowptr, "rowptr");
    TORCH_CHECK(rowptr.size(0) >= 1, "rowptr must have at least one element");
    c10::cuda::CUDAGuard guard(rowptr.device());
    int32_t out = 0;
    const int rc = gespmm_max_row_nnz(rowptr.size(0) - 1, rowptr.data_ptr<int>(), &out, at::cuda::getCurrentCUDAStream());
    TORCH_CHECK(rc == GESPMM_OK, "gespmm_max_row_nnz failed: ", gespmm_error_string(rc));
    return out;
}

PYBIND11_MODULE(spmm, m)
{
    m.doc() = "spmm in CSR format. csr_spmm is the kernel with edge value. csr2csc provides the format transformation";
    m.def("csr_spmm", &csr_spmm, "CSR SPMM");
    m.def("csr_spmm_no_edge_value", &csr_spmm_no_edge_value, "CSR SPMM NO EDGE VALUE");
    m.def("csr2csc", &csr2csc, "csr2csc");
    m.def("csr_spmm_ex", &csr_spmm_ex, "CSR SPMM with per-call options and fused row / column scaling and bias",
          pybind11::arg("A_rowptr"), pybind11::arg("A_colind"), pybind11::arg("A_csrVal"), pybind11::arg("B"),
          pybind11::arg("sequential") = false, pybind11::arg("max_row_nnz") = -1, pybind11::arg("row_scale") = pybind11::none(),
          pybind11::arg("col_scale") = pybind11::none(), pybind11::arg("bias") = pybind11::none());
    m.def("max_row_nnz", &max_row_nnz, "longest row of a device CSR");
    m.def("row_sum_is_sequential", &row_sum_is_sequential, "does this operator sum such a row in CSR order (bit-identical to the reference)?",
          pybind11::arg("K"), pybind11::arg("row_nnz"), pybind11::arg("sequential") = false);
}
