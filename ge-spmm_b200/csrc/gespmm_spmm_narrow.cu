// gespmm_spmm_narrow.cu -- narrow B in 16-byte slices (K <= 64): sub-warp walker (re-associated) and row-parallel walker (sequential)
#include "gespmm_spmm_kernels.cuh"

namespace gespmm_detail {

template <bool VALUED>
cudaError_t narrow(int mode, bool rows, int K, const Args &a)
{
    if (rows) {
        if (mode == 1) return dispatch_rows<VALUED, true, false>(K, a);
        if (mode == 2) return dispatch_rows<VALUED, false, true>(K, a);
        if (mode == 3) return dispatch_rows<VALUED, false, true, false>(K, a);
        return dispatch_rows<VALUED, false, false>(K, a);
    }
    if (mode == 1) return dispatch_sub<VALUED, true, false>(K, a);
    if (mode == 2) return dispatch_sub<VALUED, false, true>(K, a);
    if (mode == 3) return dispatch_sub<VALUED, false, true, false>(K, a);
    return dispatch_sub<VALUED, false, false>(K, a);
}
cudaError_t run_narrow(int mode, bool valued, bool rows, int K, const Args &a)
{
    return valued ? narrow<true>(mode, rows, K, a) : narrow<false>(mode, rows, K, a);
}

}  // namespace gespmm_detail
