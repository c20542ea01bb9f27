// gespmm_spmm_lanegroup.cu -- K <= 16 in 4-byte slices: sub-warp walker (default) and row-group kernel (sequential order)
#include "gespmm_spmm_kernels.cuh"

namespace gespmm_detail {

template <bool VALUED>
cudaError_t lanegroup(int mode, bool rowgroup, int K, const Args &a)
{
    if (rowgroup) {
        if (mode == 1) return dispatch_rowgroup<VALUED, true, false>(K, a);
        if (mode >= 2) return dispatch_rowgroup<VALUED, false, true>(K, a);
        return dispatch_rowgroup<VALUED, false, false>(K, a);
    }
    if (mode == 1) return dispatch_sub1<VALUED, true, false>(K, a);
    if (mode == 2) return dispatch_sub1<VALUED, false, true>(K, a);
    if (mode == 3) return dispatch_sub1<VALUED, false, true, false>(K, a);
    return dispatch_sub1<VALUED, false, false>(K, a);
}
cudaError_t run_lanegroup(int mode, bool valued, bool rowgroup, int K, const Args &a)
{
    return valued ? lanegroup<true>(mode, rowgroup, K, a) : lanegroup<false>(mode, rowgroup, K, a);
}

}  // namespace gespmm_detail
