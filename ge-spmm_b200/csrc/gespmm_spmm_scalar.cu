// gespmm_spmm_scalar.cu -- any K, any 4-byte alignment: the ring walker on 4-byte slices, and the register-staged scalar walker (comparisons)
#include "gespmm_spmm_kernels.cuh"

namespace gespmm_detail {

template <bool VALUED>
cudaError_t scalar(int mode, bool reg, int V, const Args &a)
{
    if (reg) {
        if (mode == 1) return dispatch_scalar<VALUED, true, false>(V, a);
        if (mode >= 2) return dispatch_scalar<VALUED, false, true>(V, a);
        return dispatch_scalar<VALUED, false, false>(V, a);
    }
    if (mode == 1) return dispatch_ring1<VALUED, true, false>(V, a);
    if (mode == 2) return dispatch_ring1<VALUED, false, true>(V, a);
    if (mode == 3) return dispatch_ring1<VALUED, false, true, false>(V, a);
    return dispatch_ring1<VALUED, false, false>(V, a);
}
cudaError_t run_scalar(int mode, bool valued, bool reg, int V, const Args &a)
{
    return valued ? scalar<true>(mode, reg, V, a) : scalar<false>(mode, reg, V, a);
}

}  // namespace gespmm_detail
