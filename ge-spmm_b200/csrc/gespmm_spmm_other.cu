// gespmm_spmm_other.cu -- the TMA bulk-copy walker (cp.async.bulk + mbarrier) and the register-staged walker on aligned operands
#include "gespmm_spmm_kernels.cuh"

namespace gespmm_detail {

template <bool VALUED>
cudaError_t other(bool bulk, bool hint, int V, bool masked, const Args &a)
{
    if (!bulk) return dispatch_reg4<VALUED>(V, a);
    if (hint) return masked ? launch<WalkerBulk<VALUED, true, true>, 1, true, 24>(a) : launch<WalkerBulk<VALUED, false, true>, 1, true, 24>(a);
    return masked ? launch<WalkerBulk<VALUED, true, false>, 1, true, 24>(a) : launch<WalkerBulk<VALUED, false, false>, 1, true, 24>(a);
}
cudaError_t run_other(bool valued, bool bulk, bool hint, int V, bool masked, const Args &a)
{
    return valued ? other<true>(bulk, hint, V, masked, a) : other<false>(bulk, hint, V, masked, a);
}

}  // namespace gespmm_detail
