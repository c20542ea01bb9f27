// gespmm_spmm_ring_unvalued.cu -- the ring walker (cp.async gather ring, 16-byte slices), unvalued products
#include "gespmm_spmm_kernels.cuh"

namespace gespmm_detail {

cudaError_t run_ring_unvalued(int mode, bool peer, bool hint, int V, bool masked, const Args &a)
{
    constexpr bool VALUED = false;
    if (peer) return dispatch_ring<VALUED, true, false>(V, a, masked);
    if (mode == 1) return dispatch_ring<VALUED, false, true>(V, a, masked);
    if (mode == 2) return launch_ring<1, VALUED, false, false, true, false, true>(a, masked);
    if (mode == 3) return launch_ring<1, VALUED, false, false, true, false, false>(a, masked);
    if (hint) return launch_ring<1, VALUED, false, false, false, true>(a, masked);
    return dispatch_ring<VALUED, false, false>(V, a, masked);
}

}  // namespace gespmm_detail
