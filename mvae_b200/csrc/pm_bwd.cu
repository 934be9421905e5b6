// pm_bwd.cu — backward instantiations of the fused product-manifold kernel (see pm_kernels_impl.cuh).
#define MVAE_PM_BWD 1
// Register budgets of the backward instantiations (pm_kernels_impl.cuh: blocks of 256 threads per SM the launch bounds
// promise).  Dimensions <= 2: 3 blocks = 80 registers, not the header's 4 = 64.  At 64 the kernel spilled (224 bytes of
// stores, 284 of loads per thread: 126 local-memory instructions spread over all its loops, in a kernel that is bound by
// instruction issue), and the tighter budget bought nothing: the h2,s2,e2 launch runs 5 warps per CTA and shared memory
// caps it at 5 CTAs per SM — 25 warps x 80 registers x 32 lanes = 64 000 of the SM's 65 536 registers.
#define MVAE_PM_MINB(MAXN, BWD) ((MAXN) == 2 ? 3 : (MAXN) == 4 ? 3 : (MAXN) == 8 ? 2 : 1)
#include "pm_kernels_impl.cuh"

namespace mvae {
int launch_pm_backward(PmParams& p, void* stream) { return launch_pm(p, stream); }
}  // namespace mvae
