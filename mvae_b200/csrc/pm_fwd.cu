// pm_fwd.cu — forward instantiations of the fused product-manifold kernel (see pm_kernels_impl.cuh).
#define MVAE_PM_BWD 0
#include "pm_kernels_impl.cuh"

namespace mvae {
int launch_pm_forward(PmParams& p, void* stream) { return launch_pm(p, stream); }
}  // namespace mvae
