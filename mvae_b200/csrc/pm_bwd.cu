// pm_bwd.cu — backward instantiations of the fused product-manifold kernel (see pm_kernels_impl.cuh).
#define MVAE_PM_BWD 1
#include "pm_kernels_impl.cuh"

namespace mvae {
int launch_pm_backward(PmParams& p, void* stream) { return launch_pm(p, stream); }
}  // namespace mvae
