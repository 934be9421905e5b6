// pm_params.cuh — launch parameters shared by the product-manifold translation units.
#pragma once
#include "mvae_common.cuh"

namespace mvae {


constexpr int kPmThreads = 128;

struct PmParams {
  mvae_pm_desc desc;
  int64_t B;
  const float* ml;
  const float* eps;
  const float* radius;
  // forward outputs
  float* z;
  float* kl;
  float* mu;
  float* sigma;
  uint32_t* flag;
  // backward
  const float* gz;
  const float* gkl;
  float gkl_scalar;
  float* gml;
  float* gradius;
  // tiling
  int S;  // samples per CTA tile (multiple of 32)
  int ldp_ml, ldp_eps, ldp_z, ldp_c;
  FastDiv fd_ml, fd_eps, fd_z, fd_c, fd_S;
  int vec_ok;  // all global pointers 16-byte aligned
};


int launch_pm_forward(PmParams& p, void* stream);
int launch_pm_backward(PmParams& p, void* stream);

}  // namespace mvae
