// pm_params.cuh — launch parameters shared by the product-manifold translation units.
#pragma once
#include "mvae_common.cuh"

namespace mvae {

constexpr int kPmMaxThreads = 512;  // register budget: 128 per thread

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
  int S;          // samples per tile (multiple of 32)
  int n_tiles;    // ceil(B / S)
  int vec_ok;     // all global pointers 16-byte aligned: tiles move with bulk async copies (TMA)
  int zero_gml;   // backward: some column of a gml row is owned by no component -> tiles are zero-filled first
};

int launch_pm_forward(PmParams& p, void* stream);
int launch_pm_backward(PmParams& p, void* stream);

}  // namespace mvae
