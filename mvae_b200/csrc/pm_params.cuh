// pm_params.cuh — launch parameters shared by the product-manifold translation units.
#pragma once
#include "mvae_common.cuh"

namespace mvae {

constexpr int kPmMaxWarps = 8;
constexpr int kPmMaxThreads = 32 * kPmMaxWarps;  // with min-blocks launch bounds: 64 / 80 / 128 registers per thread

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
  int vec_ok;     // all global pointers 16-byte aligned: tiles move with bulk async copies (TMA)
  int zero_gml;   // backward: some column of a gml row is owned by no component -> tiles are zero-filled first
  // Tiling, computed on the host so that the kernel reads every offset as an immediate constant-bank operand
  // (offsets in floats from the start of dynamic shared memory unless named bytes).
  int S;            // samples per tile = 32 * nb
  int nb;           // 32-sample blocks per tile
  int n_warps;      // warps per CTA (one-component-per-warp variant: sum of the per-component warp counts)
  // one-component-per-warp variant: warp w works on component w_ci[w]: blocks w_blk0[w], + w_bstride[w], ...
  // (w_nblk[w] of them) of every tile.  Components get warps in proportion to their arithmetic cost.
  uint8_t w_ci[kPmMaxWarps], w_blk0[kPmMaxWarps], w_nblk[kPmMaxWarps], w_bstride[kPmMaxWarps];
  int nst;          // input stages of the shared-memory ring (2..8)
  int n_tiles;      // ceil(B / S)
  int n_bulk_tiles; // tiles [0, n_bulk_tiles) are full and aligned: moved by bulk copies
  int in_base, in_stage, in_ml, in_eps, in_gz, in_gkl;
  int out_base, out_stage, out_a, out_b, out_c, out_d;  // fwd: z, kl, mu, sigma   bwd: gml
  int bar_base;
  uint32_t bytes_ml, bytes_eps, bytes_z, bytes_c, bytes_in;  // bulk copy sizes of one tile
};

int launch_pm_forward(PmParams& p, void* stream);
int launch_pm_backward(PmParams& p, void* stream);

}  // namespace mvae
