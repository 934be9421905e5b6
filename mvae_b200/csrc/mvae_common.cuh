// mvae_common.cuh — shared host/device helpers of libmvae_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mvae_b200.h"

namespace mvae {

// Thread-local record of the last CUDA error seen by an entry point (mvae_last_cuda_error()).
void note_cuda_error(cudaError_t e);

#define MVAE_CUDA_TRY(expr)                    \
  do {                                         \
    cudaError_t _e = (expr);                   \
    if (_e != cudaSuccess) {                   \
      ::mvae::note_cuda_error(_e);             \
      return MVAE_ERR_CUDA;                    \
    }                                          \
  } while (0)

// Kernel launches are checked with cudaPeekAtLastError (does not clear a sticky error, does not sync).
#define MVAE_LAUNCH_CHECK() MVAE_CUDA_TRY(cudaPeekAtLastError())

// Per-device cached attributes (SM count, compute capability).  Filled lazily, never synchronises.
struct DeviceInfo {
  int sm_count = 0;
  int cc_major = 0;
  int cc_minor = 0;
  int max_smem_optin = 0;
};
int get_device_info(DeviceInfo* out);  // returns mvae_status

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------------------------------------
// Exact division of small unsigned numbers by a runtime constant: q = n / d for n, d < 2^16
// (magic = ceil(2^32 / d); the error term n*e/(d*2^32) < 1/d because n*e < 2^32).
struct FastDiv {
  uint32_t d;
  uint32_t magic;
};
static inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  f.magic = d <= 1 ? 0u : (uint32_t)(((1ull << 32) + d - 1) / d);
  return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, FastDiv f) { return f.d <= 1 ? n : __umulhi(n, f.magic); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

}  // namespace mvae
