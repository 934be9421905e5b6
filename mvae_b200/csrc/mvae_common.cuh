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

// diagnostics: region of the timeline buffer for the next stamped launch (null when timeline mode is off)
unsigned long long* debug_timeline_region(int kind, long long ncta, int a, int b, int c);

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
// ---------------------------------------------------------------------------------------------------------
// Programmatic dependent launch.  A train step is ~14 kernels of 5-30 us each inside one CUDA graph: with plain stream
// order every kernel boundary costs launch latency plus the drain of the previous grid.  Kernels launched through
// launch_pdl() may be scheduled while their predecessor is still running; they call pdl_launch_dependents() first (so
// that THEIR successor can be scheduled early too), do whatever does not depend on the predecessor (barrier / TMEM
// set-up), and then pdl_wait(), which returns once the predecessor grid has completed and its writes are visible.
// On by default (MVAE_PDL=0 disables): inside the step's CUDA graph it is worth ~4 % of the cfg2 step.
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, FastDiv f) { return f.d <= 1 ? n : __umulhi(n, f.magic); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

}  // namespace mvae
