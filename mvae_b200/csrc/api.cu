// api.cu — status strings, ABI version, cached device attributes of libmvae_b200.so.
#include <stdlib.h>

#include <mutex>

#include "mvae_common.cuh"

namespace mvae {

static thread_local int g_last_cuda_error = 0;
void note_cuda_error(cudaError_t e) { g_last_cuda_error = (int)e; }

int get_device_info(DeviceInfo* out) {
  constexpr int kMaxDev = 64;
  static DeviceInfo cache[kMaxDev];
  static bool have[kMaxDev] = {};
  static std::mutex mu;
  int dev = 0;
  MVAE_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) return MVAE_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lock(mu);
  if (!have[dev]) {
    DeviceInfo d;
    MVAE_CUDA_TRY(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
    MVAE_CUDA_TRY(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    MVAE_CUDA_TRY(cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    MVAE_CUDA_TRY(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cache[dev] = d;
    have[dev] = true;
  }
  *out = cache[dev];
  if (out->cc_major != 10) return MVAE_ERR_NOT_SM100;
  return MVAE_OK;
}

// diagnostics of the fused latent block (latent_impl.cuh): read by launch_latent on every launch
unsigned long long* g_lat_stamps = nullptr;
int g_lat_debug_flags = 0;
// diagnostics of the tcgen05 GEMM (gemm_sm100.cu)
unsigned long long* g_gemm_stamps = nullptr;
int g_gemm_debug_flags = 0;
// timeline mode (mvae_debug_timeline): every stamped launch gets its OWN 16-word-per-CTA region of one buffer, in launch
// (= capture) order, so that the launches of a captured step can be laid side by side after a replay
unsigned long long* g_dbg_cursor = nullptr;
static int g_dbg_log[256][5];
static int g_dbg_log_n = 0;
unsigned long long* debug_timeline_region(int kind, long long ncta, int a, int b, int c) {
  if (!g_dbg_cursor || g_dbg_log_n >= 256) return nullptr;
  unsigned long long* r = g_dbg_cursor;
  g_dbg_cursor += ncta * 16;
  int* e = g_dbg_log[g_dbg_log_n++];
  e[0] = kind; e[1] = (int)ncta; e[2] = a; e[3] = b; e[4] = c;
  return r;
}

bool pdl_enabled() {
  static const bool on = [] {
    // on by default (round 2: 0.181 -> 0.174 ms on the captured cfg2 step: a kernel's barrier / TMEM set-up runs under
    // its predecessor's tail); MVAE_PDL=0 turns it off
    const char* e = getenv("MVAE_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

}  // namespace mvae

// ---- thin CUDA-runtime helpers for the host-side pipeline of train_epoch (no kernels) ----
extern "C" int mvae_rt_memcpy_async(void* dst, const void* src, size_t bytes, void* stream) {
  if (!dst || !src) return MVAE_ERR_INVALID_ARGUMENT;
  if (bytes == 0) return MVAE_OK;
  MVAE_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, mvae::as_stream(stream)));
  return MVAE_OK;
}

extern "C" int mvae_rt_event_record(void* event, void* stream) {
  if (!event) return MVAE_ERR_INVALID_ARGUMENT;
  MVAE_CUDA_TRY(cudaEventRecord(reinterpret_cast<cudaEvent_t>(event), mvae::as_stream(stream)));
  return MVAE_OK;
}

extern "C" int mvae_rt_stream_wait_event(void* stream, void* event) {
  if (!event) return MVAE_ERR_INVALID_ARGUMENT;
  MVAE_CUDA_TRY(cudaStreamWaitEvent(mvae::as_stream(stream), reinterpret_cast<cudaEvent_t>(event), 0));
  return MVAE_OK;
}

extern "C" int mvae_debug_latent(unsigned long long* stamps, int32_t flags) {
  mvae::g_lat_stamps = stamps;
  mvae::g_lat_debug_flags = flags;
  return MVAE_OK;
}

extern "C" int mvae_debug_gemm(unsigned long long* stamps, int32_t flags) {
  mvae::g_gemm_stamps = stamps;
  mvae::g_gemm_debug_flags = flags;
  return MVAE_OK;
}

extern "C" int mvae_debug_timeline(unsigned long long* buffer) {
  mvae::g_dbg_cursor = buffer;
  if (buffer) mvae::g_dbg_log_n = 0;
  return MVAE_OK;
}

extern "C" int mvae_debug_timeline_log(int32_t* out, int32_t max_entries) {
  int n = mvae::g_dbg_log_n < max_entries ? mvae::g_dbg_log_n : max_entries;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < 5; ++j) out[i * 5 + j] = mvae::g_dbg_log[i][j];
  return n;
}

extern "C" const char* mvae_strerror(int status) {
  switch (status) {
    case MVAE_OK: return "ok";
    case MVAE_ERR_INVALID_ARGUMENT: return "invalid argument";
    case MVAE_ERR_UNSUPPORTED: return "unsupported request";
    case MVAE_ERR_CUDA: return "CUDA runtime error (see mvae_last_cuda_error)";
    case MVAE_ERR_ALIGNMENT: return "pointer or leading dimension violates the documented alignment";
    case MVAE_ERR_NOT_SM100: return "device is not compute capability 10.x (this library is sm_100a only)";
    default: return "unknown mvae status";
  }
}

extern "C" int mvae_abi_version(void) { return MVAE_ABI_VERSION; }

extern "C" int mvae_last_cuda_error(void) { return mvae::g_last_cuda_error; }

extern "C" int mvae_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  mvae::DeviceInfo d;
  int rc = mvae::get_device_info(&d);
  if (rc != MVAE_OK && rc != MVAE_ERR_NOT_SM100) return rc;
  if (sm_count) *sm_count = d.sm_count;
  if (cc_major) *cc_major = d.cc_major;
  if (cc_minor) *cc_minor = d.cc_minor;
  return rc;
}
