// dp_step.cu — the data-parallel optimizer step as ONE kernel over NVLink peer memory (mvae_dp_adam_step).
//
// The reference has no distributed mode; its ELBO is a SUM over the batch (stats.py:200-202), so N ranks holding B/N
// samples each reproduce the single-GPU step at batch B when the gradient / statistics bucket is SUMMED across ranks
// before the update (Trainer.build_optimizer's Adam + the curvature SGD, train.py:327-360, utils.py:148-180).
// With NCCL that is all-reduce -> Adam -> weight refresh: three dependent launches whose latency (not bandwidth — the
// bucket is 2.5 MB) is exposed at the end of every 0.3 ms step.  Here the collective and the update are one kernel:
//
//   barrier A   every rank has finished its backward pass                 (flags in peer memory, release/acquire.sys)
//   reduce-scatter + Adam
//               rank r owns the slice [r n/N, (r+1) n/N) of the parameters: it sums that slice of the gradient over
//               all ranks with 128-bit loads from the peers' buckets (NVSwitch gives every peer full bandwidth),
//               applies Adam with ITS slice of the moments (the optimizer state is sharded N ways), and
//   all-gather  stores the new parameter values straight into every peer's parameter buffer;
//               the 3+2C statistics / radius-gradient tail is summed redundantly by every rank (same order, so the
//               replicas stay bit-identical) and the radii take their SGD step locally
//   barrier B   every rank has read my bucket and written my parameters
//
// Everything the kernel needs to know about the step (Adam step count, barrier epoch) lives on the device, so the launch
// is captured once in the step's CUDA graph.  Spin loops are bounded (clock64) and raise an error word instead of
// hanging the GPU if a peer never arrives.
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include "mvae_common.cuh"

namespace mvae {

constexpr int kDpThreads = 512;
constexpr int kDpFlagStride = 32;  // uint32 per flag slot (one 128-byte line each)

struct DpPlaneTarget {
  int64_t begin, end;  // float range of the flat parameter buffer holding a [rows, cols] matrix (both % 4 == 0)
  int cols, ld, planes;
  uint16_t* base;
  int64_t plane_stride;
};

struct DpParams {
  mvae_dp_comm comm;
  int64_t n4;        // float4 elements of the network-parameter bucket
  int64_t n_net;
  int n_tail, C;
  float* m;
  float* v;
  float lr, b1, b2, eps;
  int32_t* step_dev;
  float* radius;
  float radius_lr;
  const float* radius_mask;
  float* tail_out;
  uint32_t* sync;    // local: [0] epoch, [1] arrival counter, [2] release word, [3] error
  long long spin_limit;
  int n_targets;     // weight matrices whose split-bf16 planes are refreshed after the all-gather
  DpPlaneTarget t[4];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {  // peer (or own) memory, never through a stale L1 line
  float4 r;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ float ld_peer1(const float* p) {
  float r;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
  return r;
}

// Grid-wide barrier that also spans the ranks: every CTA arrives on a local counter; CTA 0 then publishes the barrier
// index `b` in every peer's flag slot for this rank, waits until every peer has published >= b in ours, and releases
// the other CTAs.  Counters are monotonic across launches (b grows by 2 per step).
__device__ void dp_barrier(const DpParams& p, uint32_t b) {
  __syncthreads();
  uint32_t* sync = p.sync;
  const int world = p.comm.world, rank = p.comm.rank;
  if (threadIdx.x == 0) {
    __threadfence_system();
    atomicAdd(&sync[1], 1u);
  }
  if (blockIdx.x == 0) {
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
      const uint32_t target = b * gridDim.x;
      while (ld_acquire_gpu(&sync[1]) < target)
        if (clock64() - t0 > p.spin_limit) {
          sync[3] = 1u;
          break;
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < world) {
      const int r = threadIdx.x;
      st_release_sys(p.comm.flags[r] + rank * kDpFlagStride, b);
      while (ld_acquire_sys(p.comm.flags[rank] + r * kDpFlagStride) < b)
        if (clock64() - t0 > p.spin_limit) {
          sync[3] = 2u;
          break;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicExch(&sync[2], b);
    }
  } else if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (ld_acquire_gpu(&sync[2]) < b)
      if (clock64() - t0 > p.spin_limit) {
        sync[3] = 3u;
        break;
      }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kDpThreads) dp_adam_kernel(const __grid_constant__ DpParams p) {
  const int world = p.comm.world, rank = p.comm.rank;
  const uint32_t epoch = p.sync[0];  // completed steps; every CTA reads it before barrier A, CTA 0 bumps it after B
  const int32_t step = *p.step_dev + 1;
  dp_barrier(p, 2u * epoch + 1u);

  // ---- reduce-scatter + Adam + all-gather on my slice ----
  const double st = (double)step;
  const float step_size = (float)((double)p.lr / (1.0 - pow((double)p.b1, st)));
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(1.0 - pow((double)p.b2, st)));
  const float w1 = 1.f - p.b1, w2 = 1.f - p.b2;
  const int64_t per = (p.n4 + world - 1) / world;
  const int64_t lo = rank * per, hi = min(p.n4, lo + per);
  // Two float4 per thread and iteration: all 2 x world peer loads are in flight before the first is consumed (the
  // loop is latency bound: a peer load is a few microseconds over NVLink).
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += 2 * stride) {
    const int64_t idx[2] = {i0, i0 + stride};
    float4 t[2][MVAE_DP_MAX_RANKS];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int r = 0; r < MVAE_DP_MAX_RANKS; ++r)
        if (r < world && idx[u] < hi) t[u][r] = ld_peer4(p.comm.bucket[r] + 4 * idx[u]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = idx[u];
      if (i >= hi) break;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < MVAE_DP_MAX_RANKS; ++r)  // fixed rank order
        if (r < world) {
          g.x += t[u][r].x;
          g.y += t[u][r].y;
          g.z += t[u][r].z;
          g.w += t[u][r].w;
        }
      float4 mi = reinterpret_cast<float4*>(p.m)[i], vi = reinterpret_cast<float4*>(p.v)[i];
      float4 pi = reinterpret_cast<float4*>(p.comm.flat[rank])[i];
#define MVAE_ADAM1(c)                                       \
  mi.c = mi.c + (g.c - mi.c) * w1;                          \
  vi.c = vi.c * p.b2 + w2 * (g.c * g.c);                    \
  pi.c = pi.c - step_size * (mi.c / (sqrtf(vi.c) * inv_bc2_sqrt + p.eps));
      MVAE_ADAM1(x) MVAE_ADAM1(y) MVAE_ADAM1(z) MVAE_ADAM1(w)
#undef MVAE_ADAM1
      reinterpret_cast<float4*>(p.m)[i] = mi;
      reinterpret_cast<float4*>(p.v)[i] = vi;
      for (int r = 0; r < world; ++r) reinterpret_cast<float4*>(p.comm.flat[r])[i] = pi;
    }
  }
  // ---- statistics / radius-gradient tail: every rank sums it (same order), radii step locally ----
  if (blockIdx.x == gridDim.x - 1) {
    for (int t = threadIdx.x; t < p.n_tail; t += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < world; ++r) s += ld_peer1(p.comm.bucket[r] + p.n_net + t);
      p.tail_out[t] = s;
      if (t < p.C && p.radius && p.radius_lr != 0.f) {
        const float mk = p.radius_mask ? p.radius_mask[t] : 1.f;
        p.radius[t] = p.radius[t] - p.radius_lr * (s * mk);
      }
    }
  }
  __threadfence_system();
  dp_barrier(p, 2u * epoch + 2u);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.sync[0] = epoch + 1u;
    *p.step_dev = step;
  }
  // ---- every slice of the parameters has arrived: refresh the split-bf16 planes of the GEMM weights locally ----
  for (int t = 0; t < p.n_targets; ++t) {
    const DpPlaneTarget& tg = p.t[t];
    const int64_t n4 = (tg.end - tg.begin) >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      const float4 w = ld_peer4(p.comm.flat[rank] + tg.begin + 4 * i);  // peers wrote it: not through L1
      const int64_t rel = 4 * i;
      const int64_t r = rel / tg.cols;
      const int c = (int)(rel - r * tg.cols);
      float v4[4] = {w.x, w.y, w.z, w.w};
      for (int pl = 0; pl < tg.planes; ++pl) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(v4[0], v4[1]), h1 = __floats2bfloat162_rn(v4[2], v4[3]);
        const uint32_t u0 = *reinterpret_cast<const uint32_t*>(&h0), u1 = *reinterpret_cast<const uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(tg.base + pl * tg.plane_stride + r * tg.ld + c) = make_uint2(u0, u1);
        v4[0] -= __uint_as_float(u0 << 16);
        v4[1] -= __uint_as_float(u0 & 0xFFFF0000u);
        v4[2] -= __uint_as_float(u1 << 16);
        v4[3] -= __uint_as_float(u1 & 0xFFFF0000u);
      }
    }
  }
}

}  // namespace mvae

using namespace mvae;

extern "C" int mvae_dp_alloc(size_t bytes, void** dev_ptr) {
  if (!dev_ptr || bytes == 0) return MVAE_ERR_INVALID_ARGUMENT;
  void* p = nullptr;
  MVAE_CUDA_TRY(cudaMalloc(&p, bytes));
  MVAE_CUDA_TRY(cudaMemset(p, 0, bytes));
  MVAE_CUDA_TRY(cudaDeviceSynchronize());
  *dev_ptr = p;
  return MVAE_OK;
}

extern "C" int mvae_dp_free(void* dev_ptr) {
  if (dev_ptr) MVAE_CUDA_TRY(cudaFree(dev_ptr));
  return MVAE_OK;
}

extern "C" int mvae_dp_ipc_export(void* dev_ptr, uint8_t* handle_out) {
  if (!dev_ptr || !handle_out) return MVAE_ERR_INVALID_ARGUMENT;
  static_assert(sizeof(cudaIpcMemHandle_t) == MVAE_DP_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  MVAE_CUDA_TRY(cudaIpcGetMemHandle(&h, dev_ptr));
  memcpy(handle_out, &h, sizeof(h));
  return MVAE_OK;
}

extern "C" int mvae_dp_ipc_open(const uint8_t* handle, void** peer_ptr) {
  if (!handle || !peer_ptr) return MVAE_ERR_INVALID_ARGUMENT;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  MVAE_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *peer_ptr = p;
  return MVAE_OK;
}

extern "C" int mvae_dp_ipc_close(void* peer_ptr) {
  if (peer_ptr) MVAE_CUDA_TRY(cudaIpcCloseMemHandle(peer_ptr));
  return MVAE_OK;
}

extern "C" int mvae_dp_adam_step(const mvae_dp_comm* comm, int64_t n_net, int32_t n_tail, int32_t C, float* exp_avg,
                                 float* exp_avg_sq, float lr, float beta1, float beta2, float eps, int32_t* step_dev,
                                 float* radius, float radius_lr, const float* radius_mask, float* tail_out,
                                 uint32_t* sync_words, int32_t n_targets, const int64_t* target_begin,
                                 const int32_t* target_rows, const mvae_planes* targets, void* stream) {
  if (!comm || comm->world < 1 || comm->world > MVAE_DP_MAX_RANKS || comm->rank < 0 || comm->rank >= comm->world ||
      n_net < 0 || (n_net & 3) || n_tail < 0 || C < 0 || C > n_tail || !exp_avg || !exp_avg_sq || !step_dev ||
      !tail_out || !sync_words)
    return MVAE_ERR_INVALID_ARGUMENT;
  for (int r = 0; r < comm->world; ++r) {
    if (!comm->bucket[r] || !comm->flat[r] || !comm->flags[r]) return MVAE_ERR_INVALID_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(comm->bucket[r]) & 15) || (reinterpret_cast<uintptr_t>(comm->flat[r]) & 15))
      return MVAE_ERR_ALIGNMENT;
  }
  if ((reinterpret_cast<uintptr_t>(exp_avg) & 15) || (reinterpret_cast<uintptr_t>(exp_avg_sq) & 15))
    return MVAE_ERR_ALIGNMENT;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  DpParams p;
  memset(&p, 0, sizeof(p));
  p.comm = *comm;
  p.n_net = n_net;
  p.n4 = n_net / 4;
  p.n_tail = n_tail;
  p.C = C;
  p.m = exp_avg;
  p.v = exp_avg_sq;
  p.lr = lr;
  p.b1 = beta1;
  p.b2 = beta2;
  p.eps = eps;
  p.step_dev = step_dev;
  p.radius = radius;
  p.radius_lr = radius_lr;
  p.radius_mask = radius_mask;
  p.tail_out = tail_out;
  p.sync = sync_words;
  if (n_targets < 0 || n_targets > 4 || (n_targets > 0 && (!target_begin || !target_rows || !targets)))
    return MVAE_ERR_INVALID_ARGUMENT;
  p.n_targets = n_targets;
  for (int t = 0; t < n_targets; ++t) {
    const mvae_planes& pl = targets[t];
    if (!pl.base || pl.planes < 1 || pl.planes > 3 || pl.rows != target_rows[t] || pl.ld < pl.cols)
      return MVAE_ERR_INVALID_ARGUMENT;
    if ((target_begin[t] & 3) || (pl.cols & 3) || (pl.ld & 7) || (pl.plane_stride & 7) || target_begin[t] < 0 ||
        target_begin[t] + (int64_t)target_rows[t] * pl.cols > n_net || (reinterpret_cast<uintptr_t>(pl.base) & 15))
      return MVAE_ERR_ALIGNMENT;
    p.t[t].begin = target_begin[t];
    p.t[t].end = target_begin[t] + (int64_t)target_rows[t] * pl.cols;
    p.t[t].cols = pl.cols;
    p.t[t].ld = pl.ld;
    p.t[t].planes = pl.planes;
    p.t[t].base = pl.base;
    p.t[t].plane_stride = pl.planes > 1 ? pl.plane_stride : 0;
  }
  p.spin_limit = 4000000000ll;  // ~2 s of SM clocks: a peer that has not arrived by then never will
  // The CTAs spin on each other, so all of them must be resident at once: a FIXED grid of at most one CTA per SM
  // (the barrier counters assume the same grid on every launch; 64 registers x 512 threads fit one SM).
  const int grid = di.sm_count;
  dp_adam_kernel<<<grid, kDpThreads, 0, as_stream(stream)>>>(p);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}
