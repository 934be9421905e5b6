// dp_step.cu — the data-parallel optimizer step as kernels over NVLink peer memory (mvae_dp_step).
//
// The reference has no distributed mode; its ELBO is a SUM over the batch (stats.py:200-202), so N ranks holding B/N
// samples each reproduce the single-GPU step at batch B when the gradient / statistics bucket is SUMMED across ranks
// before the update (Trainer.build_optimizer's Adam + the curvature SGD, train.py:327-360, utils.py:148-180).
// With NCCL that is all-reduce -> Adam -> weight refresh: three dependent launches whose latency (not bandwidth — the
// bucket is 2.5 MB) is exposed at the end of every 0.2 ms step.  Here the collective and the update are one kernel
// per RANGE of the flat parameter buffer, so that a range whose gradient is complete early (fc_logits: half of the
// parameters, ready before the latent backward pass starts) is exchanged on a side branch UNDER the rest of the
// backward pass, and only the last range (fc_e0, heads, fc_d0 + the statistics tail) sits at the end of the step:
//
//   phase A     "my gradients of this range are complete": one flag store per peer (release.sys), issued when the
//               kernel starts (stream order already guarantees the local backward pass is done — no grid-wide arrival);
//               every CTA polls the flags the peers stored into ITS rank's memory (local loads).
//   reduce-scatter + Adam
//               rank r owns the slice [r n/N, (r+1) n/N) of the range: it sums that slice of the gradient over
//               all ranks with 128-bit loads from the peers' buckets (NVSwitch gives every peer full bandwidth),
//               applies Adam with ITS slice of the moments (the optimizer state is sharded N ways), and
//               writes the new values into ITS OWN parameter buffer;
//               the 3+2C statistics / radius-gradient tail is summed redundantly by every rank (same order, so the
//               replicas stay bit-identical), the "curvature" gradients are clipped (vae.py:161-163) and the radii take
//               their SGD step locally
//   phase B     the LAST CTA to finish (atomic ticket) tells every rank, itself included, "my slice is final and I am
//               done reading your bucket"; every CTA polls those flags, then
//   all-gather  PULLS the peers' slices with 128-bit peer loads into its own buffer (no remote stores anywhere: a load
//               completes when its data arrives, a pushed store would need a system-scope fence that waits for the
//               remote writes to drain) and refreshes the split-bf16 planes of the GEMM weights on the way.
//
// Everything the kernel needs to know about the step (Adam step count, flag epoch) lives on the device, so the launches
// are captured in the step's CUDA graph.  Waits are bounded (%globaltimer): a peer that never arrives makes the kernel
// raise a STICKY error word, write no parameter, and turn every later launch into a no-op; the host checks the word
// (it travels with the step's statistics) and raises.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mvae_common.cuh"

namespace mvae {

constexpr int kDpThreads = 512;
constexpr int kDpFlagStride = 32;  // uint32 per flag slot (one 128-byte line each)
constexpr int kDpErrWord = 8;      // index of the sticky error word in sync_words

struct DpPlaneTarget {
  int64_t begin, end;  // float range of the flat parameter buffer holding a [rows, cols] matrix (both % 4 == 0)
  int cols, ld, planes;
  uint16_t* base;
  int64_t plane_stride;
};

struct DpParams {
  mvae_dp_comm comm;
  int64_t lo4, hi4;  // float4 range of this launch
  int64_t n_net;
  int channel, do_tail;
  int n_tail, C;
  float* m;
  float* v;
  float lr, b1, b2, eps;
  int32_t* step_dev;
  float* radius;
  float radius_lr;
  const float* radius_mask;
  const float* clip_mask;
  float clip_max_norm;
  float* tail_out;
  uint32_t* sync;    // local: [4 ch + 0] epoch, [4 ch + 1] arrival ticket, [8] sticky error
  unsigned long long timeout_ns;
  int n_targets;     // weight matrices whose split-bf16 planes are refreshed after the all-gather
  DpPlaneTarget t[8];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
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
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// flag slot of (channel, phase, source rank) inside a rank's flag area
__device__ __forceinline__ int dp_slot(int channel, int phase, int src) {
  return ((channel * 2 + phase) * MVAE_DP_MAX_RANKS + src) * kDpFlagStride;
}

// Threads 0..world-1 wait until peer `tid` has stored >= e into this rank's slot; false (and the sticky error word set)
// if one of them ran out of time.  All threads of the CTA get the same answer.
__device__ bool dp_wait_peers(const DpParams& p, int phase, uint32_t e, int* s_err) {
  const int world = p.comm.world, rank = p.comm.rank;
  if ((int)threadIdx.x < world) {
    const uint32_t* f = p.comm.flags[rank] + dp_slot(p.channel, phase, threadIdx.x);
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(f) < e) {
      if (global_ns() - t0 > p.timeout_ns) {
        const uint32_t code = 1u + (uint32_t)phase;
        atomicCAS(&p.sync[kDpErrWord], 0u, code);
        p.tail_out[p.n_tail] = (float)code;  // the word the host reads with the step's statistics
        *s_err = 1;
        break;
      }
    }
  }
  __syncthreads();
  return *s_err == 0;
}

__global__ void __launch_bounds__(kDpThreads) dp_step_kernel(const __grid_constant__ DpParams p) {
  __shared__ int s_err, s_last, s_dead;
  const int world = p.comm.world, rank = p.comm.rank;
  uint32_t* sync = p.sync + 4 * p.channel;
  if (threadIdx.x == 0) s_err = 0, s_last = 0, s_dead = (p.sync[kDpErrWord] != 0u);
  __syncthreads();
  if (s_dead) return;  // a peer timed out earlier: parameters stay frozen until the host raises
  const uint32_t e = sync[0] + 1u;     // launches of this channel so far + 1; bumped by the last CTA after every CTA read it
  const int32_t step = *p.step_dev + 1;  // bumped by the launch that owns the tail, the last one of a step
  __syncthreads();

  // time stamps of the phases (low 32 bits of %globaltimer, ns) for the host to read after a run: channel 1 (the launch
  // at the end of the step) -> sync[9..13] = start, A passed, slice delivered, B passed, planes refreshed;
  // channel 0 -> sync[2], sync[3] = start, end
  const bool stamp = blockIdx.x == 0 && threadIdx.x == 0;
  if (stamp) p.sync[p.channel ? 9 : 2] = (uint32_t)global_ns();
  // ---- everything that does not depend on the peers happens BEFORE the wait: bias corrections, and the first
  //      (usually only) pass's moments and parameters, which nobody but this rank ever writes ----
  const double st = (double)step;
  const float step_size = (float)((double)p.lr / (1.0 - pow((double)p.b1, st)));
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(1.0 - pow((double)p.b2, st)));
  const float w1 = 1.f - p.b1, w2 = 1.f - p.b2;
  const int64_t per = (p.hi4 - p.lo4 + world - 1) / world;
  const int64_t lo = p.lo4 + rank * per, hi = min(p.hi4, lo + per);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // My slice is dealt out in contiguous chunks, one per CTA, so that EVERY SM issues a few of the peer loads: an SM can
  // keep only so many loads in flight, and with the slice packed into the first CTAs (512 threads x 2 x world loads
  // each) they went out in several rounds of one NVLink round trip each (measured: the slowest CTA finished 12 us
  // after phase A at N = 8, the first warp of CTA 0 after 4 us).
  const int64_t per_cta = ((hi - lo) + gridDim.x - 1) / gridDim.x;
  const int64_t c_lo = lo + blockIdx.x * per_cta, c_hi = min(hi, c_lo + per_cta);
  const int64_t first = c_lo + threadIdx.x;
  float4 m0[2], v0[2], p0[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int64_t i = first + u * blockDim.x;
    if (i < c_hi) {
      m0[u] = reinterpret_cast<const float4*>(p.m)[i];
      v0[u] = reinterpret_cast<const float4*>(p.v)[i];
      p0[u] = reinterpret_cast<const float4*>(p.comm.flat[rank])[i];
    }
  }

  // ---- phase A: announce that my gradients are complete, wait for every peer's announcement ----
  // (st.release.sys is cumulative: it covers the gradient writes of the preceding kernels, which happen before this
  // thread by stream order — no separate system-scope fence, which costs microseconds)
  if (blockIdx.x == 0 && (int)threadIdx.x < world)
    st_release_sys(p.comm.flags[threadIdx.x] + dp_slot(p.channel, 0, rank), e);
  if (!dp_wait_peers(p, 0, e, &s_err)) return;  // nothing has been written yet
  if (stamp && p.channel) p.sync[10] = (uint32_t)global_ns();
  // per-CTA stamps of the launch at the end of the step: sync[16 + 2 b] = CTA b passed phase A, [17 + 2 b] = it has
  // finished its slice (MVAE_DP_SYNC_WORDS leaves room for 256 CTAs)
  const bool cta_stamp = p.channel && threadIdx.x == 0 && blockIdx.x < 256;
  if (cta_stamp) p.sync[16 + 2 * blockIdx.x] = (uint32_t)global_ns();

  // the CTA that owns the tail puts its peer loads in flight first (they complete under the slice work below)
  const bool tail_cta = p.do_tail && blockIdx.x == gridDim.x - 1;
  float tail_v[MVAE_DP_MAX_RANKS];
  if (tail_cta && (int)threadIdx.x < p.n_tail) {
#pragma unroll
    for (int r = 0; r < MVAE_DP_MAX_RANKS; ++r)
      if (r < world) tail_v[r] = ld_peer1(p.comm.bucket[r] + p.n_net + threadIdx.x);
  }
  // ---- reduce-scatter + Adam on my slice of the range: the new values go into MY parameter buffer only ----
  // Two float4 per thread and iteration: all 2 x world peer loads are in flight before the first is consumed (the
  // loop is latency bound: a peer load is a few microseconds over NVLink).
  for (int64_t i0 = first; i0 < c_hi; i0 += 2 * blockDim.x) {
    const int64_t idx[2] = {i0, i0 + blockDim.x};
    float4 t[2][MVAE_DP_MAX_RANKS];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int r = 0; r < MVAE_DP_MAX_RANKS; ++r)
        if (r < world && idx[u] < c_hi) t[u][r] = ld_peer4(p.comm.bucket[r] + 4 * idx[u]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = idx[u];
      if (i >= c_hi) break;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < MVAE_DP_MAX_RANKS; ++r)  // fixed rank order
        if (r < world) {
          g.x += t[u][r].x;
          g.y += t[u][r].y;
          g.z += t[u][r].z;
          g.w += t[u][r].w;
        }
      float4 mi, vi, pi;
      if (i0 == first) {
        mi = m0[u], vi = v0[u], pi = p0[u];
      } else {
        mi = reinterpret_cast<const float4*>(p.m)[i];
        vi = reinterpret_cast<const float4*>(p.v)[i];
        pi = reinterpret_cast<const float4*>(p.comm.flat[rank])[i];
      }
#define MVAE_ADAM1(c)                                       \
  mi.c = mi.c + (g.c - mi.c) * w1;                          \
  vi.c = vi.c * p.b2 + w2 * (g.c * g.c);                    \
  pi.c = pi.c - step_size * (mi.c / (sqrtf(vi.c) * inv_bc2_sqrt + p.eps));
      MVAE_ADAM1(x) MVAE_ADAM1(y) MVAE_ADAM1(z) MVAE_ADAM1(w)
#undef MVAE_ADAM1
      reinterpret_cast<float4*>(p.m)[i] = mi;
      reinterpret_cast<float4*>(p.v)[i] = vi;
      reinterpret_cast<float4*>(p.comm.flat[rank])[i] = pi;
    }
  }
  if (stamp && p.channel) p.sync[14] = (uint32_t)global_ns();
  // ---- statistics / radius-gradient tail: every rank sums it (same order), clips, and steps its radii locally ----
  if (tail_cta) {
    __shared__ float s_clip;
    if ((int)threadIdx.x < p.n_tail) {  // n_tail = 2 C + 3 <= 195 < blockDim.x
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < MVAE_DP_MAX_RANKS; ++r)  // fixed rank order on every rank: replicas stay bit-identical
        if (r < world) s += tail_v[r];
      p.tail_out[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      // torch.nn.utils.clip_grad_norm_(curvature parameters, max_norm) on the rank-summed gradient (vae.py:161-163)
      float coef = 1.f;
      if (p.clip_mask) {
        float ss = 0.f;
        for (int t = 0; t < p.C; ++t) ss += p.clip_mask[t] * p.tail_out[t] * p.tail_out[t];
        coef = fminf(1.f, p.clip_max_norm / (sqrtf(ss) + 1e-6f));
      }
      s_clip = coef;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < p.C; t += blockDim.x) {
      float s = p.tail_out[t];
      if (p.clip_mask && p.clip_mask[t] != 0.f) p.tail_out[t] = s = s * s_clip;
      if (p.radius && p.radius_lr != 0.f) {
        const float mk = p.radius_mask ? p.radius_mask[t] : 1.f;
        p.radius[t] = p.radius[t] - p.radius_lr * (s * mk);
      }
    }
  }

  // ---- phase B: the last CTA to get here tells every rank (this one included) that my slice is final ----
  // Device-scope fence per thread; the system-scope fence is the publishing threads' (below): the release is cumulative
  // over everything those threads have observed through the ticket.  (A system-scope fence in every thread costs
  // 7 us here — measured — even though nothing but local memory was written.)
  __threadfence();
  if (stamp && p.channel) p.sync[15] = (uint32_t)global_ns();
  if (cta_stamp) p.sync[17 + 2 * blockIdx.x] = (uint32_t)global_ns();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t ticket = atomicAdd(&sync[1], 1u);
    __threadfence();
    s_last = (ticket + 1u == gridDim.x);
  }
  __syncthreads();
  if (s_last) {
    if ((int)threadIdx.x < world)  // cumulative release: everything observed through the ticket is covered
      st_release_sys(p.comm.flags[threadIdx.x] + dp_slot(p.channel, 1, rank), e);
    if (threadIdx.x == 0) {
      if (p.channel) p.sync[11] = (uint32_t)global_ns();
      sync[1] = 0u;  // ticket counter for the next launch of this channel
      sync[0] = e;
      if (p.do_tail) {
        *p.step_dev = step;
        p.tail_out[p.n_tail] = 0.f;
      }
    }
  }
  if (!dp_wait_peers(p, 1, e, &s_err)) return;
  if (stamp && p.channel) p.sync[12] = (uint32_t)global_ns();

  // ---- all-gather by PULLING: every slice of the range is final in its owner's buffer; read the peers' slices over
  //      NVLink (a load needs no fence — a pushed store would need a system-scope fence that waits for the remote
  //      writes to drain: measured 8 us), keep a local copy, and refresh the split-bf16 planes of the GEMM weights ----
  // (four loads in flight per thread: the loop is bound by the latency of a peer load, not by its bandwidth)
  constexpr int kGatherUnroll = 4;
  for (int64_t i0 = p.lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < p.hi4; i0 += kGatherUnroll * stride) {
    float4 wv[kGatherUnroll];
    int own[kGatherUnroll];
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < p.hi4) {
        int owner = (int)((i - p.lo4) / per);
        own[u] = owner < world ? owner : world - 1;
        wv[u] = ld_peer4(p.comm.flat[own[u]] + 4 * i);  // peers' (or my CTAs') fresh values: not through L1
      }
    }
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= p.hi4) break;
      const float4 w = wv[u];
      if (own[u] != rank) reinterpret_cast<float4*>(p.comm.flat[rank])[i] = w;
      const int64_t idx = 4 * i;
      for (int t = 0; t < p.n_targets; ++t) {
        const DpPlaneTarget& tg = p.t[t];
        if (idx < tg.begin || idx >= tg.end) continue;
        const int64_t rel = idx - tg.begin;
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
        break;
      }
    }
  }
  if (stamp) p.sync[p.channel ? 13 : 3] = (uint32_t)global_ns();
}

// Re-align the ranks: returns when every rank has launched it (flags of "channel" 2; epoch in sync[6]).
__global__ void dp_rendezvous_kernel(const mvae_dp_comm comm, uint32_t* sync, unsigned long long timeout_ns) {
  if (sync[kDpErrWord] != 0u) return;
  const uint32_t e = sync[6] + 1u;
  if ((int)threadIdx.x < comm.world) {
    st_release_sys(comm.flags[threadIdx.x] + dp_slot(2, 0, comm.rank), e);
    const uint32_t* f = comm.flags[comm.rank] + dp_slot(2, 0, threadIdx.x);
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(f) < e)
      if (global_ns() - t0 > timeout_ns) {
        atomicCAS(&sync[kDpErrWord], 0u, 3u);
        break;
      }
  }
  __syncthreads();
  if (threadIdx.x == 0) sync[6] = e;
}

}  // namespace mvae

using namespace mvae;

static unsigned long long dp_timeout_ns() {
  // a peer that has not arrived within this time never will (MVAE_DP_TIMEOUT_S, default 30 s: rank skew of an
  // evaluation pass or a checkpoint between two steps is legitimate, a dead peer is not)
  double timeout_s = 30.0;
  if (const char* env = getenv("MVAE_DP_TIMEOUT_S")) {
    const double v = atof(env);
    if (v > 0.0) timeout_s = v;
  }
  return (unsigned long long)(timeout_s * 1e9);
}

extern "C" int mvae_dp_rendezvous(const mvae_dp_comm* comm, uint32_t* sync_words, void* stream) {
  if (!comm || !sync_words || comm->world < 1 || comm->world > MVAE_DP_MAX_RANKS || comm->rank < 0 ||
      comm->rank >= comm->world)
    return MVAE_ERR_INVALID_ARGUMENT;
  for (int r = 0; r < comm->world; ++r)
    if (!comm->flags[r]) return MVAE_ERR_INVALID_ARGUMENT;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  dp_rendezvous_kernel<<<1, 32, 0, as_stream(stream)>>>(*comm, sync_words, dp_timeout_ns());
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

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

extern "C" int mvae_dp_step(const mvae_dp_comm* comm, const mvae_dp_step_args* a, void* stream) {
  if (!comm || !a || comm->world < 1 || comm->world > MVAE_DP_MAX_RANKS || comm->rank < 0 ||
      comm->rank >= comm->world || a->n_net < 0 || (a->n_net & 3) || a->begin < 0 || a->end < a->begin ||
      a->end > a->n_net || (a->begin & 3) || (a->end & 3) || a->channel < 0 || a->channel >= MVAE_DP_CHANNELS ||
      a->n_tail < 0 || a->n_tail > 512 || a->C < 0 || a->C > a->n_tail || !a->exp_avg || !a->exp_avg_sq || !a->step_dev || !a->tail_out ||
      !a->sync_words)
    return MVAE_ERR_INVALID_ARGUMENT;
  for (int r = 0; r < comm->world; ++r) {
    if (!comm->bucket[r] || !comm->flat[r] || !comm->flags[r]) return MVAE_ERR_INVALID_ARGUMENT;
    if ((reinterpret_cast<uintptr_t>(comm->bucket[r]) & 15) || (reinterpret_cast<uintptr_t>(comm->flat[r]) & 15))
      return MVAE_ERR_ALIGNMENT;
  }
  if ((reinterpret_cast<uintptr_t>(a->exp_avg) & 15) || (reinterpret_cast<uintptr_t>(a->exp_avg_sq) & 15))
    return MVAE_ERR_ALIGNMENT;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  DpParams p;
  memset(&p, 0, sizeof(p));
  p.comm = *comm;
  p.n_net = a->n_net;
  p.lo4 = a->begin / 4;
  p.hi4 = a->end / 4;
  p.channel = a->channel;
  p.do_tail = a->do_tail ? 1 : 0;
  p.n_tail = a->n_tail;
  p.C = a->C;
  p.m = a->exp_avg;
  p.v = a->exp_avg_sq;
  p.lr = a->lr;
  p.b1 = a->beta1;
  p.b2 = a->beta2;
  p.eps = a->eps;
  p.step_dev = a->step_dev;
  p.radius = a->radius;
  p.radius_lr = a->radius_lr;
  p.radius_mask = a->radius_mask;
  p.clip_mask = a->clip_mask;
  p.clip_max_norm = a->clip_max_norm;
  p.tail_out = a->tail_out;
  p.sync = a->sync_words;
  if (a->n_targets < 0 || a->n_targets > 8 ||
      (a->n_targets > 0 && (!a->target_begin || !a->target_rows || !a->targets)))
    return MVAE_ERR_INVALID_ARGUMENT;
  p.n_targets = a->n_targets;
  int64_t refresh4 = 0;
  for (int t = 0; t < a->n_targets; ++t) {
    const mvae_planes& pl = a->targets[t];
    if (!pl.base || pl.planes < 1 || pl.planes > 3 || pl.rows != a->target_rows[t] || pl.ld < pl.cols)
      return MVAE_ERR_INVALID_ARGUMENT;
    const int64_t tb = a->target_begin[t], te = tb + (int64_t)a->target_rows[t] * pl.cols;
    if ((tb & 3) || (pl.cols & 3) || (pl.ld & 7) || (pl.plane_stride & 7) || (reinterpret_cast<uintptr_t>(pl.base) & 15))
      return MVAE_ERR_ALIGNMENT;
    if (tb < a->begin || te > a->end) return MVAE_ERR_INVALID_ARGUMENT;  // refreshed from THIS launch's range only
    p.t[t].begin = tb;
    p.t[t].end = te;
    p.t[t].cols = pl.cols;
    p.t[t].ld = pl.ld;
    p.t[t].planes = pl.planes;
    p.t[t].base = pl.base;
    p.t[t].plane_stride = pl.planes > 1 ? pl.plane_stride : 0;
    refresh4 = refresh4 > (te - tb) / 4 ? refresh4 : (te - tb) / 4;
  }
  p.timeout_ns = dp_timeout_ns();
  // The CTAs wait for each other (ticket in phase B), so all of them must be resident at once: at most one CTA per SM.
  // Enough CTAs for one pass over my slice (two float4 per thread) and two passes of the gather over the range.
  const int64_t per4 = (p.hi4 - p.lo4 + comm->world - 1) / comm->world;
  int64_t want = (per4 + 2 * kDpThreads - 1) / (2 * kDpThreads);
  const int64_t want_gather = (p.hi4 - p.lo4 + 2 * kDpThreads - 1) / (2 * kDpThreads);
  if (want_gather > want) want = want_gather;
  (void)refresh4;
  int grid = (int)(want < 1 ? 1 : (want > di.sm_count ? di.sm_count : want));
  if (a->max_ctas > 0) {
    if (grid > a->max_ctas) grid = a->max_ctas;
  } else {
    grid = di.sm_count;  // the launch at the end of the step has the machine to itself: spread the peer loads over all SMs
  }
  dp_step_kernel<<<grid, kDpThreads, 0, as_stream(stream)>>>(p);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}
