// pm_kernels_impl.cuh — the fused per-sample product-manifold kernels (K3 of SURVEY.md §2.2) behind
// mvae_pm_forward / mvae_pm_backward (include/mvae_b200.h).  Arithmetic: pm_math.cuh.
//
// Design.  The kernels are PERSISTENT: gridDim = SMs x resident CTAs, each CTA walks tiles of S
// consecutive samples (S = 32 or 64).  The rows of a tile are contiguous in HBM, so every tensor of a tile moves with
// ONE bulk asynchronous copy (TMA, cp.async.bulk) issued by one thread: inputs (ml, eps[, gz, gkl]) land in a
// two-stage shared-memory ring guarded by mbarriers — the copy of tile k+1 is in flight while tile k is computed —
// and outputs (z, kl[, mu, sigma] / gml) leave from a double-buffered staging tile with bulk stores.  No thread spends
// issue slots on address arithmetic for global memory, which matters because the kernels are instruction-issue bound.
// Work items are (component, 32 samples): a warp keeps ONE component for the whole launch, so its descriptor and the
// curvature constants (R, 1/R, R^2, ...) sit in registers, there is no divergence, and the C components of a sample
// run in parallel on C warps.  The backward kernel recomputes the forward from its inputs and accumulates dR in a
// register across all tiles of the CTA: one warp reduction and one global atomic per warp per launch.
// Ragged last tiles and unaligned pointers take a cooperative load / store path through the same shared-memory tiles.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "pm_math.cuh"
#include "pm_params.cuh"

#ifndef MVAE_PM_BWD
#error "define MVAE_PM_BWD to 0 (forward kernels) or 1 (backward kernels) before including pm_kernels_impl.cuh"
#endif

namespace mvae {

using namespace pm;

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t pm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pm_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pm_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pm_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// global -> shared bulk copy, completion counted in bytes on an mbarrier
__device__ __forceinline__ void pm_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void pm_bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void pm_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void pm_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void pm_bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (bulk stores read them)
__device__ __forceinline__ void pm_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// cooperative copies of the ragged / unaligned path
__device__ __forceinline__ void coop_load(float* __restrict__ s, const float* __restrict__ g, int total) {
  for (int i = threadIdx.x; i < total; i += blockDim.x) s[i] = __ldg(g + i);
}
__device__ __forceinline__ void coop_store(float* __restrict__ g, const float* __restrict__ s, int total) {
  for (int i = threadIdx.x; i < total; i += blockDim.x) g[i] = s[i];
}

// ------------------------------------------------------------------------------------------------ one item
// What a warp needs to know about its component, staged once in shared memory per CTA (3 x 128-bit loads).
struct __align__(16) ItemInfo {
  int type, n, l_n, d;
  int m_off, l_off, eps_off, z_off;
  CompConst K;  // R, 1/R
  float rp;
  int pad;
};

template <int N, bool BWD, bool WANT_MS>
__device__ __forceinline__ bool run_item(const ItemInfo& c, const float* ml_row, const float* eps_row, float* z_row,
                                         float* kl_slot, float* mu_row, float* sigma_row, const float* gz_row,
                                         float gkl, float* gml_row, float* gR_acc) {
  CompOut<N> o;
  const int n = N > 0 ? N : c.n;
  constexpr int CN = Cap<N>::n;
  // operands are read from the shared-memory tile where they are needed (keeps the register footprint small)
  const float* m = ml_row + c.m_off;
  const float* l = ml_row + c.l_off;
  const float* e = eps_row + c.eps_off;
  const float* gz = BWD ? gz_row + c.z_off : nullptr;
  float gm[BWD ? CN : 1], gl[BWD ? CN : 1];
  float gR = 0.f;
  switch (c.type) {
    case MVAE_EUCLIDEAN: comp_e<N, BWD>(n, c.l_n, m, l, e, o, gz, gkl, gm, gl); break;
    case MVAE_HYPERBOLOID: comp_hsp<N, BWD, kHyp, WANT_MS>(n, c.l_n, m, l, e, c.K, o, gz, gkl, gm, gl, &gR); break;
    case MVAE_SPHERE: comp_hsp<N, BWD, kSph, WANT_MS>(n, c.l_n, m, l, e, c.K, o, gz, gkl, gm, gl, &gR); break;
    default: comp_hsp<N, BWD, kPoi, WANT_MS>(n, c.l_n, m, l, e, c.K, o, gz, gkl, gm, gl, &gR); break;
  }
  if (BWD) {
    *gR_acc += gR;
    if (N > 0) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        gml_row[c.m_off + j] = gm[j];
        if (j < c.l_n) gml_row[c.l_off + j] = gl[j];
      }
    } else {
      for (int j = 0; j < n; ++j) {
        gml_row[c.m_off + j] = gm[j];
        if (j < c.l_n) gml_row[c.l_off + j] = gl[j];
      }
    }
    return true;
  }
  // a sum is non-finite iff a term is (up to overflow of finite terms near FLT_MAX, which the flag may also report)
  float chk = o.kl;
  const int d = c.d;
  if (N > 0) {
#pragma unroll
    for (int k = 0; k < N + 1; ++k)
      if (k < d) {
        z_row[c.z_off + k] = o.z[k];
        chk += o.z[k];
        if (WANT_MS) mu_row[c.z_off + k] = o.mu[k];
      }
    if (WANT_MS) {
#pragma unroll
      for (int j = 0; j < N; ++j) sigma_row[c.eps_off + j] = o.sigma[j];
    }
  } else {
    for (int k = 0; k < d; ++k) {
      z_row[c.z_off + k] = o.z[k];
      chk += o.z[k];
      if (WANT_MS) mu_row[c.z_off + k] = o.mu[k];
    }
    if (WANT_MS)
      for (int j = 0; j < n; ++j) sigma_row[c.eps_off + j] = o.sigma[j];
  }
  *kl_slot = o.kl;
  return chk - chk == 0.f;
}

template <bool BWD, int MAXN, bool WANT_MS>
__device__ __forceinline__ bool dispatch_item(const ItemInfo& c, const float* ml_row, const float* eps_row,
                                              float* z_row, float* kl_slot, float* mu_row, float* sigma_row,
                                              const float* gz_row, float gkl, float* gml_row, float* gR_acc) {
  // MAXN > 0: only dimensions <= MAXN are compiled in (register budget = that of the widest one); MAXN == 0: all
  // static dimensions plus the runtime-dimension path.
#define MVAE_CASE(NN)                                                                                               \
  case NN:                                                                                                          \
    if constexpr (MAXN == 0 || NN <= MAXN)                                                                          \
      return run_item<NN, BWD, WANT_MS>(c, ml_row, eps_row, z_row, kl_slot, mu_row, sigma_row, gz_row, gkl, gml_row, \
                                        gR_acc);                                                                    \
    else                                                                                                            \
      return true;
  switch (c.n) {
    MVAE_CASE(1)
    MVAE_CASE(2)
    MVAE_CASE(3)
    MVAE_CASE(4)
    MVAE_CASE(5)
    MVAE_CASE(6)
    MVAE_CASE(8)
    default:
      if constexpr (MAXN == 0)
        return run_item<0, BWD, WANT_MS>(c, ml_row, eps_row, z_row, kl_slot, mu_row, sigma_row, gz_row, gkl, gml_row,
                                         gR_acc);
      else
        return true;  // unreachable: the host picks the kernel whose MAXN covers every component
  }
#undef MVAE_CASE
}

// ------------------------------------------------------------------------------------------------ tile layout
// Shared memory (floats): [ItemInfo x C] | 2 x input stage | 2 x output stage | 2 mbarriers.  Every tile is a
// multiple of 128 bytes (S is a multiple of 32) so all bulk copies are 16-byte aligned.
struct PmLayout {
  int info_floats;
  int in_ml, in_eps, in_gz, in_gkl, in_stage;      // offsets inside an input stage, stage size
  int out_a, out_b, out_c, out_d, out_stage;       // fwd: z, kl, mu, sigma    bwd: gml
  int total_floats;
};
__host__ __device__ inline PmLayout pm_layout(int C, int ld_ml, int ld_eps, int ld_z, int S, bool bwd, bool want_ms,
                                              bool has_gkl) {
  PmLayout L;
  L.info_floats = (C * (int)(sizeof(ItemInfo) / 4) + 31) & ~31;
  int o = 0;
  L.in_ml = o;
  o += S * ld_ml;
  L.in_eps = o;
  o += S * ld_eps;
  L.in_gz = o;
  if (bwd) o += S * ld_z;
  L.in_gkl = o;
  if (bwd && has_gkl) o += S * C;
  L.in_stage = o;
  o = 0;
  L.out_a = o;
  o += bwd ? S * ld_ml : S * ld_z;
  L.out_b = o;
  if (!bwd) o += S * C;
  L.out_c = o;
  if (!bwd && want_ms) o += S * ld_z;
  L.out_d = o;
  if (!bwd && want_ms) o += S * ld_eps;
  L.out_stage = o;
  L.total_floats = L.info_floats + 2 * L.in_stage + 2 * L.out_stage + 8;
  return L;
}

// SINGLE: blockDim = 32 x (warp-items per tile), every warp keeps ONE (component, 32-sample block) for the whole launch.
// !SINGLE: more warp-items than warps (many components): warps loop over the items of a tile.
template <bool BWD, int MAXN, bool WANT_MS, bool SINGLE>
__device__ __forceinline__ void pm_kernel_body(const PmParams& p) {
  extern __shared__ __align__(128) float smem[];
  const int C = p.desc.C, S = p.S;
  const int ld_ml = p.desc.ld_ml, ld_eps = p.desc.ld_eps, ld_z = p.desc.ld_z;
  const bool has_gkl = BWD && p.gkl != nullptr;
  const PmLayout L = pm_layout(C, ld_ml, ld_eps, ld_z, S, BWD, WANT_MS, has_gkl);
  ItemInfo* info = reinterpret_cast<ItemInfo*>(smem);
  float* in_base = smem + L.info_floats;
  float* out_base = in_base + 2 * L.in_stage;
  const uint32_t bar0 = pm_smem_u32(out_base + 2 * L.out_stage);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int i = tid; i < C; i += blockDim.x) {
    const mvae_component c = p.desc.comp[i];
    ItemInfo ii;
    ii.type = c.type;
    ii.n = c.n;
    ii.l_n = c.l_n;
    ii.d = c.d;
    ii.m_off = c.m_off;
    ii.l_off = c.l_off;
    ii.eps_off = c.eps_off;
    ii.z_off = c.z_off;
    ii.rp = (p.radius && c.type != MVAE_EUCLIDEAN) ? __ldg(p.radius + i) : 1.f;
    ii.K = make_const(ii.rp);
    ii.pad = 0;
    info[i] = ii;
  }
  if (tid == 0) {
    pm_mbar_init(bar0, 1);
    pm_mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int wpc = S >> 5;            // warps per component inside a tile
  const int n_items = C * wpc;       // warp-items per tile
  const uint32_t in_bytes = (uint32_t)L.in_stage * 4u;

  auto use_bulk = [&](int tile) { return p.vec_ok && (int64_t)(tile + 1) * S <= p.B; };
  auto issue_loads = [&](int tile, int st) {  // one thread
    const int64_t row0 = (int64_t)tile * S;
    const uint32_t bar = bar0 + 8u * st;
    const uint32_t dst = pm_smem_u32(in_base + st * L.in_stage);
    pm_mbar_expect_tx(bar, in_bytes);
    pm_bulk_g2s(dst + 4u * L.in_ml, p.ml + row0 * ld_ml, (uint32_t)(S * ld_ml) * 4u, bar);
    pm_bulk_g2s(dst + 4u * L.in_eps, p.eps + row0 * ld_eps, (uint32_t)(S * ld_eps) * 4u, bar);
    if (BWD) {
      pm_bulk_g2s(dst + 4u * L.in_gz, p.gz + row0 * ld_z, (uint32_t)(S * ld_z) * 4u, bar);
      if (has_gkl) pm_bulk_g2s(dst + 4u * L.in_gkl, p.gkl + row0 * C, (uint32_t)(S * C) * 4u, bar);
    }
  };

  // the component this warp keeps (single-item mode)
  ItemInfo mine;
  int my_ci = 0, my_sub = 0;
  if (SINGLE) {
    my_ci = warp / wpc;
    mine = info[my_ci];
    my_sub = (warp % wpc) << 5;
  }
  float gR_acc = 0.f;
  bool finite = true;

  int tile = blockIdx.x;
  if (tid == 0 && tile < p.n_tiles && use_bulk(tile)) issue_loads(tile, 0);
  for (int k = 0; tile < p.n_tiles; ++k, tile += gridDim.x) {
    const int st = k & 1;
    const int64_t row0 = (int64_t)tile * S;
    const int rows = (int)min((int64_t)S, p.B - row0);
    const bool bulk = use_bulk(tile);
    float* sin_ = in_base + st * L.in_stage;
    float* sout = out_base + st * L.out_stage;
    const int next = tile + gridDim.x;
    if (tid == 0 && next < p.n_tiles && use_bulk(next)) issue_loads(next, st ^ 1);
    if (bulk) {
      pm_mbar_wait(bar0 + 8u * st, (uint32_t)(k >> 1) & 1u);
    } else {
      coop_load(sin_ + L.in_ml, p.ml + row0 * ld_ml, rows * ld_ml);
      coop_load(sin_ + L.in_eps, p.eps + row0 * ld_eps, rows * ld_eps);
      if (BWD) {
        coop_load(sin_ + L.in_gz, p.gz + row0 * ld_z, rows * ld_z);
        if (has_gkl) coop_load(sin_ + L.in_gkl, p.gkl + row0 * C, rows * C);
      }
      __syncthreads();
    }
    if (BWD && p.zero_gml) {
      for (int i = tid; i < S * ld_ml; i += blockDim.x) sout[L.out_a + i] = 0.f;
      __syncthreads();
    }
    // ---- compute ----
    if (SINGLE) {
      const int sidx = my_sub + lane;
      if (sidx < rows) {
        const float gkl = BWD ? (has_gkl ? sin_[L.in_gkl + sidx * C + my_ci] : p.gkl_scalar) : 0.f;
        finite &= dispatch_item<BWD, MAXN, WANT_MS>(
            mine, sin_ + L.in_ml + sidx * ld_ml, sin_ + L.in_eps + sidx * ld_eps, sout + L.out_a + sidx * ld_z,
            sout + L.out_b + sidx * C + my_ci, sout + L.out_c + sidx * ld_z, sout + L.out_d + sidx * ld_eps,
            sin_ + L.in_gz + sidx * ld_z, gkl, sout + L.out_a + sidx * ld_ml, &gR_acc);
      }
    } else {
      for (int w = warp; w < n_items; w += nwarps) {
        const int ci = w / wpc;
        const int sidx = ((w % wpc) << 5) + lane;
        float gR = 0.f;
        if (sidx < rows) {
          const ItemInfo c = info[ci];
          const float gkl = BWD ? (has_gkl ? sin_[L.in_gkl + sidx * C + ci] : p.gkl_scalar) : 0.f;
          finite &= dispatch_item<BWD, MAXN, WANT_MS>(
              c, sin_ + L.in_ml + sidx * ld_ml, sin_ + L.in_eps + sidx * ld_eps, sout + L.out_a + sidx * ld_z,
              sout + L.out_b + sidx * C + ci, sout + L.out_c + sidx * ld_z, sout + L.out_d + sidx * ld_eps,
              sin_ + L.in_gz + sidx * ld_z, gkl, sout + L.out_a + sidx * ld_ml, &gR);
          gR *= radius_d(c.rp);
        }
        if (BWD && p.gradius) {
          gR = warp_sum(gR);
          if (lane == 0 && gR != 0.f) atomicAdd(p.gradius + ci, gR);
        }
      }
    }
    // ---- store ----
    if (bulk) {
      pm_fence_async();
      if (tid == 0) pm_bulk_wait_read0();  // the store of tile k-1 (other staging buffer) has left shared memory
      __syncthreads();
      if (tid == 0) {
        const uint32_t src = pm_smem_u32(sout);
        if (BWD) {
          pm_bulk_s2g(p.gml + row0 * ld_ml, src + 4u * L.out_a, (uint32_t)(S * ld_ml) * 4u);
        } else {
          pm_bulk_s2g(p.z + row0 * ld_z, src + 4u * L.out_a, (uint32_t)(S * ld_z) * 4u);
          pm_bulk_s2g(p.kl + row0 * C, src + 4u * L.out_b, (uint32_t)(S * C) * 4u);
          if (WANT_MS) {
            pm_bulk_s2g(p.mu + row0 * ld_z, src + 4u * L.out_c, (uint32_t)(S * ld_z) * 4u);
            pm_bulk_s2g(p.sigma + row0 * ld_eps, src + 4u * L.out_d, (uint32_t)(S * ld_eps) * 4u);
          }
        }
        pm_bulk_commit();
      }
    } else {
      __syncthreads();
      if (BWD) {
        coop_store(p.gml + row0 * ld_ml, sout + L.out_a, rows * ld_ml);
      } else {
        coop_store(p.z + row0 * ld_z, sout + L.out_a, rows * ld_z);
        coop_store(p.kl + row0 * C, sout + L.out_b, rows * C);
        if (WANT_MS) {
          coop_store(p.mu + row0 * ld_z, sout + L.out_c, rows * ld_z);
          coop_store(p.sigma + row0 * ld_eps, sout + L.out_d, rows * ld_eps);
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0) pm_bulk_wait_all0();  // shared memory must outlive the last bulk store
  if (BWD) {
    if (SINGLE && p.gradius) {
      const float g = warp_sum(gR_acc * radius_d(mine.rp));
      if (lane == 0 && g != 0.f) atomicAdd(p.gradius + my_ci, g);
    }
  } else if (p.flag) {
    const unsigned bad = __ballot_sync(0xffffffffu, !finite);
    if (bad && lane == 0) atomicOr(p.flag, 1u);
  }
}

#if !MVAE_PM_BWD
template <int MAXN, bool WANT_MS, bool SINGLE>
__global__ void __launch_bounds__(kPmMaxThreads, 1) pm_forward_kernel(const __grid_constant__ PmParams p) {
  pm_kernel_body<false, MAXN, WANT_MS, SINGLE>(p);
}
#else
template <int MAXN, bool SINGLE>
__global__ void __launch_bounds__(kPmMaxThreads, 1) pm_backward_kernel(const __grid_constant__ PmParams p) {
  pm_kernel_body<true, MAXN, false, SINGLE>(p);
}
#endif

// ------------------------------------------------------------------------------------------------ host side
static int launch_pm(PmParams& p, void* stream) {
  constexpr bool bwd = MVAE_PM_BWD != 0;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  const mvae_pm_desc& D = p.desc;
  const bool want_ms = !bwd && (p.mu != nullptr || p.sigma != nullptr);
  if (want_ms && (!p.mu || !p.sigma)) return MVAE_ERR_INVALID_ARGUMENT;
  const bool has_gkl = bwd && p.gkl != nullptr;
  int maxn = 0;
  bool dyn = false;
  for (int i = 0; i < D.C; ++i) {
    const int n = D.comp[i].n;
    dyn = dyn || !(n >= 1 && n <= 8 && n != 7);
    maxn = n > maxn ? n : maxn;
  }
  if (bwd) {
    // does every column of a gml row belong to exactly one component?  (packed descriptors: yes)
    int owned = 0;
    for (int i = 0; i < D.C; ++i) owned += D.comp[i].n + D.comp[i].l_n;
    p.zero_gml = owned != D.ld_ml;
  }
  // one warp per (component, 32 samples) when that fits a CTA, else the looping variant with all dimensions compiled in
  int S = 64;
  if (p.B < (int64_t)64 * 8 * di.sm_count || D.C * 2 * 32 > kPmMaxThreads) S = 32;
  {
    static const char* s_env = getenv("MVAE_PM_TILE");  // tuning aid: force the tile height (32 | 64)
    if (s_env && (atoi(s_env) == 32 || (atoi(s_env) == 64 && D.C * 2 * 32 <= kPmMaxThreads))) S = atoi(s_env);
  }
  const bool single = !dyn && D.C * 32 <= kPmMaxThreads;
  void (*kern)(PmParams);
#if MVAE_PM_BWD
  kern = !single ? pm_backward_kernel<0, false> : maxn <= 2 ? pm_backward_kernel<2, true>
                                               : maxn <= 4   ? pm_backward_kernel<4, true>
                                                             : pm_backward_kernel<8, true>;
#else
  if (want_ms)
    kern = !single ? pm_forward_kernel<0, true, false> : maxn <= 2 ? pm_forward_kernel<2, true, true>
                                                     : maxn <= 4   ? pm_forward_kernel<4, true, true>
                                                                   : pm_forward_kernel<8, true, true>;
  else
    kern = !single ? pm_forward_kernel<0, false, false> : maxn <= 2 ? pm_forward_kernel<2, false, true>
                                                      : maxn <= 4   ? pm_forward_kernel<4, false, true>
                                                                    : pm_forward_kernel<8, false, true>;
#endif
  // Tile height: 64 samples when the batch still yields several tiles per resident CTA, else 32 (small batches want
  // as many CTAs as possible).
  cudaFuncAttributes fa;
  MVAE_CUDA_TRY(cudaFuncGetAttributes(&fa, kern));
  int threads = 0, blocks_per_sm = 0;
  size_t smem = 0;
  for (;; S = 32) {
    const PmLayout L = pm_layout(D.C, D.ld_ml, D.ld_eps, D.ld_z, S, bwd, want_ms, has_gkl);
    smem = (size_t)L.total_floats * 4;
    int warps = D.C * (S / 32);
    if (!single) {
      if (warps > 8) warps = 8;  // looping variant: 128-register kernels, keep several CTAs resident
    }
    threads = warps * 32;
    bool ok = smem <= (size_t)di.max_smem_optin && threads <= fa.maxThreadsPerBlock &&
              threads * fa.numRegs <= 65536;
    if (ok) {
      if (smem > 48 * 1024)
        MVAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MVAE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, threads, smem));
      ok = blocks_per_sm >= 1;
    }
    if (ok && (S == 32 || blocks_per_sm * threads >= 768)) break;  // S = 64 only if it keeps >= 24 warps per SM
    if (S == 32) return MVAE_ERR_UNSUPPORTED;
  }
  p.S = S;
  const int64_t tiles = (p.B + S - 1) / S;
  if (tiles > 0x7fffffff) return MVAE_ERR_UNSUPPORTED;
  p.n_tiles = (int)tiles;
  int64_t grid = (int64_t)blocks_per_sm * di.sm_count;
  if (grid > tiles) grid = tiles;
  kern<<<(unsigned)grid, threads, smem, as_stream(stream)>>>(p);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

}  // namespace mvae
