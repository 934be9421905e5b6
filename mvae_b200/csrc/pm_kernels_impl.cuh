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
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pm_item.cuh"
#include "pm_params.cuh"

#ifndef MVAE_PM_BWD
#error "define MVAE_PM_BWD to 0 (forward kernels) or 1 (backward kernels) before including pm_kernels_impl.cuh"
#endif

namespace mvae {

// ------------------------------------------------------------------------------------------------ tile layout
// Shared memory (floats): [ItemInfo x C] | nst x input stage | 2 x output stage | nst mbarriers.  Every tile is a
// multiple of 128 bytes (S is a multiple of 32) so all bulk copies are 16-byte aligned.  The host computes every
// offset (fill_layout) so that the kernel reads them as constant-bank operands.
constexpr int kPmMaxStages = 8;

static inline size_t fill_layout(PmParams& p, bool bwd, bool want_ms, bool has_gkl) {
  const int C = p.desc.C, ld_ml = p.desc.ld_ml, ld_eps = p.desc.ld_eps, ld_z = p.desc.ld_z, S = p.S;
  const int info_floats = (C * (int)(sizeof(ItemInfo) / 4) + 31) & ~31;
  int o = 0;
  p.in_ml = o;
  o += S * ld_ml;
  p.in_eps = o;
  o += S * ld_eps;
  p.in_gz = o;
  if (bwd) o += S * ld_z;
  p.in_gkl = o;
  if (bwd && has_gkl) o += S * C;
  p.in_stage = o;
  o = 0;
  p.out_a = o;
  o += bwd ? S * ld_ml : S * ld_z;
  p.out_b = o;
  if (!bwd) o += S * C;
  p.out_c = o;
  if (!bwd && want_ms) o += S * ld_z;
  p.out_d = o;
  if (!bwd && want_ms) o += S * ld_eps;
  p.out_stage = o;
  p.in_base = info_floats;
  p.out_base = p.in_base + p.nst * p.in_stage;
  p.bar_base = p.out_base + 2 * p.out_stage;
  p.bytes_ml = 4u * (uint32_t)(S * ld_ml);
  p.bytes_eps = 4u * (uint32_t)(S * ld_eps);
  p.bytes_z = 4u * (uint32_t)(S * ld_z);
  p.bytes_c = 4u * (uint32_t)(S * C);
  p.bytes_in = 4u * (uint32_t)p.in_stage;
  return (size_t)(p.bar_base + 2 * kPmMaxStages) * 4;
}

// The persistent tile loop.  TYPE >= 0: this warp keeps ONE component of static (TYPE, N) for the whole launch
// (blockDim = 32 x the warps of all components) — no dispatch, every shared-memory offset final.  TYPE < 0: looping variant (more
// warp-items than warps: many components), items are dispatched one by one.
template <bool BWD, int MAXN, bool WANT_MS, int TYPE, int N>
__device__ __forceinline__ void pm_tile_loop(const PmParams& p, float* smem, const ItemInfo* info, int my_ci,
                                             uint32_t smem0, uint32_t bar0) {
  constexpr bool SINGLE = TYPE >= 0;
  const int C = p.desc.C;
  const int ld_ml = p.desc.ld_ml, ld_eps = p.desc.ld_eps, ld_z = p.desc.ld_z;
  const bool has_gkl = BWD && p.gkl != nullptr;
  const bool check = !BWD && p.flag != nullptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  auto issue_loads = [&](int tile, int st) {  // one thread
    const int64_t row0 = (int64_t)tile * p.S;
    const uint32_t bar = bar0 + 8u * st;
    const uint32_t dst = smem0 + 4u * (uint32_t)(p.in_base + st * p.in_stage);
    pm_mbar_expect_tx(bar, p.bytes_in);
    pm_bulk_g2s(dst + 4u * p.in_ml, p.ml + row0 * ld_ml, p.bytes_ml, bar);
    pm_bulk_g2s(dst + 4u * p.in_eps, p.eps + row0 * ld_eps, p.bytes_eps, bar);
    if (BWD) {
      pm_bulk_g2s(dst + 4u * p.in_gz, p.gz + row0 * ld_z, p.bytes_z, bar);
      if (has_gkl) pm_bulk_g2s(dst + 4u * p.in_gkl, p.gkl + row0 * C, p.bytes_c, bar);
    }
  };

  // this thread's operand offsets (floats from smem[0]) inside stage 0 / staging buffer 0, block 0
  ItemInfo mine;
  int r0 = lane;
  int o_m = 0, o_l = 0, o_e = 0, o_gz = 0, o_gkl = 0, o_z = 0, o_kl = 0, o_mu = 0, o_sg = 0, o_gm = 0, o_gl = 0;
  if (SINGLE) {
    mine = info[my_ci];
    r0 = ((int)p.w_blk0[warp] << 5) + lane;
    o_m = p.in_base + p.in_ml + r0 * ld_ml + mine.m_off;
    o_l = p.in_base + p.in_ml + r0 * ld_ml + mine.l_off;
    o_e = p.in_base + p.in_eps + r0 * ld_eps + mine.eps_off;
    o_gz = p.in_base + p.in_gz + r0 * ld_z + mine.z_off;
    o_gkl = p.in_base + p.in_gkl + r0 * C + my_ci;
    o_z = p.out_base + p.out_a + r0 * ld_z + mine.z_off;
    o_kl = p.out_base + p.out_b + r0 * C + my_ci;
    o_mu = p.out_base + p.out_c + r0 * ld_z + mine.z_off;
    o_sg = p.out_base + p.out_d + r0 * ld_eps + mine.eps_off;
    o_gm = p.out_base + p.out_a + r0 * ld_ml + mine.m_off;
    o_gl = p.out_base + p.out_a + r0 * ld_ml + mine.l_off;
  }
  // per-warp block walk: nblk blocks, srows rows apart
  const int nblk = SINGLE ? (int)p.w_nblk[warp] : 0;
  const int srows = SINGLE ? (int)p.w_bstride[warp] << 5 : 0;
  const int rs_ml = srows * ld_ml, rs_eps = srows * ld_eps, rs_z = srows * ld_z, rs_c = srows * C;
  float* const b_m = smem + o_m;
  float* const b_l = smem + o_l;
  float* const b_e = smem + o_e;
  float* const b_gz = smem + o_gz;
  float* const b_gkl = smem + o_gkl;
  float* const b_z = smem + o_z;
  float* const b_kl = smem + o_kl;
  float* const b_mu = smem + o_mu;
  float* const b_sg = smem + o_sg;
  float* const b_gm = smem + o_gm;
  float* const b_gl = smem + o_gl;
  float gR_acc = 0.f;
  bool finite = true;

  int tile = blockIdx.x;
  if (tid == 0) {  // prologue: nst - 1 tiles in flight
    int t = tile;
    for (int i = 0; i < p.nst - 1 && t < p.n_bulk_tiles; ++i, t += gridDim.x) issue_loads(t, i);
  }
  int st = 0;
  int in_off = 0, out_off = 0;  // float offsets of the current input stage / output staging buffer
  uint32_t parity = 0;
  for (; tile < p.n_tiles; tile += gridDim.x) {
    const bool bulk = tile < p.n_bulk_tiles;
    int rows = p.S;
    if (tid == 0) {
      // the stage consumed in the previous iteration is free: refill it with the tile nst - 1 ahead
      const int ahead = tile + (p.nst - 1) * (int)gridDim.x;  // n_tiles * 8 stays far below 2^31 (host check)
      if (ahead < p.n_bulk_tiles) issue_loads(ahead, st == 0 ? p.nst - 1 : st - 1);
    }
    if (bulk) {
      pm_mbar_wait(bar0 + 8u * st, parity);
    } else {
      const int64_t row0 = (int64_t)tile * p.S;
      rows = (int)min((int64_t)p.S, p.B - row0);
      float* sin_ = smem + p.in_base + in_off;
      coop_load(sin_ + p.in_ml, p.ml + row0 * ld_ml, rows * ld_ml);
      coop_load(sin_ + p.in_eps, p.eps + row0 * ld_eps, rows * ld_eps);
      if (BWD) {
        coop_load(sin_ + p.in_gz, p.gz + row0 * ld_z, rows * ld_z);
        if (has_gkl) coop_load(sin_ + p.in_gkl, p.gkl + row0 * C, rows * C);
      }
      __syncthreads();
    }
    if (BWD && p.zero_gml) {
      float* g0 = smem + p.out_base + out_off + p.out_a;
      for (int i = tid; i < p.S * ld_ml; i += blockDim.x) g0[i] = 0.f;
      __syncthreads();
    }
    // ---- compute ----
    if constexpr (SINGLE) {
      // operand pointers are loop-carried (one add per block) so that the shared-memory base is formed once
      const float* pm = b_m + in_off;
      const float* pl = b_l + in_off;
      const float* pe = b_e + in_off;
      const float* pgz = b_gz + in_off;
      const float* pgkl = b_gkl + in_off;
      float* pz = b_z + out_off;
      float* pkl = b_kl + out_off;
      float* pmu = b_mu + out_off;
      float* psg = b_sg + out_off;
      float* pgm = b_gm + out_off;
      float* pgl = b_gl + out_off;
      int row = r0;
#pragma unroll 1
      for (int r = 0; r < nblk; ++r) {
        if (row < rows) {
          const float gkl = BWD ? (has_gkl ? *pgkl : p.gkl_scalar) : 0.f;
          finite &= run_item<N, TYPE, BWD, WANT_MS>(mine.n, mine.l_n, mine.K, pm, pl, pe, pz, pkl, pmu, psg, pgz, gkl,
                                                    pgm, pgl, &gR_acc, check);
        }
        row += srows;
        pm += rs_ml;
        pl += rs_ml;
        pe += rs_eps;
        if (BWD) {
          pgz += rs_z;
          pgkl += rs_c;
          pgm += rs_ml;
          pgl += rs_ml;
        } else {
          pz += rs_z;
          pkl += rs_c;
          if (WANT_MS) {
            pmu += rs_z;
            psg += rs_eps;
          }
        }
      }
    } else {
      const int nwarps = blockDim.x >> 5, blocks = p.nb;  // 32-sample blocks per tile
      float* sin_ = smem + p.in_base + in_off;
      float* sout = smem + p.out_base + out_off;
      for (int w = warp; w < C * blocks; w += nwarps) {
        const int ci = w / blocks;
        const int sidx = ((w - ci * blocks) << 5) + lane;
        float gR = 0.f;
        if (sidx < rows) {
          const ItemInfo c = info[ci];
          const float gkl = BWD ? (has_gkl ? sin_[p.in_gkl + sidx * C + ci] : p.gkl_scalar) : 0.f;
          finite &= dispatch_item<BWD, MAXN, WANT_MS>(
              c, sin_ + p.in_ml + sidx * ld_ml, sin_ + p.in_eps + sidx * ld_eps, sout + p.out_a + sidx * ld_z,
              sout + p.out_b + sidx * C + ci, sout + p.out_c + sidx * ld_z, sout + p.out_d + sidx * ld_eps,
              sin_ + p.in_gz + sidx * ld_z, gkl, sout + p.out_a + sidx * ld_ml, &gR, check);
          gR *= c.dfac;
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
      if (tid == 0) pm_bulk_wait_read0();  // the store of the previous tile (other staging buffer) has left shared memory
      __syncthreads();
      if (tid == 0) {
        const int64_t row0 = (int64_t)tile * p.S;
        const uint32_t src = smem0 + 4u * (uint32_t)(p.out_base + out_off);
        if (BWD) {
          pm_bulk_s2g(p.gml + row0 * ld_ml, src + 4u * p.out_a, p.bytes_ml);
        } else {
          pm_bulk_s2g(p.z + row0 * ld_z, src + 4u * p.out_a, p.bytes_z);
          pm_bulk_s2g(p.kl + row0 * C, src + 4u * p.out_b, p.bytes_c);
          if (WANT_MS) {
            pm_bulk_s2g(p.mu + row0 * ld_z, src + 4u * p.out_c, p.bytes_z);
            pm_bulk_s2g(p.sigma + row0 * ld_eps, src + 4u * p.out_d, p.bytes_eps);
          }
        }
        pm_bulk_commit();
      }
    } else {
      const int64_t row0 = (int64_t)tile * p.S;
      const float* sout = smem + p.out_base + out_off;
      __syncthreads();
      if (BWD) {
        coop_store(p.gml + row0 * ld_ml, sout + p.out_a, rows * ld_ml);
      } else {
        coop_store(p.z + row0 * ld_z, sout + p.out_a, rows * ld_z);
        coop_store(p.kl + row0 * C, sout + p.out_b, rows * C);
        if (WANT_MS) {
          coop_store(p.mu + row0 * ld_z, sout + p.out_c, rows * ld_z);
          coop_store(p.sigma + row0 * ld_eps, sout + p.out_d, rows * ld_eps);
        }
      }
      __syncthreads();
    }
    in_off += p.in_stage;
    if (++st == p.nst) {
      st = 0;
      in_off = 0;
      parity ^= 1u;
    }
    out_off = out_off ? 0 : p.out_stage;
  }
  if (tid == 0) pm_bulk_wait_all0();  // shared memory must outlive the last bulk store
  if (BWD) {
    if (SINGLE && p.gradius) {
      const float g = warp_sum(gR_acc * mine.dfac);
      if (lane == 0 && g != 0.f) atomicAdd(p.gradius + my_ci, g);
    }
  } else if (p.flag) {
    const unsigned bad = __ballot_sync(0xffffffffu, !finite);
    if (bad && lane == 0) atomicOr(p.flag, 1u);
  }
}

template <bool BWD, int MAXN, bool WANT_MS, bool SINGLE>
__device__ __forceinline__ void pm_kernel_body(const PmParams& p) {
  extern __shared__ __align__(128) float smem[];
  const int C = p.desc.C;
  ItemInfo* info = reinterpret_cast<ItemInfo*>(smem);
  const uint32_t smem0 = pm_smem_u32(smem);
  const uint32_t bar0 = smem0 + 4u * (uint32_t)p.bar_base;
  const int tid = threadIdx.x;
  stage_items(info, p.desc, p.radius);
  if (tid == 0) {
    for (int i = 0; i < p.nst; ++i) pm_mbar_init(bar0 + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if constexpr (SINGLE) {
    // Every warp of the CTA runs the same number of tile iterations, so the __syncthreads inside the per-(type, n)
    // copies of the loop pair up across warps that took different cases.
    const int my_ci = p.w_ci[tid >> 5];
    const int type = info[my_ci].type, n = info[my_ci].n;
#define MVAE_PM_LOOP_E(NN) pm_tile_loop<BWD, MAXN, WANT_MS, MVAE_EUCLIDEAN, NN>(p, smem, info, my_ci, smem0, bar0);
#define MVAE_PM_LOOP_H(NN) pm_tile_loop<BWD, MAXN, WANT_MS, MVAE_HYPERBOLOID, NN>(p, smem, info, my_ci, smem0, bar0);
#define MVAE_PM_LOOP_S(NN) pm_tile_loop<BWD, MAXN, WANT_MS, MVAE_SPHERE, NN>(p, smem, info, my_ci, smem0, bar0);
#define MVAE_PM_LOOP_P(NN) pm_tile_loop<BWD, MAXN, WANT_MS, MVAE_POINCARE, NN>(p, smem, info, my_ci, smem0, bar0);
#define MVAE_PM_LOOP_D(NN) pm_tile_loop<BWD, MAXN, WANT_MS, MVAE_PROJ_SPHERE, NN>(p, smem, info, my_ci, smem0, bar0);
    switch (type) {
      case MVAE_EUCLIDEAN: MVAE_PM_FOR_DIMS(MAXN, n, MVAE_PM_LOOP_E) break;
      case MVAE_HYPERBOLOID: MVAE_PM_FOR_DIMS(MAXN, n, MVAE_PM_LOOP_H) break;
      case MVAE_SPHERE: MVAE_PM_FOR_DIMS(MAXN, n, MVAE_PM_LOOP_S) break;
      case MVAE_POINCARE: MVAE_PM_FOR_DIMS(MAXN, n, MVAE_PM_LOOP_P) break;
      default: MVAE_PM_FOR_DIMS(MAXN, n, MVAE_PM_LOOP_D) break;
    }
#undef MVAE_PM_LOOP_D
#undef MVAE_PM_LOOP_E
#undef MVAE_PM_LOOP_H
#undef MVAE_PM_LOOP_S
#undef MVAE_PM_LOOP_P
  } else {
    pm_tile_loop<BWD, MAXN, WANT_MS, -1, 0>(p, smem, info, 0, smem0, bar0);
  }
}

// Register budgets (launch bounds): the kernels are latency-sensitive, so the narrow-dimension variants are held to
// 48 / 64 / 80 registers to keep 32+ warps resident; the widest ones take what they need.
#ifndef MVAE_PM_MINB
#define MVAE_PM_MINB(MAXN, BWD) ((MAXN) == 2 ? ((BWD) ? 4 : 5) : (MAXN) == 4 ? ((BWD) ? 3 : 4) : (MAXN) == 8 ? ((BWD) ? 2 : 3) : 1)
#endif
#if !MVAE_PM_BWD
template <int MAXN, bool WANT_MS, bool SINGLE>
__global__ void __launch_bounds__(kPmMaxThreads, MVAE_PM_MINB(MAXN, false))
    pm_forward_kernel(const __grid_constant__ PmParams p) {
  pm_kernel_body<false, MAXN, WANT_MS, SINGLE>(p);
}
#else
template <int MAXN, bool SINGLE>
__global__ void __launch_bounds__(kPmMaxThreads, MVAE_PM_MINB(MAXN, true))
    pm_backward_kernel(const __grid_constant__ PmParams p) {
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
  // SINGLE variant: one warp per (component, 32-sample block column) when C warps fit a CTA; else the looping variant
  // with all dimensions compiled in.
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
  // Tile shape.  A tile is nb blocks of 32 samples.  One-component-per-warp variant: component c gets w_c warps
  // (w_c divides nb), each walking nb / w_c blocks; the w_c are chosen so that the warps of a CTA finish a tile
  // together (arithmetic cost model below) — warps that wait at the tile barrier for the slowest component are lost
  // issue capacity.  More blocks per tile amortise the per-tile costs (barrier, mbarrier wait, the issuing thread's
  // copies); the ring depth nst keeps enough bulk copies in flight to cover HBM latency.  Small batches want many small
  // tiles instead (one per SM at least).
  cudaFuncAttributes fa;
  MVAE_CUDA_TRY(cudaFuncGetAttributes(&fa, kern));
  auto cost_of = [&](const mvae_component& c) {  // issue slots per item, from the SASS of the static variants
    const int base = c.type == MVAE_EUCLIDEAN ? 20 : c.type == MVAE_SPHERE ? 125 : c.type == MVAE_POINCARE ? 115 :
                     (c.type == MVAE_PROJ_SPHERE || c.type == MVAE_UNIVERSAL) ? 140 : 105;
    return (bwd ? 3 : 2) * (base + (c.type == MVAE_EUCLIDEAN ? 15 : 25) * c.n) / 2;
  };
  int wc[MVAE_MAX_COMPONENTS];
  // best split of <= kPmMaxWarps warps over the components for nb blocks per tile; returns efficiency in 1/1024
  auto plan = [&](int nb, int* w_out) {
    int w[MVAE_MAX_COMPONENTS], total = D.C;
    for (int i = 0; i < D.C; ++i) w[i] = 1;
    int best_eff = -1;
    for (;;) {
      int sum_cost = 0, max_load = 0, arg = 0;
      for (int i = 0; i < D.C; ++i) {
        const int load = cost_of(D.comp[i]) * (nb / w[i]);
        sum_cost += cost_of(D.comp[i]) * nb;
        if (load > max_load) {
          max_load = load;
          arg = i;
        }
      }
      const int eff = (int)((int64_t)sum_cost * 1024 / ((int64_t)total * max_load));
      if (eff > best_eff) {
        best_eff = eff;
        for (int i = 0; i < D.C; ++i) w_out[i] = w[i];
      }
      // give the most loaded component the next divisor of nb
      int nw = w[arg] + 1;
      while (nw <= nb && nb % nw) ++nw;
      if (nw > nb || total - w[arg] + nw > kPmMaxWarps) break;
      total += nw - w[arg];
      w[arg] = nw;
    }
    return best_eff;
  };
  int nb = 1;
  if (single) {
    const int bytes_per_block = 32 * 4 * (D.ld_ml + D.ld_eps + D.ld_z + D.C + (bwd ? D.ld_ml : 0));
    int best = -1;
    for (int cand = 1; cand <= 8; ++cand) {
      if (cand * bytes_per_block > 40 * 1024 && cand > 1) break;       // tiles of wide products are big already
      if (p.B < (int64_t)32 * cand * 4 * di.sm_count && cand > 1) break;  // small batch: smallest tiles
      int w_try[MVAE_MAX_COMPONENTS];
      const int eff = plan(cand, w_try) + 16 * cand;  // ties (and near-ties) go to the bigger tile
      if (eff > best) {
        best = eff;
        nb = cand;
      }
    }
  }
  // For the chosen nb (shrunk if shared memory would cap residency below ~24 warps per SM or below what the registers
  // allow): the shallowest ring whose copies in flight on one SM reach 64 KiB (HBM latency x bandwidth per SM).
  int S = 32 * nb, threads = 0, blocks_per_sm = 0;
  size_t smem = 0;
  auto configure = [&](int nb_, int nst_, int* blocks, size_t* bytes) -> int {  // fills p for (nb_, nst_)
    p.S = 32 * nb_;
    p.nb = nb_;
    p.nst = nst_;
    *bytes = fill_layout(p, bwd, want_ms, has_gkl);
    int warps = 0;
    if (single) {
      plan(nb_, wc);
      for (int i = 0; i < D.C; ++i)
        for (int j = 0; j < wc[i]; ++j, ++warps) {
          p.w_ci[warps] = (uint8_t)i;
          p.w_blk0[warps] = (uint8_t)j;
          p.w_bstride[warps] = (uint8_t)wc[i];
          p.w_nblk[warps] = (uint8_t)(nb_ / wc[i]);
        }
    } else {
      warps = D.C * nb_;
      if (warps > kPmMaxWarps) warps = kPmMaxWarps;  // looping variant: 128-register kernels, several CTAs resident
    }
    p.n_warps = warps;
    const int th = warps * 32;
    *blocks = 0;
    if (*bytes > (size_t)di.max_smem_optin || th > fa.maxThreadsPerBlock || th * fa.numRegs > 65536) return MVAE_OK;
    if (*bytes > 48 * 1024)
      MVAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*bytes));
    MVAE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, kern, th, *bytes));
    return MVAE_OK;
  };
  int forced_nst = 0;
  {
    static const char* tune = getenv("MVAE_PM_TUNE");  // tuning aid: "nb,nst"
    int a = 0, b = 0;
    if (tune && sscanf(tune, "%d,%d", &a, &b) == 2 && a >= 1 && a <= 8 && b >= 2 && b <= kPmMaxStages) {
      nb = a;
      forced_nst = b;
    }
  }
  int nst = 2;
  for (;; --nb) {
    int blocks = 0, reg_blocks = 0, pick = 0;
    size_t bytes = 0;
    rc = configure(nb, 2, &blocks, &bytes);
    if (rc != MVAE_OK) return rc;
    if (blocks >= 1) {
      const int th = p.n_warps * 32;
      MVAE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&reg_blocks, kern, th, 0));
      int want = (768 + th - 1) / th;
      if (want > reg_blocks) want = reg_blocks;
      for (int cand = 2; cand <= (forced_nst ? forced_nst : 4); ++cand) {
        rc = configure(nb, cand, &blocks, &bytes);
        if (rc != MVAE_OK) return rc;
        if (blocks < want && !(nb == 1 && cand == 2 && blocks >= 1)) break;
        pick = cand;
        if (!forced_nst && (int64_t)(cand - 1) * p.bytes_in * blocks >= 64 * 1024) break;
      }
    }
    if (pick) {
      nst = pick;
      break;
    }
    if (nb == 1) return MVAE_ERR_UNSUPPORTED;
  }
  rc = configure(nb, nst, &blocks_per_sm, &smem);
  if (rc != MVAE_OK) return rc;
  S = 32 * nb;
  threads = p.n_warps * 32;
  {
    static const bool dbg = getenv("MVAE_PM_DEBUG") != nullptr;
    if (dbg) {
      fprintf(stderr, "[mvae pm %s] B=%lld C=%d single=%d nb=%d nst=%d warps=%d regs=%d smem=%zu blocks/SM=%d w=", bwd ? "bwd" : "fwd",
              (long long)p.B, D.C, (int)single, nb, nst, p.n_warps, fa.numRegs, smem, blocks_per_sm);
      if (single)
        for (int i = 0; i < D.C; ++i) fprintf(stderr, "%d%s", wc[i], i + 1 < D.C ? "," : "");
      fprintf(stderr, "\n");
    }
  }
  const int64_t tiles = (p.B + S - 1) / S;
  if (tiles > (0x7fffffff >> 4)) return MVAE_ERR_UNSUPPORTED;
  p.n_tiles = (int)tiles;
  p.n_bulk_tiles = p.vec_ok ? (int)(p.B / S) : 0;
  int64_t grid = (int64_t)blocks_per_sm * di.sm_count;
  if (grid > tiles) grid = tiles;
  kern<<<(unsigned)grid, threads, smem, as_stream(stream)>>>(p);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

}  // namespace mvae
