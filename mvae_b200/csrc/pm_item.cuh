// pm_item.cuh — shared by the fused product-manifold kernels (pm_kernels_impl.cuh) and the fused latent-block kernels
// (latent_fused.cu): bulk-copy / mbarrier PTX wrappers, the per-component descriptor staged in shared memory, and
// run_item / dispatch_item, which run ONE (component, sample) through pm_math.cuh with every operand in shared memory.
#pragma once
#include <math.h>
#include <stdint.h>

#include "mvae_common.cuh"
#include "pm_math.cuh"

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
  float dfac;   // d(radius)/d(raw parameter): clamp(relu(.)) for radius parameters, d|kappa|^(-1/2)/dkappa for 'u'
};

// One (component, sample) item with every operand pointer final (component offsets applied): m, l, e [, gz] in the input
// stage; z, kl [, mu, sigma] or gm, gl in the output staging tile.  Returns false if a produced value is non-finite
// (only evaluated when `check`).
template <int N, int TYPE, bool BWD, bool WANT_MS>
__device__ __forceinline__ bool run_item(int n_rt, int l_n, const CompConst& K, const float* m, const float* l,
                                         const float* e, float* z, float* kl, float* mu, float* sigma,
                                         const float* gz, float gkl, float* gm_out, float* gl_out, float* gR_acc,
                                         bool check) {
  CompOut<N> o;
  const int n = N > 0 ? N : n_rt;
  constexpr int CN = Cap<N>::n;
  constexpr bool AMB = TYPE == MVAE_HYPERBOLOID || TYPE == MVAE_SPHERE;  // ambient dimension n + 1
  float gm[BWD ? CN : 1], gl[BWD ? CN : 1];
  float gR = 0.f;
  if (TYPE == MVAE_EUCLIDEAN) comp_e<N, BWD>(n, l_n, m, l, e, o, gz, gkl, gm, gl);
  else if (TYPE == MVAE_HYPERBOLOID) comp_hsp<N, BWD, kHyp, WANT_MS>(n, l_n, m, l, e, K, o, gz, gkl, gm, gl, &gR);
  else if (TYPE == MVAE_SPHERE) comp_hsp<N, BWD, kSph, WANT_MS>(n, l_n, m, l, e, K, o, gz, gkl, gm, gl, &gR);
  else if (TYPE == MVAE_POINCARE) comp_hsp<N, BWD, kPoi, WANT_MS>(n, l_n, m, l, e, K, o, gz, gkl, gm, gl, &gR);
  else comp_hsp<N, BWD, kPsp, WANT_MS>(n, l_n, m, l, e, K, o, gz, gkl, gm, gl, &gR);
  if (BWD) {
    *gR_acc += gR;
    if (N > 0) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        gm_out[j] = gm[j];
        if (j < l_n) gl_out[j] = gl[j];
      }
    } else {
      for (int j = 0; j < n; ++j) {
        gm_out[j] = gm[j];
        if (j < l_n) gl_out[j] = gl[j];
      }
    }
    return true;
  }
  const int d = AMB ? n + 1 : n;
  if (N > 0) {
#pragma unroll
    for (int k = 0; k < N + 1; ++k)
      if (k < d) {
        z[k] = o.z[k];
        if (WANT_MS) mu[k] = o.mu[k];
      }
    if (WANT_MS) {
#pragma unroll
      for (int j = 0; j < N; ++j) sigma[j] = o.sigma[j];
    }
  } else {
    for (int k = 0; k < d; ++k) {
      z[k] = o.z[k];
      if (WANT_MS) mu[k] = o.mu[k];
    }
    if (WANT_MS)
      for (int j = 0; j < n; ++j) sigma[j] = o.sigma[j];
  }
  *kl = o.kl;
  if (!check) return true;
  // KL is a function of every produced value except, for Euclidean components, of eps (z = mu + eps sigma): there
  // the z coordinates join the sum.  A sum is non-finite iff a term is (up to overflow near FLT_MAX).
  float chk = o.kl;
  if (TYPE == MVAE_EUCLIDEAN) {
    if (N > 0) {
#pragma unroll
      for (int k = 0; k < N; ++k) chk += o.z[k];
    } else {
      for (int k = 0; k < d; ++k) chk += o.z[k];
    }
  }
  return chk - chk == 0.f;
}

// Runtime (type, n) -> static instantiation.  MAXN > 0: only dimensions <= MAXN are compiled in (register budget =
// that of the widest one); MAXN == 0: all static dimensions plus the runtime-dimension path.
#define MVAE_PM_FOR_DIMS(MAXN, n, CALL)                                     \
  switch (n) {                                                              \
    case 1: CALL(1) break;                                                  \
    case 2: CALL(2) break;                                                  \
    case 3: if constexpr (MAXN == 0 || MAXN >= 3) { CALL(3) } break;        \
    case 4: if constexpr (MAXN == 0 || MAXN >= 4) { CALL(4) } break;        \
    case 5: if constexpr (MAXN == 0 || MAXN >= 5) { CALL(5) } break;        \
    case 6: if constexpr (MAXN == 0 || MAXN >= 6) { CALL(6) } break;        \
    case 8: if constexpr (MAXN == 0 || MAXN >= 8) { CALL(8) } break;        \
    default: if constexpr (MAXN == 0) { CALL(0) } break;                    \
  }

// looping variant: dispatch per item
template <bool BWD, int MAXN, bool WANT_MS>
__device__ __forceinline__ bool dispatch_item(const ItemInfo& c, const float* ml_row, const float* eps_row,
                                              float* z_row, float* kl_slot, float* mu_row, float* sigma_row,
                                              const float* gz_row, float gkl, float* gml_row, float* gR_acc,
                                              bool check) {
  bool ok = true;
#define MVAE_PM_ITEM(TT, NN)                                                                                          \
  ok = run_item<NN, TT, BWD, WANT_MS>(c.n, c.l_n, c.K, ml_row + c.m_off, ml_row + c.l_off, eps_row + c.eps_off,       \
                                      z_row + c.z_off, kl_slot, mu_row + c.z_off, sigma_row + c.eps_off,              \
                                      gz_row + c.z_off, gkl, gml_row + c.m_off, gml_row + c.l_off, gR_acc, check);
#define MVAE_PM_E(NN) MVAE_PM_ITEM(MVAE_EUCLIDEAN, NN)
#define MVAE_PM_H(NN) MVAE_PM_ITEM(MVAE_HYPERBOLOID, NN)
#define MVAE_PM_S(NN) MVAE_PM_ITEM(MVAE_SPHERE, NN)
#define MVAE_PM_P(NN) MVAE_PM_ITEM(MVAE_POINCARE, NN)
#define MVAE_PM_D(NN) MVAE_PM_ITEM(MVAE_PROJ_SPHERE, NN)
  switch (c.type) {
    case MVAE_EUCLIDEAN: MVAE_PM_FOR_DIMS(MAXN, c.n, MVAE_PM_E) break;
    case MVAE_HYPERBOLOID: MVAE_PM_FOR_DIMS(MAXN, c.n, MVAE_PM_H) break;
    case MVAE_SPHERE: MVAE_PM_FOR_DIMS(MAXN, c.n, MVAE_PM_S) break;
    case MVAE_POINCARE: MVAE_PM_FOR_DIMS(MAXN, c.n, MVAE_PM_P) break;
    default: MVAE_PM_FOR_DIMS(MAXN, c.n, MVAE_PM_D) break;
  }
#undef MVAE_PM_D
#undef MVAE_PM_E
#undef MVAE_PM_H
#undef MVAE_PM_S
#undef MVAE_PM_P
#undef MVAE_PM_ITEM
  return ok;
}

// Stage the per-component descriptors (+ curvature constants from the raw radius parameters) in shared memory.
__device__ __forceinline__ void stage_items(ItemInfo* info, const mvae_pm_desc& desc, const float* radius) {
  for (int i = threadIdx.x; i < desc.C; i += blockDim.x) {
    const mvae_component c = desc.comp[i];
    ItemInfo ii;
    ii.type = c.type;
    ii.n = c.n;
    ii.l_n = c.l_n;
    ii.d = c.d;
    ii.m_off = c.m_off;
    ii.l_off = c.l_off;
    ii.eps_off = c.eps_off;
    ii.z_off = c.z_off;
    ii.rp = (radius && c.type != MVAE_EUCLIDEAN) ? __ldg(radius + i) : 1.f;
    if (c.type == MVAE_UNIVERSAL) {
      // universal.py:64-74: the manifold is chosen by the sign of the learnable curvature (the reference does it with a
      // host-side `if` on every call); radius = relu(1 / sqrt(|kappa|)) (:30-31)
      const float kappa = ii.rp, ak = fabsf(kappa);
      if (ak > 1e-6f) {
        ii.type = kappa < 0.f ? MVAE_POINCARE : MVAE_PROJ_SPHERE;
        ii.rp = rsqrtf(ak);
        ii.dfac = (kappa < 0.f ? 0.5f : -0.5f) * ii.rp / ak;
      } else {
        ii.type = MVAE_EUCLIDEAN;
        ii.rp = 1.f;
        ii.dfac = 0.f;
      }
    } else {
      ii.dfac = radius_d(ii.rp);
    }
    ii.K = make_const(ii.rp);
    info[i] = ii;
  }
}

}  // namespace mvae
