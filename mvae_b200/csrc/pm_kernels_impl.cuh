// pm_kernels.cu — the fused per-sample product-manifold kernels (K3 of SURVEY.md §2.2) behind
// mvae_pm_forward / mvae_pm_backward (include/mvae_b200.h).
//
// Data movement (HBM-bound design): a CTA owns a tile of S consecutive samples.  The tile's rows of
// `ml`, `eps` (and `gz`) are contiguous in HBM, so they are fetched with coalesced 128-bit loads and scattered
// into shared memory with an ODD row stride (bank-conflict-free when a thread later walks one sample's row).
// The per-component descriptor and the clamped radii R_c are staged in shared memory once per CTA.  Work items
// are (component, sample) pairs laid out component-major, so a warp always executes one manifold type at one
// dimension (no divergence) and all C components of a sample proceed in parallel.  Results go back through
// shared memory and leave with coalesced 128-bit stores.  Nothing is read twice from HBM.
#pragma once
#include <math.h>
#include <string.h>

#include "manifold_math.cuh"
#include "pm_params.cuh"

#ifndef MVAE_PM_BWD
#error "define MVAE_PM_BWD to 0 (forward kernels) or 1 (backward kernels) before including pm_kernels_impl.cuh"
#endif

namespace mvae {

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w));
}

// rows x ld floats, contiguous in global at g  <->  the same dense layout in shared memory: straight 128-bit copies,
// no index arithmetic (the kernel is instruction-issue bound; bank conflicts of the later strided row reads cost less
// than scattering into padded rows).
__device__ __forceinline__ void tile_load(float* __restrict__ s, const float* __restrict__ g, int total, bool vec) {
  int done = 0;
  if (vec) {
    const int nvec = total >> 2;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x)
      reinterpret_cast<float4*>(s)[i] = ldg_stream4(g + 4 * i);
    done = nvec << 2;
  }
  for (int i = done + threadIdx.x; i < total; i += blockDim.x) s[i] = __ldg(g + i);
}

__device__ __forceinline__ void tile_store(float* __restrict__ g, const float* __restrict__ s, int total, bool vec) {
  int done = 0;
  if (vec) {
    const int nvec = total >> 2;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) stg_stream4(g + 4 * i, reinterpret_cast<const float4*>(s)[i]);
    done = nvec << 2;
  }
  for (int i = done + threadIdx.x; i < total; i += blockDim.x) g[i] = s[i];
}

// One (component, sample) item.  Returns false if any produced value is non-finite.
template <int N, bool BWD>
__device__ __forceinline__ bool run_item(const mvae_component& c, float R, float rp, const float* ml_row,
                                         const float* eps_row, float* z_row, float* kl_slot, float* mu_row,
                                         float* sigma_row, const float* gz_row, float gkl, float* gml_row,
                                         float* gR_acc) {
  CompOut<N> o;
  const float* m = ml_row + c.m_off;
  const float* l = ml_row + c.l_off;
  const float* e = eps_row + c.eps_off;
  const float* gz = BWD ? gz_row + c.z_off : nullptr;
  float* gm = BWD ? gml_row + c.m_off : nullptr;
  float* gl = BWD ? gml_row + c.l_off : nullptr;
  float gR = 0.f;
  const int n = N > 0 ? N : c.n;
  switch (c.type) {
    case MVAE_EUCLIDEAN: comp_e<N, BWD>(n, c.l_n, m, l, e, o, gz, gkl, gm, gl); break;
    case MVAE_HYPERBOLOID: comp_h<N, BWD>(n, c.l_n, m, l, e, R, o, gz, gkl, gm, gl, &gR); break;
    case MVAE_SPHERE: comp_s<N, BWD>(n, c.l_n, m, l, e, R, o, gz, gkl, gm, gl, &gR); break;
    default: comp_p<N, BWD>(n, c.l_n, m, l, e, R, o, gz, gkl, gm, gl, &gR); break;
  }
  if (BWD) {
    *gR_acc += gR * radius_d(rp);
    return true;
  }
  bool finite = isfinite(o.kl);
  const int d = c.d;
#pragma unroll
  for (int k = 0; k < Cap<N>::d; ++k)
    if (k < d) {
      z_row[c.z_off + k] = o.z[k];
      finite = finite && isfinite(o.z[k]);
      if (mu_row) mu_row[c.z_off + k] = o.mu[k];
    }
  if (sigma_row) {
#pragma unroll
    for (int j = 0; j < Cap<N>::n; ++j)
      if (j < n) sigma_row[c.eps_off + j] = o.sigma[j];
  }
  *kl_slot = o.kl;
  return finite;
}

template <bool BWD, int MAXN>
__device__ __forceinline__ bool dispatch_item(const mvae_component& c, float R, float rp, const float* ml_row,
                                              const float* eps_row, float* z_row, float* kl_slot, float* mu_row,
                                              float* sigma_row, const float* gz_row, float gkl, float* gml_row,
                                              float* gR_acc) {
  // MAXN > 0: only dimensions <= MAXN are compiled in (register budget = that of the widest one); MAXN == 0: all
  // static dimensions plus the runtime-dimension path.
#define MVAE_CASE(NN)                                                                                              \
  case NN:                                                                                                         \
    if constexpr (MAXN == 0 || NN <= MAXN)                                                                         \
      return run_item<NN, BWD>(c, R, rp, ml_row, eps_row, z_row, kl_slot, mu_row, sigma_row, gz_row, gkl, gml_row, \
                               gR_acc);                                                                            \
    else                                                                                                           \
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
        return run_item<0, BWD>(c, R, rp, ml_row, eps_row, z_row, kl_slot, mu_row, sigma_row, gz_row, gkl, gml_row,
                                gR_acc);
      else
        return true;  // unreachable: the host picks the kernel whose MAXN covers every component
  }
#undef MVAE_CASE
}

// shared-memory carve-up (floats); must match pm_smem_floats()
struct PmSmem {
  mvae_component* comp;
  float* R;
  float* rp;
  float* gR;
  float* ml;
  float* eps;
  float* a;  // fwd: z     bwd: gz
  float* b;  // fwd: kl    bwd: gkl (optional)
  float* c;  // fwd: mu    bwd: gml
  float* d;  // fwd: sigma
};

__device__ __forceinline__ PmSmem carve(const PmParams& p, float* base, bool bwd) {
  PmSmem s;
  const int C = p.desc.C;
  s.comp = reinterpret_cast<mvae_component*>(base);
  float* f = base + C * (int)(sizeof(mvae_component) / sizeof(float));
  s.R = f;
  f += C;
  s.rp = f;
  f += C;
  s.gR = f;
  f += C;
  f += (4 - ((3 * C) & 3)) & 3;  // keep tiles 16-byte aligned
  const int S = p.S;  // multiple of 32, so every tile size is a multiple of 4 floats (16-byte aligned tiles)
  s.ml = f;
  f += S * p.ldp_ml;
  s.eps = f;
  f += S * p.ldp_eps;
  s.a = f;
  f += S * p.ldp_z;
  s.b = f;
  f += S * p.ldp_c;
  s.c = f;
  f += bwd ? S * p.ldp_ml : S * p.ldp_z;
  s.d = f;
  return s;
}

static size_t pm_smem_floats(const mvae_pm_desc& D, int S, bool bwd, int ldp_ml, int ldp_eps, int ldp_z, int ldp_c) {
  size_t f = (size_t)D.C * (sizeof(mvae_component) / sizeof(float)) + 3 * (size_t)D.C + 4;
  f += (size_t)S * (ldp_ml + ldp_eps + ldp_z + ldp_c);
  f += bwd ? (size_t)S * ldp_ml : (size_t)S * (ldp_z + ldp_eps);
  return f;
}

#if !MVAE_PM_BWD
template <int MAXN>
__global__ void __launch_bounds__(kPmThreads) pm_forward_kernel(const __grid_constant__ PmParams p) {
  extern __shared__ __align__(16) float smem[];
  const PmSmem s = carve(p, smem, false);
  const int C = p.desc.C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s.comp[i] = p.desc.comp[i];
    float rp = (p.radius && p.desc.comp[i].type != MVAE_EUCLIDEAN) ? __ldg(p.radius + i) : 1.f;
    s.rp[i] = rp;
    s.R[i] = radius_of(rp);
  }
  const int64_t row0 = (int64_t)blockIdx.x * p.S;
  const int rows = (int)min((int64_t)p.S, p.B - row0);
  const bool vec = p.vec_ok != 0;
  tile_load(s.ml, p.ml + row0 * p.desc.ld_ml, rows * p.desc.ld_ml, vec);
  tile_load(s.eps, p.eps + row0 * p.desc.ld_eps, rows * p.desc.ld_eps, vec);
  __syncthreads();
  const bool want_mu = p.mu != nullptr, want_sigma = p.sigma != nullptr;
  bool finite = true;
  const int items = C * p.S;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int ci = (int)fastdiv((uint32_t)it, p.fd_S);
    const int sidx = it - ci * p.S;
    if (sidx >= rows) continue;
    const mvae_component c = s.comp[ci];
    finite &= dispatch_item<false, MAXN>(c, s.R[ci], s.rp[ci], s.ml + sidx * p.ldp_ml, s.eps + sidx * p.ldp_eps,
                                   s.a + sidx * p.ldp_z, s.b + sidx * p.ldp_c + ci,
                                   want_mu ? s.c + sidx * p.ldp_z : nullptr,
                                   want_sigma ? s.d + sidx * p.ldp_eps : nullptr, nullptr, 0.f, nullptr, nullptr);
  }
  if (p.flag) {
    unsigned bad = __ballot_sync(0xffffffffu, !finite);
    if (bad && (threadIdx.x & 31) == 0) atomicOr(p.flag, 1u);
  }
  __syncthreads();
  tile_store(p.z + row0 * p.desc.ld_z, s.a, rows * p.desc.ld_z, vec);
  tile_store(p.kl + row0 * C, s.b, rows * C, vec);
  if (want_mu) tile_store(p.mu + row0 * p.desc.ld_z, s.c, rows * p.desc.ld_z, vec);
  if (want_sigma) tile_store(p.sigma + row0 * p.desc.ld_eps, s.d, rows * p.desc.ld_eps, vec);
}

#else
template <int MAXN>
__global__ void __launch_bounds__(kPmThreads) pm_backward_kernel(const __grid_constant__ PmParams p) {
  extern __shared__ __align__(16) float smem[];
  const PmSmem s = carve(p, smem, true);
  const int C = p.desc.C;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s.comp[i] = p.desc.comp[i];
    float rp = (p.radius && p.desc.comp[i].type != MVAE_EUCLIDEAN) ? __ldg(p.radius + i) : 1.f;
    s.rp[i] = rp;
    s.R[i] = radius_of(rp);
    s.gR[i] = 0.f;
  }
  const int64_t row0 = (int64_t)blockIdx.x * p.S;
  const int rows = (int)min((int64_t)p.S, p.B - row0);
  const bool vec = p.vec_ok != 0;
  tile_load(s.ml, p.ml + row0 * p.desc.ld_ml, rows * p.desc.ld_ml, vec);
  tile_load(s.eps, p.eps + row0 * p.desc.ld_eps, rows * p.desc.ld_eps, vec);
  tile_load(s.a, p.gz + row0 * p.desc.ld_z, rows * p.desc.ld_z, vec);
  if (p.gkl) tile_load(s.b, p.gkl + row0 * C, rows * C, vec);
  for (int i = threadIdx.x; i < p.S * p.ldp_ml; i += blockDim.x) s.c[i] = 0.f;  // columns no component owns
  __syncthreads();
  const int items = C * p.S;
  // p.S is a multiple of 32 and blockDim.x too, so a warp's 32 items always share one component.
  for (int it0 = (threadIdx.x & ~31); it0 < items; it0 += blockDim.x) {
    const int it = it0 + (threadIdx.x & 31);
    const int ci = (int)fastdiv((uint32_t)it0, p.fd_S);
    const int sidx = it - ci * p.S;
    float gR = 0.f;
    if (sidx < rows) {
      const mvae_component c = s.comp[ci];
      const float gkl = p.gkl ? s.b[sidx * p.ldp_c + ci] : p.gkl_scalar;
      dispatch_item<true, MAXN>(c, s.R[ci], s.rp[ci], s.ml + sidx * p.ldp_ml, s.eps + sidx * p.ldp_eps, nullptr, nullptr,
                          nullptr, nullptr, s.a + sidx * p.ldp_z, gkl, s.c + sidx * p.ldp_ml, &gR);
    }
    gR = warp_sum(gR);
    if ((threadIdx.x & 31) == 0 && gR != 0.f) atomicAdd(&s.gR[ci], gR);
  }
  __syncthreads();
  tile_store(p.gml + row0 * p.desc.ld_ml, s.c, rows * p.desc.ld_ml, vec);
  if (p.gradius)
    for (int i = threadIdx.x; i < C; i += blockDim.x)
      if (s.gR[i] != 0.f) atomicAdd(p.gradius + i, s.gR[i]);
}

#endif  // MVAE_PM_BWD

// Choose the tile height: the largest S in {128, 64, 32} that keeps >= 2 CTAs per SM worth of shared memory and
// still yields at least ~2 waves of CTAs for small batches.
static int pick_tile(const mvae_pm_desc& D, int64_t B, bool bwd, int sm_count, int max_smem, PmParams* p) {
  p->ldp_ml = D.ld_ml;
  p->ldp_eps = D.ld_eps;
  p->ldp_z = D.ld_z;
  p->ldp_c = D.C;
  const int cand[3] = {128, 64, 32};
  int S = 0;
  for (int i = 0; i < 3; ++i) {
    size_t bytes = 4 * pm_smem_floats(D, cand[i], bwd, p->ldp_ml, p->ldp_eps, p->ldp_z, p->ldp_c);
    const bool fits2 = bytes * 2 + 2048 <= (size_t)max_smem + 1024;
    const bool fits1 = bytes <= (size_t)max_smem;
    const int64_t tiles = (B + cand[i] - 1) / cand[i];
    if ((fits2 || (i == 2 && fits1)) && (tiles >= 2ll * sm_count || i == 2)) {
      S = cand[i];
      break;
    }
  }
  return S;
}

static int launch_pm(PmParams& p, void* stream) {
  constexpr bool bwd = MVAE_PM_BWD != 0;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  const int S = pick_tile(p.desc, p.B, bwd, di.sm_count, di.max_smem_optin, &p);
  if (S == 0) return MVAE_ERR_UNSUPPORTED;
  p.S = S;
  p.fd_ml = make_fastdiv(p.desc.ld_ml);
  p.fd_eps = make_fastdiv(p.desc.ld_eps);
  p.fd_z = make_fastdiv(p.desc.ld_z);
  p.fd_c = make_fastdiv(p.desc.C);
  p.fd_S = make_fastdiv(S);
  const size_t smem = 4 * pm_smem_floats(p.desc, S, bwd, p.ldp_ml, p.ldp_eps, p.ldp_z, p.ldp_c);
  const int64_t tiles = (p.B + S - 1) / S;
  if (tiles > 0x7fffffff) return MVAE_ERR_UNSUPPORTED;
  int maxn = 0;
  bool dyn = false;
  for (int i = 0; i < p.desc.C; ++i) {
    const int n = p.desc.comp[i].n;
    dyn = dyn || !(n >= 1 && n <= 8 && n != 7);
    maxn = n > maxn ? n : maxn;
  }
#if MVAE_PM_BWD
#define MVAE_PM_KERNEL pm_backward_kernel
#else
#define MVAE_PM_KERNEL pm_forward_kernel
#endif
  void (*kern)(PmParams) = dyn ? MVAE_PM_KERNEL<0> : maxn <= 2 ? MVAE_PM_KERNEL<2> : maxn <= 4 ? MVAE_PM_KERNEL<4>
                                                                                             : MVAE_PM_KERNEL<8>;
  if (smem > 48 * 1024) MVAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)tiles, kPmThreads, smem, as_stream(stream)>>>(p);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

}  // namespace mvae
