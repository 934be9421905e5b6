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
#include <math.h>
#include <string.h>

#include "manifold_math.cuh"

namespace mvae {

constexpr int kPmThreads = 128;

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
  // tiling
  int S;  // samples per CTA tile (multiple of 32)
  int ldp_ml, ldp_eps, ldp_z, ldp_c;
  FastDiv fd_ml, fd_eps, fd_z, fd_c, fd_S;
  int vec_ok;  // all global pointers 16-byte aligned
};

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

// rows x ld floats, contiguous in global at g  ->  shared tile with row stride ldp
__device__ __forceinline__ void tile_load(float* __restrict__ s, const float* __restrict__ g, int rows, int ld, int ldp,
                                          FastDiv fd, bool vec) {
  const int total = rows * ld;
  int done = 0;
  if (vec) {
    const int nvec = total >> 2;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      float4 v = ldg_stream4(g + 4 * i);
      uint32_t idx = 4u * i;
      uint32_t r = fastdiv(idx, fd);
      uint32_t c = idx - r * ld;
      float* row = s + r * ldp;
      float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        row[c] = vals[k];
        if (++c == (uint32_t)ld) {
          c = 0;
          row += ldp;
        }
      }
    }
    done = nvec << 2;
  }
  for (int i = done + threadIdx.x; i < total; i += blockDim.x) {
    uint32_t r = fastdiv((uint32_t)i, fd);
    uint32_t c = i - r * ld;
    s[r * ldp + c] = __ldg(g + i);
  }
}

__device__ __forceinline__ void tile_store(float* __restrict__ g, const float* __restrict__ s, int rows, int ld,
                                           int ldp, FastDiv fd, bool vec) {
  const int total = rows * ld;
  int done = 0;
  if (vec) {
    const int nvec = total >> 2;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      uint32_t idx = 4u * i;
      uint32_t r = fastdiv(idx, fd);
      uint32_t c = idx - r * ld;
      const float* row = s + r * ldp;
      float vals[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        vals[k] = row[c];
        if (++c == (uint32_t)ld) {
          c = 0;
          row += ldp;
        }
      }
      stg_stream4(g + 4 * i, make_float4(vals[0], vals[1], vals[2], vals[3]));
    }
    done = nvec << 2;
  }
  for (int i = done + threadIdx.x; i < total; i += blockDim.x) {
    uint32_t r = fastdiv((uint32_t)i, fd);
    uint32_t c = i - r * ld;
    g[i] = s[r * ldp + c];
  }
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

template <bool BWD, bool DYN>
__device__ __forceinline__ bool dispatch_item(const mvae_component& c, float R, float rp, const float* ml_row,
                                              const float* eps_row, float* z_row, float* kl_slot, float* mu_row,
                                              float* sigma_row, const float* gz_row, float gkl, float* gml_row,
                                              float* gR_acc) {
#define MVAE_CASE(NN)                                                                                            \
  case NN:                                                                                                       \
    return run_item<NN, BWD>(c, R, rp, ml_row, eps_row, z_row, kl_slot, mu_row, sigma_row, gz_row, gkl, gml_row, \
                             gR_acc);
  switch (c.n) {
    MVAE_CASE(1)
    MVAE_CASE(2)
    MVAE_CASE(3)
    MVAE_CASE(4)
    MVAE_CASE(5)
    MVAE_CASE(6)
    MVAE_CASE(8)
    default:
      if constexpr (DYN)
        return run_item<0, BWD>(c, R, rp, ml_row, eps_row, z_row, kl_slot, mu_row, sigma_row, gz_row, gkl, gml_row,
                                gR_acc);
      else
        return true;  // unreachable: the host picks the DYN kernel when any dimension is outside the static set
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
  const int S = p.S;
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

template <bool DYN>
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
  tile_load(s.ml, p.ml + row0 * p.desc.ld_ml, rows, p.desc.ld_ml, p.ldp_ml, p.fd_ml, vec);
  tile_load(s.eps, p.eps + row0 * p.desc.ld_eps, rows, p.desc.ld_eps, p.ldp_eps, p.fd_eps, vec);
  __syncthreads();
  const bool want_mu = p.mu != nullptr, want_sigma = p.sigma != nullptr;
  bool finite = true;
  const int items = C * p.S;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int ci = (int)fastdiv((uint32_t)it, p.fd_S);
    const int sidx = it - ci * p.S;
    if (sidx >= rows) continue;
    const mvae_component c = s.comp[ci];
    finite &= dispatch_item<false, DYN>(c, s.R[ci], s.rp[ci], s.ml + sidx * p.ldp_ml, s.eps + sidx * p.ldp_eps,
                                   s.a + sidx * p.ldp_z, s.b + sidx * p.ldp_c + ci,
                                   want_mu ? s.c + sidx * p.ldp_z : nullptr,
                                   want_sigma ? s.d + sidx * p.ldp_eps : nullptr, nullptr, 0.f, nullptr, nullptr);
  }
  if (p.flag) {
    unsigned bad = __ballot_sync(0xffffffffu, !finite);
    if (bad && (threadIdx.x & 31) == 0) atomicOr(p.flag, 1u);
  }
  __syncthreads();
  tile_store(p.z + row0 * p.desc.ld_z, s.a, rows, p.desc.ld_z, p.ldp_z, p.fd_z, vec);
  tile_store(p.kl + row0 * C, s.b, rows, C, p.ldp_c, p.fd_c, vec);
  if (want_mu) tile_store(p.mu + row0 * p.desc.ld_z, s.c, rows, p.desc.ld_z, p.ldp_z, p.fd_z, vec);
  if (want_sigma) tile_store(p.sigma + row0 * p.desc.ld_eps, s.d, rows, p.desc.ld_eps, p.ldp_eps, p.fd_eps, vec);
}

template <bool DYN>
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
  tile_load(s.ml, p.ml + row0 * p.desc.ld_ml, rows, p.desc.ld_ml, p.ldp_ml, p.fd_ml, vec);
  tile_load(s.eps, p.eps + row0 * p.desc.ld_eps, rows, p.desc.ld_eps, p.ldp_eps, p.fd_eps, vec);
  tile_load(s.a, p.gz + row0 * p.desc.ld_z, rows, p.desc.ld_z, p.ldp_z, p.fd_z, vec);
  if (p.gkl) tile_load(s.b, p.gkl + row0 * C, rows, C, p.ldp_c, p.fd_c, vec);
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
      dispatch_item<true, DYN>(c, s.R[ci], s.rp[ci], s.ml + sidx * p.ldp_ml, s.eps + sidx * p.ldp_eps, nullptr, nullptr,
                          nullptr, nullptr, s.a + sidx * p.ldp_z, gkl, s.c + sidx * p.ldp_ml, &gR);
    }
    gR = warp_sum(gR);
    if ((threadIdx.x & 31) == 0 && gR != 0.f) atomicAdd(&s.gR[ci], gR);
  }
  __syncthreads();
  tile_store(p.gml + row0 * p.desc.ld_ml, s.c, rows, p.desc.ld_ml, p.ldp_ml, p.fd_ml, vec);
  if (p.gradius)
    for (int i = threadIdx.x; i < C; i += blockDim.x)
      if (s.gR[i] != 0.f) atomicAdd(p.gradius + i, s.gR[i]);
}

static int validate_desc(const mvae_pm_desc* d) {
  if (!d || d->C < 1 || d->C > MVAE_MAX_COMPONENTS) return MVAE_ERR_INVALID_ARGUMENT;
  if (d->ld_ml < 1 || d->ld_eps < 1 || d->ld_z < 1) return MVAE_ERR_INVALID_ARGUMENT;
  if (d->ld_ml > 16384 || d->ld_z > 16384) return MVAE_ERR_UNSUPPORTED;
  for (int i = 0; i < d->C; ++i) {
    const mvae_component& c = d->comp[i];
    if (c.type < MVAE_EUCLIDEAN || c.type > MVAE_PROJ_SPHERE) return MVAE_ERR_INVALID_ARGUMENT;
    if (c.type == MVAE_PROJ_SPHERE) return MVAE_ERR_UNSUPPORTED;
    if (c.n < 1) return MVAE_ERR_INVALID_ARGUMENT;
    if (c.n > kDynMaxN) return MVAE_ERR_UNSUPPORTED;
    const int d_expect = (c.type == MVAE_HYPERBOLOID || c.type == MVAE_SPHERE) ? c.n + 1 : c.n;
    if (c.d != d_expect) return MVAE_ERR_INVALID_ARGUMENT;
    if (c.l_n != c.n && c.l_n != 1) return MVAE_ERR_INVALID_ARGUMENT;
    if (c.m_off < 0 || c.m_off + c.n > d->ld_ml || c.l_off < 0 || c.l_off + c.l_n > d->ld_ml)
      return MVAE_ERR_INVALID_ARGUMENT;
    if (c.eps_off < 0 || c.eps_off + c.n > d->ld_eps) return MVAE_ERR_INVALID_ARGUMENT;
    if (c.z_off < 0 || c.z_off + c.d > d->ld_z) return MVAE_ERR_INVALID_ARGUMENT;
  }
  return MVAE_OK;
}

static bool aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Choose the tile height: the largest S in {128, 64, 32} that keeps >= 2 CTAs per SM worth of shared memory and
// still yields at least ~2 waves of CTAs for small batches.
static int pick_tile(const mvae_pm_desc& D, int64_t B, bool bwd, int sm_count, int max_smem, PmParams* p) {
  p->ldp_ml = D.ld_ml | 1;
  p->ldp_eps = D.ld_eps | 1;
  p->ldp_z = D.ld_z | 1;
  p->ldp_c = D.C | 1;
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

static int launch_pm(bool bwd, PmParams& p, void* stream) {
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
  bool dyn = false;
  for (int i = 0; i < p.desc.C; ++i) {
    const int n = p.desc.comp[i].n;
    dyn = dyn || !(n >= 1 && n <= 8 && n != 7);
  }
  void (*kern)(PmParams) = bwd ? (dyn ? pm_backward_kernel<true> : pm_backward_kernel<false>)
                               : (dyn ? pm_forward_kernel<true> : pm_forward_kernel<false>);
  if (smem > 48 * 1024) MVAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)tiles, kPmThreads, smem, as_stream(stream)>>>(p);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

}  // namespace mvae

using namespace mvae;

extern "C" int mvae_pm_desc_init(mvae_pm_desc* D, int32_t C, const int32_t* types, const int32_t* dims,
                                 int32_t scalar) {
  if (!D || !types || !dims || C < 1 || C > MVAE_MAX_COMPONENTS) return MVAE_ERR_INVALID_ARGUMENT;
  memset(D, 0, sizeof(*D));
  int ml = 0, e = 0, z = 0;
  for (int i = 0; i < C; ++i) {
    mvae_component* c = &D->comp[i];
    if (dims[i] < 1 || types[i] < MVAE_EUCLIDEAN || types[i] > MVAE_PROJ_SPHERE) return MVAE_ERR_INVALID_ARGUMENT;
    c->type = types[i];
    c->n = dims[i];
    c->d = (types[i] == MVAE_HYPERBOLOID || types[i] == MVAE_SPHERE) ? dims[i] + 1 : dims[i];
    c->l_n = scalar ? 1 : dims[i];
    c->m_off = ml;
    ml += c->n;
    c->l_off = ml;
    ml += c->l_n;
    c->eps_off = e;
    e += c->n;
    c->z_off = z;
    z += c->d;
  }
  D->C = C;
  D->ld_ml = ml;
  D->ld_eps = e;
  D->ld_z = z;
  return MVAE_OK;
}

extern "C" int mvae_pm_forward(const mvae_pm_desc* desc, int64_t B, const float* ml, const float* eps,
                               const float* radius, float* z, float* kl, float* mu, float* sigma,
                               uint32_t* nonfinite_flag, void* stream) {
  int rc = validate_desc(desc);
  if (rc != MVAE_OK) return rc;
  if (B < 0 || (B > 0 && (!ml || !eps || !z || !kl))) return MVAE_ERR_INVALID_ARGUMENT;
  if (B == 0) return MVAE_OK;
  PmParams p;
  memset(&p, 0, sizeof(p));
  p.desc = *desc;
  p.B = B;
  p.ml = ml;
  p.eps = eps;
  p.radius = radius;
  p.z = z;
  p.kl = kl;
  p.mu = mu;
  p.sigma = sigma;
  p.flag = nonfinite_flag;
  p.vec_ok = aligned16(ml) && aligned16(eps) && aligned16(z) && aligned16(kl) && aligned16(mu) && aligned16(sigma);
  return launch_pm(false, p, stream);
}

extern "C" int mvae_pm_backward(const mvae_pm_desc* desc, int64_t B, const float* ml, const float* eps,
                                const float* radius, const float* gz, const float* gkl, float gkl_scalar, float* gml,
                                float* gradius, void* stream) {
  int rc = validate_desc(desc);
  if (rc != MVAE_OK) return rc;
  if (B < 0 || (B > 0 && (!ml || !eps || !gz || !gml))) return MVAE_ERR_INVALID_ARGUMENT;
  if (B == 0) return MVAE_OK;
  PmParams p;
  memset(&p, 0, sizeof(p));
  p.desc = *desc;
  p.B = B;
  p.ml = ml;
  p.eps = eps;
  p.radius = radius;
  p.gz = gz;
  p.gkl = gkl;
  p.gkl_scalar = gkl_scalar;
  p.gml = gml;
  p.gradius = gradius;
  p.vec_ok = aligned16(ml) && aligned16(eps) && aligned16(gz) && aligned16(gkl) && aligned16(gml);
  return launch_pm(true, p, stream);
}
