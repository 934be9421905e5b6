// pm_kernels.cu — C-ABI entry points of the fused product-manifold kernels (mvae_pm_desc_init / mvae_pm_forward /
// mvae_pm_backward, include/mvae_b200.h): argument validation and descriptor packing.  The kernels live in
// pm_kernels_impl.cuh and are instantiated in pm_fwd.cu / pm_bwd.cu.
#include <string.h>

#include "pm_math.cuh"
#include "pm_params.cuh"

namespace mvae {

static int validate_desc(const mvae_pm_desc* d) {
  if (!d || d->C < 1 || d->C > MVAE_MAX_COMPONENTS) return MVAE_ERR_INVALID_ARGUMENT;
  if (d->ld_ml < 1 || d->ld_eps < 1 || d->ld_z < 1) return MVAE_ERR_INVALID_ARGUMENT;
  if (d->ld_ml > 16384 || d->ld_z > 16384) return MVAE_ERR_UNSUPPORTED;
  for (int i = 0; i < d->C; ++i) {
    const mvae_component& c = d->comp[i];
    if (c.type < MVAE_EUCLIDEAN || c.type > MVAE_UNIVERSAL) return MVAE_ERR_INVALID_ARGUMENT;
    if (c.n < 1) return MVAE_ERR_INVALID_ARGUMENT;
    if (c.n > pm::kDynMaxN) return MVAE_ERR_UNSUPPORTED;
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

}  // namespace mvae

using namespace mvae;

extern "C" int mvae_pm_desc_init(mvae_pm_desc* D, int32_t C, const int32_t* types, const int32_t* dims,
                                 int32_t scalar) {
  if (!D || !types || !dims || C < 1 || C > MVAE_MAX_COMPONENTS) return MVAE_ERR_INVALID_ARGUMENT;
  memset(D, 0, sizeof(*D));
  int ml = 0, e = 0, z = 0;
  for (int i = 0; i < C; ++i) {
    mvae_component* c = &D->comp[i];
    if (dims[i] < 1 || types[i] < MVAE_EUCLIDEAN || types[i] > MVAE_UNIVERSAL) return MVAE_ERR_INVALID_ARGUMENT;
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
  return launch_pm_forward(p, stream);
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
  return launch_pm_backward(p, stream);
}
