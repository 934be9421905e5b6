// pm_emul.cpp — TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
// Compiles the per-sample arithmetic of the fused product-manifold kernels (mvae_b200/csrc/pm_math.cuh) as plain
// C++ (libm stand-ins for the MUFU ops) so that tests/test_pm_math_host.py can check the algebra of the closed forms
// and of the hand-derived reverse sweep against the oracle on a machine without a GPU.
#include "../../include/mvae_b200.h"
#include "../../mvae_b200/csrc/pm_math.cuh"

using namespace mvae::pm;

template <int N, bool BWD>
static void run(const mvae_component& c, float rp, const float* ml, const float* eps, float* z, float* kl, float* mu,
                float* sigma, const float* gz, float gkl, float* gml, float* gR) {
  CompOut<N> o;
  const int n = N > 0 ? N : c.n;
  const CompConst K = make_const(rp);
  float g = 0.f;
  const float* m = ml + c.m_off;
  const float* l = ml + c.l_off;
  const float* e = eps + c.eps_off;
  const float* gzc = BWD ? gz + c.z_off : nullptr;
  float* gm = BWD ? gml + c.m_off : nullptr;
  float* gl = BWD ? gml + c.l_off : nullptr;
  switch (c.type) {
    case MVAE_EUCLIDEAN: comp_e<N, BWD>(n, c.l_n, m, l, e, o, gzc, gkl, gm, gl); break;
    case MVAE_HYPERBOLOID: comp_hsp<N, BWD, kHyp, true>(n, c.l_n, m, l, e, K, o, gzc, gkl, gm, gl, &g); break;
    case MVAE_SPHERE: comp_hsp<N, BWD, kSph, true>(n, c.l_n, m, l, e, K, o, gzc, gkl, gm, gl, &g); break;
    case MVAE_POINCARE: comp_hsp<N, BWD, kPoi, true>(n, c.l_n, m, l, e, K, o, gzc, gkl, gm, gl, &g); break;
    default: comp_hsp<N, BWD, kPsp, true>(n, c.l_n, m, l, e, K, o, gzc, gkl, gm, gl, &g); break;
  }
  if (BWD) {
    *gR += g * radius_d(rp);
    return;
  }
  for (int k = 0; k < c.d; ++k) {
    z[c.z_off + k] = o.z[k];
    if (mu) mu[c.z_off + k] = o.mu[k];
  }
  if (sigma)
    for (int j = 0; j < n; ++j) sigma[c.eps_off + j] = o.sigma[j];
  *kl = o.kl;
}

template <bool BWD>
static void dispatch(const mvae_component& c, float rp, const float* ml, const float* eps, float* z, float* kl,
                     float* mu, float* sigma, const float* gz, float gkl, float* gml, float* gR, int force_dyn) {
  if (force_dyn) return run<0, BWD>(c, rp, ml, eps, z, kl, mu, sigma, gz, gkl, gml, gR);
  switch (c.n) {
    case 1: return run<1, BWD>(c, rp, ml, eps, z, kl, mu, sigma, gz, gkl, gml, gR);
    case 2: return run<2, BWD>(c, rp, ml, eps, z, kl, mu, sigma, gz, gkl, gml, gR);
    case 3: return run<3, BWD>(c, rp, ml, eps, z, kl, mu, sigma, gz, gkl, gml, gR);
    case 4: return run<4, BWD>(c, rp, ml, eps, z, kl, mu, sigma, gz, gkl, gml, gR);
    case 5: return run<5, BWD>(c, rp, ml, eps, z, kl, mu, sigma, gz, gkl, gml, gR);
    case 6: return run<6, BWD>(c, rp, ml, eps, z, kl, mu, sigma, gz, gkl, gml, gR);
    case 8: return run<8, BWD>(c, rp, ml, eps, z, kl, mu, sigma, gz, gkl, gml, gR);
    default: return run<0, BWD>(c, rp, ml, eps, z, kl, mu, sigma, gz, gkl, gml, gR);
  }
}

extern "C" void pm_emul_forward(const mvae_pm_desc* D, int64_t B, const float* ml, const float* eps,
                                const float* radius, float* z, float* kl, float* mu, float* sigma, int force_dyn) {
  for (int64_t b = 0; b < B; ++b)
    for (int ci = 0; ci < D->C; ++ci) {
      const mvae_component& c = D->comp[ci];
      const float rp = (radius && c.type != MVAE_EUCLIDEAN) ? radius[ci] : 1.f;
      dispatch<false>(c, rp, ml + b * D->ld_ml, eps + b * D->ld_eps, z + b * D->ld_z, kl + b * D->C + ci,
                      mu ? mu + b * D->ld_z : nullptr, sigma ? sigma + b * D->ld_eps : nullptr, nullptr, 0.f, nullptr,
                      nullptr, force_dyn);
    }
}

extern "C" void pm_emul_backward(const mvae_pm_desc* D, int64_t B, const float* ml, const float* eps,
                                 const float* radius, const float* gz, const float* gkl, float gkl_scalar, float* gml,
                                 double* gradius, int force_dyn) {
  for (int64_t b = 0; b < B; ++b)
    for (int ci = 0; ci < D->C; ++ci) {
      const mvae_component& c = D->comp[ci];
      const float rp = (radius && c.type != MVAE_EUCLIDEAN) ? radius[ci] : 1.f;
      float g = 0.f;
      dispatch<true>(c, rp, ml + b * D->ld_ml, eps + b * D->ld_eps, nullptr, nullptr, nullptr, nullptr,
                     gz + b * D->ld_z, gkl ? gkl[b * D->C + ci] : gkl_scalar, gml + b * D->ld_ml, &g, force_dyn);
      gradius[ci] += g;
    }
}

// elementary functions, for accuracy sweeps against libm
extern "C" void pm_emul_sincos(int64_t n, const float* x, float* s, float* c) {
  for (int64_t i = 0; i < n; ++i) sincos_cw(x[i], s + i, c + i);
}
extern "C" void pm_emul_atan2(int64_t n, const float* y, const float* x, float* r) {
  for (int64_t i = 0; i < n; ++i) r[i] = atan2_pos(y[i], x[i]);
}
