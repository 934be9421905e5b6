// manifold_math.cuh — per-sample device arithmetic of the mixed-curvature latent components (fp32).
//
// One `comp_*<N, BWD>` function per manifold runs, for ONE sample of ONE component, the whole chain
//   Component.encode -> reparametrize -> q_z.rsample_with_parts -> kl_loss (log q - log p)
// of the reference (paths relative to the reference root):
//   mt/mvae/components/component.py:63-75, mt/mvae/sampling/sampling_procedures.py:91-116,145-155,
//   mt/mvae/distributions/wrapped_normal.py:70-103, mt/mvae/ops/{hyperbolics,spherical,euclidean,poincare}.py,
//   guarded scalar math mt/mvae/ops/common.py:28-147 (LeakyClamp / Atanh / Acosh custom backward rules),
//   geoopt==0.1.0 poincare math for the Poincare ball (un-vendored third party; constants as in DESIGN.md).
// With BWD the function also runs the reverse sweep of that chain by recomputation (hand-derived; the
// autograd conventions of the reference — leaky clamps pass 1e-8*g outside, plain clamps pass 0,
// Acosh.backward = g / sqrt(x'^2-1) — are reproduced, not "fixed").
//
// N > 0: true dimension known at compile time (everything lives in registers); N == 0: runtime n <= kDynMaxN
// (arrays spill to local memory — slow path for unusually wide components).
#pragma once
#include "mvae_common.cuh"
#pragma nv_diag_suppress 128  // "loop is not reachable" in the forward-only instantiations

namespace mvae {

constexpr int kDynMaxN = 160;  // widest single component served (tangent dimension)

template <int N>
struct Cap {
  static constexpr int n = N > 0 ? N : kDynMaxN;
  static constexpr int d = n + 1;
};

#define MVAE_DEV __device__ __forceinline__
// Loops over coordinates are fully unrolled when the dimension is static and left rolled for the dynamic path;
// every function using MVAE_UNROLL defines `constexpr int UN`.
#define MVAE_UNROLL _Pragma("unroll UN")
#define MVAE_UN(N) constexpr int UN = (N) > 0 ? (N) + 1 : 1

constexpr float kHalfLn2Pi = 0.9189385332046727f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kMaxHyp = 85.f;  // common.py:107-114 clamp of cosh/sinh arguments

// ---- guarded scalar math, ops/common.py --------------------------------------------------------------------
MVAE_DEV float lclamp(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
MVAE_DEV float lclamp_d(float x, float lo, float hi) { return (x >= lo && x <= hi) ? 1.f : 1e-8f; }
MVAE_DEV float lclamp_lo(float x, float lo) { return x < lo ? lo : x; }
MVAE_DEV float lclamp_lo_d(float x, float lo) { return x >= lo ? 1.f : 1e-8f; }
// sqrt (common.py:117-119)
MVAE_DEV float sqrt_g(float x) { return sqrtf(lclamp_lo(x, 1e-9f)); }
MVAE_DEV float sqrt_g_d(float x, float y) { return lclamp_lo_d(x, 1e-9f) * 0.5f / y; }
// cosh & sinh of a clamped argument in one go: e = exp(|x|)
MVAE_DEV void coshsinh_g(float x, float* ch, float* sh) {
  float xc = lclamp(x, -kMaxHyp, kMaxHyp);
  *ch = coshf(xc);
  *sh = sinhf(xc);
}
// acosh (common.py:76-94)
MVAE_DEV float acosh_g(float x, float* z_out) {
  float xc = lclamp_lo(x, 1.0f + 1e-8f);
  float z = sqrt_g(xc * xc - 1.f);
  *z_out = z;
  return logf(xc + z);
}
// logsinh (common.py:122-128, logsumexp_signs :139-147); *d = dy/dx along the autograd graph
template <bool BWD>
MVAE_DEV float logsinh_g(float x, float* d) {
  float v1 = -2.f * x;
  bool arg1 = v1 > 0.f;
  float M = arg1 ? v1 : 0.f;
  float e0 = expf(0.f - M), e1 = expf(v1 - M);
  float s = e0 - e1;
  float sc = lclamp_lo(s, 1e-8f);
  float y = x + (M + logf(sc)) - kLn2;
  if (BWD) {
    float dM = arg1 ? -2.f : 0.f;
    float ds = -e0 * dM - e1 * (-2.f - dM);
    *d = 1.f + dM + lclamp_lo_d(s, 1e-8f) / sc * ds;
  }
  return y;
}
// F.softplus (beta 1, threshold 20)
MVAE_DEV float softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }
MVAE_DEV float softplus_d(float x) {
  if (x > 20.f) return 1.f;
  float z = expf(x);
  return z / (z + 1.f);
}
// radius = clamp(relu(R_param), 1e-8, 1e8) (manifold.py:73-75), plain clamp
MVAE_DEV float radius_of(float rp) {
  float r = rp > 0.f ? rp : 0.f;
  return r < 1e-8f ? 1e-8f : (r > 1e8f ? 1e8f : r);
}
MVAE_DEV float radius_d(float rp) {
  if (!(rp > 0.f)) return 0.f;
  return (rp >= 1e-8f && rp <= 1e8f) ? 1.f : 0.f;
}

// Per-sample result of one component.
template <int N>
struct CompOut {
  float mu[Cap<N>::d];
  float sigma[Cap<N>::n];
  float z[Cap<N>::d];
  float u[Cap<N>::d];
  float kl, logq, logp;
};

// sigma_j = softplus(l_j) + 1e-5 (component.py:69-72; scalar parametrization repeats one value, wrapped_normal.py:46-49)
template <int N>
MVAE_DEV void load_sigma(int n, int l_n, const float* l, float* sg) {
  MVAE_UN(N);
  if (l_n == 1) {
    float s = softplus(l[0]) + 1e-5f;
    MVAE_UNROLL
    for (int j = 0; j < Cap<N>::n; ++j)
      if (j < n) sg[j] = s;
  } else {
    MVAE_UNROLL
    for (int j = 0; j < Cap<N>::n; ++j)
      if (j < n) sg[j] = softplus(l[j]) + 1e-5f;
  }
}

// d(loss)/d l from d(loss)/d sigma
template <int N>
MVAE_DEV void store_gl(int n, int l_n, const float* l, const float* g_s, float* gl) {
  MVAE_UN(N);
  if (l_n == 1) {
    float acc = 0.f;
    MVAE_UNROLL
    for (int j = 0; j < Cap<N>::n; ++j)
      if (j < n) acc += g_s[j];
    gl[0] = acc * softplus_d(l[0]);
  } else {
    MVAE_UNROLL
    for (int j = 0; j < Cap<N>::n; ++j)
      if (j < n) gl[j] = g_s[j] * softplus_d(l[j]);
  }
}

// ================================================ EUCLIDEAN ================================================
// euclidean.py:78-79 (mu = m/2); EuclideanNormalProcedure (sampling_procedures.py:145-155):
// z = mu + eps*sigma (wrapped_distributions.py:25-27), KL(N(mu,sigma)||N(0,1)).sum(-1).
template <int N, bool BWD>
MVAE_DEV void comp_e(int n, int l_n, const float* m, const float* l, const float* e, CompOut<N>& o, const float* gz,
                     float gkl, float* gm, float* gl) {
  MVAE_UN(N);
  constexpr int CN = Cap<N>::n;
  load_sigma<N>(n, l_n, l, o.sigma);
  float kl = 0.f, lq = 0.f, lpz = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      float mu = m[j] / 2.f;
      float s = o.sigma[j];
      float z = mu + e[j] * s;
      o.mu[j] = mu;
      o.z[j] = z;
      o.u[j] = 0.f;
      float var_ratio = s * s;
      kl += 0.5f * (var_ratio + mu * mu - 1.f - logf(var_ratio));
      lq += -((z - mu) * (z - mu)) / (2.f * s * s) - logf(s) - kHalfLn2Pi;
      lpz += -(z * z) / 2.f - kHalfLn2Pi;
    }
  o.kl = kl;
  o.logq = lq;
  o.logp = lpz;
  if (!BWD) return;
  float g_s[CN];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      float mu = o.mu[j], s = o.sigma[j];
      gm[j] = (gz[j] + gkl * mu) / 2.f;
      g_s[j] = gz[j] * e[j] + gkl * (s - 1.f / s);
    }
  store_gl<N>(n, l_n, l, g_s, gl);
}

// =============================================== HYPERBOLOID ===============================================
// H._logdet (hyperbolics.py:58-65) on the Lorentz squared norm `pr` of u:
// (n-1)(log R + logsinh(r) - log r), r = sqrt_g(pr)/R.  With BWD: *g_pr = g * d(ld)/d(pr), *gR += g * d(ld)/dR.
template <bool BWD>
MVAE_DEV float h_logdet_pr(int n, float pr, float R, float g, float* g_pr, float* gR) {
  float s = sqrt_g(pr);
  float r = s / R;
  float dls = 0.f;
  float ld = (float)(n - 1) * (logf(R) + logsinh_g<BWD>(r, &dls) - logf(r));
  if (BWD) {
    float g_r = g * (float)(n - 1) * (dls - 1.f / r);
    *gR += g * (float)(n - 1) / R - g_r * r / R;
    *g_pr = (g_r / R) * sqrt_g_d(pr, s);
  }
  return ld;
}

// Forward: component.py:63-75, hyperbolics.py:114-121 (exp_map_mu0), :87-93 (PT mu0->mu), :106-111 (exp_map),
// wrapped_normal.py:84-103, hyperbolics.py:58-65 (logdet), :124-128 (inverse_exp_map), :96-103,145-148 (inverse PT).
template <int N, bool BWD>
MVAE_DEV void comp_h(int n, int l_n, const float* m, const float* l, const float* e, float R, CompOut<N>& o,
                     const float* gz, float gkl, float* gm, float* gl, float* gR_out) {
  MVAE_UN(N);
  constexpr int CN = Cap<N>::n;
  constexpr int CD = Cap<N>::d;
  const int d = n + 1;
  // ---- encode ----
  float nm2 = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) nm2 += m[j] * m[j];
  float nm = sqrtf(nm2);
  float a = nm / R;
  float dn = fmaxf(nm, 1e-12f);
  float ch, sh;
  coshsinh_g(a, &ch, &sh);
  float* mu = o.mu;
  float xn[CN];
  mu[0] = ch * R;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      xn[j] = (m[j] / dn) * R;
      mu[j + 1] = sh * xn[j];
    }
  float* sg = o.sigma;
  load_sigma<N>(n, l_n, l, sg);
  // ---- sample: v = eps*sigma; u = PT_{mu0->mu}([0,v]); z = exp_mu(u) ----
  float v[CN];
  float lp = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      v[j] = e[j] * sg[j];
      lp += mu[j + 1] * v[j];
    }
  float denom = R * (R + mu[0]);
  float coef = lp / denom;
  float* u = o.u;
  u[0] = coef * (mu[0] + R);
  float pr = u[0] * u[0];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      u[j + 1] = v[j] + coef * mu[j + 1];
      pr += u[j + 1] * u[j + 1];
    }
  pr = pr - 2.f * (u[0] * u[0]);
  float ln = sqrt_g(pr);
  float t = ln / R;
  float cht, sht;
  coshsinh_g(t, &cht, &sht);
  float* z = o.z;
  float un[CD];
  MVAE_UNROLL
  for (int k = 0; k < CD; ++k)
    if (k < d) {
      un[k] = u[k] / t;
      z[k] = cht * mu[k] + sht * un[k];
    }
  // ---- log q ----
  float nlp = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) nlp += -(v[j] * v[j]) / (2.f * (sg[j] * sg[j])) - logf(sg[j]) - kHalfLn2Pi;
  float ld = h_logdet_pr<false>(n, pr, R, 0.f, nullptr, nullptr);
  o.logq = nlp - ld;
  // ---- log p: at_point mu0 = [R,0..]; alpha = -<mu0,z>_L / R^2 ----
  float lpz = R * z[0] - 2.f * (R * z[0]);
  float alpha = -lpz / (R * R);
  float zz;
  float ach = acosh_g(alpha, &zz);
  float sq = sqrt_g(alpha * alpha - 1.f);
  float coefp = ach / sq;
  float w[CD];
  w[0] = coefp * (z[0] - alpha * R);
  float pr0 = w[0] * w[0];
  float nlp0 = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      w[j + 1] = coefp * z[j + 1];
      pr0 += w[j + 1] * w[j + 1];
      nlp0 += -(w[j + 1] * w[j + 1]) / 2.f - kHalfLn2Pi;  // v0 = w[1:] (inverse PT to mu0 leaves the tail unchanged)
    }
  pr0 = pr0 - 2.f * (w[0] * w[0]);
  float ld0 = h_logdet_pr<false>(n, pr0, R, 0.f, nullptr, nullptr);
  o.logp = nlp0 - ld0;
  o.kl = o.logq - o.logp;
  if (!BWD) return;

  // ================================ reverse sweep ================================
  float gR = 0.f;
  const float g_logq = gkl, g_logp = -gkl;
  // logp = nlp0 - ld0
  float g_pr0;
  (void)h_logdet_pr<true>(n, pr0, R, -g_logp, &g_pr0, &gR);
  float gw[CD];
  gw[0] = g_pr0 * (-2.f * w[0]);
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) gw[j + 1] = g_pr0 * 2.f * w[j + 1] + g_logp * (-w[j + 1]);
  // inverse PT: c2 = -w0/(R+R) multiplies [2R, 0...]; only its (discarded) 0-th output depends on it -> no gradient.
  // w = coefp (z - alpha mu0)
  float gzt[CD];
  float g_coefp = gw[0] * (z[0] - alpha * R);
  float g_alpha = -gw[0] * coefp * R;
  float g_mu0p0 = -gw[0] * coefp * alpha;
  gzt[0] = gz[0] + gw[0] * coefp;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      g_coefp += gw[j + 1] * z[j + 1];
      gzt[j + 1] = gz[j + 1] + gw[j + 1] * coefp;
    }
  float g_ach = g_coefp / sq;
  float g_sq = -g_coefp * ach / (sq * sq);
  g_alpha += g_sq * sqrt_g_d(alpha * alpha - 1.f, sq) * 2.f * alpha;
  g_alpha += g_ach / zz;
  // alpha = -lpz / R^2 ; lpz = mu0p0*z0 - 2 mu0p0*z0
  float g_lpz = -g_alpha / (R * R);
  gR += g_alpha * (-2.f * alpha / R);
  gzt[0] += -g_lpz * R;
  g_mu0p0 += -g_lpz * z[0];
  gR += g_mu0p0;  // mu_0 = e_0 * radius (hyperbolics.py:68-69)
  // logq = nlp - ld
  float g_pr;
  (void)h_logdet_pr<true>(n, pr, R, -g_logq, &g_pr, &gR);
  const float g_nlp = g_logq;
  // z = cht*mu + sht*un ; un = u/t
  float g_cht = 0.f, g_sht = 0.f, g_t = 0.f;
  float g_mu[CD], gu[CD];
  MVAE_UNROLL
  for (int k = 0; k < CD; ++k)
    if (k < d) {
      g_cht += gzt[k] * mu[k];
      g_sht += gzt[k] * un[k];
      g_mu[k] = gzt[k] * cht;
      float g_un = gzt[k] * sht;
      gu[k] = g_un / t;
      g_t += -g_un * u[k] / (t * t);
    }
  g_t += (g_cht * sht + g_sht * cht) * lclamp_d(t, -kMaxHyp, kMaxHyp);
  float g_ln = g_t / R;
  gR += -g_t * t / R;
  g_pr += g_ln * sqrt_g_d(pr, ln);
  gu[0] += g_pr * -2.f * u[0];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) gu[j + 1] += g_pr * 2.f * u[j + 1];
  // u0 = coef (mu0 + R); u_j = v_j + coef mu_j
  float g_coef = gu[0] * (mu[0] + R);
  g_mu[0] += gu[0] * coef;
  gR += gu[0] * coef;
  float g_v[CN];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      g_coef += gu[j + 1] * mu[j + 1];
      g_v[j] = gu[j + 1];
      g_mu[j + 1] += gu[j + 1] * coef;
    }
  float g_lp = g_coef / denom;
  float g_denom = -g_coef * coef / denom;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      g_mu[j + 1] += g_lp * v[j];
      g_v[j] += g_lp * mu[j + 1];
    }
  gR += g_denom * (2.f * R + mu[0]);
  g_mu[0] += g_denom * R;
  // nlp, v = e*sigma
  float g_s[CN];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      float s = sg[j];
      g_v[j] += g_nlp * (-v[j] / (s * s));
      g_s[j] = g_nlp * (v[j] * v[j] / (s * s * s) - 1.f / s) + g_v[j] * e[j];
    }
  store_gl<N>(n, l_n, l, g_s, gl);
  // mu0 = ch R ; mu_j = sh xn_j ; xn = m/dn*R
  float g_ch = g_mu[0] * R;
  gR += g_mu[0] * ch;
  float g_sh = 0.f, g_dn = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      g_sh += g_mu[j + 1] * xn[j];
      float g_xn = g_mu[j + 1] * sh;
      gR += g_xn * m[j] / dn;
      gm[j] = g_xn * R / dn;
      g_dn += -g_xn * R * m[j] / (dn * dn);
    }
  float g_a = (g_ch * sh + g_sh * ch) * lclamp_d(a, -kMaxHyp, kMaxHyp);
  float g_nm = (nm >= 1e-12f) ? g_dn : 0.f;
  g_nm += g_a / R;
  gR += -g_a * a / R;
  if (nm > 0.f) {
    float k = g_nm / nm;
    MVAE_UNROLL
    for (int j = 0; j < CN; ++j)
      if (j < n) gm[j] += k * m[j];
  }
  *gR_out = gR;
}

// ================================================== SPHERE ==================================================
// S._logdet (spherical.py:58-67) on nu = ||u||_2: (n-1)(log R + log clamp(|sin r|, 1e-5) - log clamp(r, 1e-5)), plain clamps.
template <bool BWD>
MVAE_DEV float s_logdet_nu(int n, float nu, float R, float g, float* g_nu, float* gR) {
  float r = nu / R;
  float sn = sinf(r);
  float as = fabsf(sn);
  float asc = fmaxf(as, 1e-5f);
  float rc = fmaxf(r, 1e-5f);
  float ld = (float)(n - 1) * (logf(R) + logf(asc) - logf(rc));
  if (BWD) {
    float k1 = g * (float)(n - 1);
    *gR += k1 / R;
    float g_r = 0.f;
    if (as >= 1e-5f) g_r += k1 / asc * (sn > 0.f ? 1.f : (sn < 0.f ? -1.f : 0.f)) * cosf(r);
    if (r >= 1e-5f) g_r += -k1 / rc;
    *gR += -g_r * r / R;
    *g_nu = g_r / R;
  }
  return ld;
}

// spherical.py:94-101 (exp_map_mu0), :74-77 (PT), :86-91 (exp_map), :104-109 (inverse_exp_map), :80-83 (inverse PT).
template <int N, bool BWD>
MVAE_DEV void comp_s(int n, int l_n, const float* m, const float* l, const float* e, float R, CompOut<N>& o,
                     const float* gz, float gkl, float* gm, float* gl, float* gR_out) {
  MVAE_UN(N);
  constexpr int CN = Cap<N>::n;
  constexpr int CD = Cap<N>::d;
  const int d = n + 1;
  float nm2 = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) nm2 += m[j] * m[j];
  float nm = sqrtf(nm2);
  float a = nm / R;
  float dn = fmaxf(nm, 1e-12f);
  float ca = cosf(a), sa = sinf(a);
  float* mu = o.mu;
  float xn[CN];
  mu[0] = ca * R;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      xn[j] = (m[j] / dn) * R;
      mu[j + 1] = sa * xn[j];
    }
  float* sg = o.sigma;
  load_sigma<N>(n, l_n, l, sg);
  float v[CN];
  float dp = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      v[j] = e[j] * sg[j];
      dp += mu[j + 1] * v[j];
    }
  float denom = R * (R + mu[0]);
  float coef = dp / denom;
  float* u = o.u;
  u[0] = 0.f - coef * (mu[0] + R);
  float nu2 = u[0] * u[0];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      u[j + 1] = v[j] - coef * mu[j + 1];
      nu2 += u[j + 1] * u[j + 1];
    }
  float nu = sqrtf(nu2);
  float t = nu / R;
  float ct = cosf(t), st = sinf(t);
  float* z = o.z;
  float un[CD];
  MVAE_UNROLL
  for (int k = 0; k < CD; ++k)
    if (k < d) {
      un[k] = u[k] / t;
      z[k] = ct * mu[k] + st * un[k];
    }
  float nlp = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) nlp += -(v[j] * v[j]) / (2.f * (sg[j] * sg[j])) - logf(sg[j]) - kHalfLn2Pi;
  float ld = s_logdet_nu<false>(n, nu, R, 0.f, nullptr, nullptr);
  o.logq = nlp - ld;
  // prior: at_point = mu0 = [R, 0..]
  float alpha = (R * z[0]) / (R * R);
  float alc = fminf(fmaxf(alpha, -1.f), 1.f);
  float ac_ = acosf(alc);
  float om = 1.f - alpha * alpha;
  float sq = sqrt_g(om);
  float coefp = ac_ / sq;
  float w[CD];
  w[0] = coefp * (z[0] - alpha * R);
  float nw2 = w[0] * w[0];
  float nlp0 = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      w[j + 1] = coefp * z[j + 1];
      nw2 += w[j + 1] * w[j + 1];
      nlp0 += -(w[j + 1] * w[j + 1]) / 2.f - kHalfLn2Pi;
    }
  float nw = sqrtf(nw2);
  float ld0 = s_logdet_nu<false>(n, nw, R, 0.f, nullptr, nullptr);
  o.logp = nlp0 - ld0;
  o.kl = o.logq - o.logp;
  if (!BWD) return;

  float gR = 0.f;
  const float g_logq = gkl, g_logp = -gkl;
  float g_nw;
  (void)s_logdet_nu<true>(n, nw, R, -g_logp, &g_nw, &gR);
  float kw = nw > 0.f ? g_nw / nw : 0.f;
  float gw[CD];
  gw[0] = kw * w[0];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) gw[j + 1] = kw * w[j + 1] + g_logp * (-w[j + 1]);
  // w0 = coefp (z0 - alpha R); w_k = coefp z_k
  float gzt[CD];
  float g_coefp = gw[0] * (z[0] - alpha * R);
  gzt[0] = gz[0] + gw[0] * coefp;
  float g_alpha = -gw[0] * coefp * R;
  gR += -gw[0] * coefp * alpha;  // via mu0[0] = R
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      g_coefp += gw[j + 1] * z[j + 1];
      gzt[j + 1] = gz[j + 1] + gw[j + 1] * coefp;
    }
  float g_ac = g_coefp / sq;
  float g_sq = -g_coefp * ac_ / (sq * sq);
  g_alpha += g_sq * sqrt_g_d(om, sq) * (-2.f * alpha);
  if (alpha >= -1.f && alpha <= 1.f) g_alpha += g_ac * (-1.f / sqrtf(1.f - alc * alc));
  // alpha = (mu0 . z)/R^2, mu0 = [R,0..]
  gzt[0] += g_alpha * R / (R * R);
  gR += g_alpha * z[0] / (R * R);
  gR += g_alpha * (-2.f * alpha / R);
  // logq
  float g_nu;
  (void)s_logdet_nu<true>(n, nu, R, -g_logq, &g_nu, &gR);
  const float g_nlp = g_logq;
  float g_ct = 0.f, g_st = 0.f, g_t = 0.f;
  float g_mu[CD], gu[CD];
  MVAE_UNROLL
  for (int k = 0; k < CD; ++k)
    if (k < d) {
      g_ct += gzt[k] * mu[k];
      g_st += gzt[k] * un[k];
      g_mu[k] = gzt[k] * ct;
      float g_un = gzt[k] * st;
      gu[k] = g_un / t;
      g_t += -g_un * u[k] / (t * t);
    }
  g_t += -g_ct * st + g_st * ct;
  g_nu += g_t / R;
  gR += -g_t * t / R;
  if (nu > 0.f) {
    float k = g_nu / nu;
    MVAE_UNROLL
    for (int k2 = 0; k2 < CD; ++k2)
      if (k2 < d) gu[k2] += k * u[k2];
  }
  // u0 = -coef (mu0+R); u_j = v_j - coef mu_j
  float g_coef = -gu[0] * (mu[0] + R);
  g_mu[0] += -gu[0] * coef;
  gR += -gu[0] * coef;
  float g_v[CN];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      g_coef += -gu[j + 1] * mu[j + 1];
      g_v[j] = gu[j + 1];
      g_mu[j + 1] += -gu[j + 1] * coef;
    }
  float g_dp = g_coef / denom;
  float g_denom = -g_coef * coef / denom;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      g_mu[j + 1] += g_dp * v[j];
      g_v[j] += g_dp * mu[j + 1];
    }
  gR += g_denom * (2.f * R + mu[0]);
  g_mu[0] += g_denom * R;
  float g_s[CN];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      float s = sg[j];
      g_v[j] += g_nlp * (-v[j] / (s * s));
      g_s[j] = g_nlp * (v[j] * v[j] / (s * s * s) - 1.f / s) + g_v[j] * e[j];
    }
  store_gl<N>(n, l_n, l, g_s, gl);
  float g_ca = g_mu[0] * R;
  gR += g_mu[0] * ca;
  float g_sa = 0.f, g_dn = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      g_sa += g_mu[j + 1] * xn[j];
      float g_xn = g_mu[j + 1] * sa;
      gR += g_xn * m[j] / dn;
      gm[j] = g_xn * R / dn;
      g_dn += -g_xn * R * m[j] / (dn * dn);
    }
  float g_a = -g_ca * sa + g_sa * ca;
  float g_nm = (nm >= 1e-12f) ? g_dn : 0.f;
  g_nm += g_a / R;
  gR += -g_a * a / R;
  if (nm > 0.f) {
    float k = g_nm / nm;
    MVAE_UNROLL
    for (int j = 0; j < CN; ++j)
      if (j < n) gm[j] += k * m[j];
  }
  *gR_out = gR;
}

// ============================================== POINCARE BALL ==============================================
// poincare.py + geoopt 0.1.0 poincare math (MIN_NORM 1e-15, tanh clamp +-15, artanh clamp 1-1e-5).
constexpr float kPMin = 1e-15f;
MVAE_DEV float cmin_d(float x, float lo) { return x >= lo ? 1.f : 0.f; }
MVAE_DEV float tanh_c(float x) { return tanhf(fminf(fmaxf(x, -15.f), 15.f)); }
MVAE_DEV float tanh_c_d(float x, float y) { return (x >= -15.f && x <= 15.f) ? (1.f - y * y) : 0.f; }
MVAE_DEV float artanh_go(float x, float* xc_out) {
  float xc = fminf(fmaxf(x, -1.0f + 1e-5f), 1.0f - 1e-5f);
  *xc_out = xc;
  return (log1pf(xc) - log1pf(-xc)) * 0.5f;
}

// mobius_add(x, y, c) over n coordinates; sv = {x2, y2, xy, den}
template <int N>
MVAE_DEV void mobius_add(int n, const float* x, const float* y, float c, float* out, float* sv) {
  MVAE_UN(N);
  constexpr int CN = Cap<N>::n;
  constexpr int CD = Cap<N>::d;
  (void)CN; (void)CD;
  float x2 = 0.f, y2 = 0.f, xy = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      x2 += x[j] * x[j];
      y2 += y[j] * y[j];
      xy += x[j] * y[j];
    }
  float A = 1.f + 2.f * c * xy + c * y2;
  float Bc = 1.f - c * x2;
  float den = 1.f + 2.f * c * xy + c * c * x2 * y2;
  float dc = fmaxf(den, kPMin);
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) out[j] = (A * x[j] + Bc * y[j]) / dc;
  sv[0] = x2;
  sv[1] = y2;
  sv[2] = xy;
  sv[3] = den;
}
// accumulates into gx, gy (either may be nullptr), *gc
template <int N>
MVAE_DEV void mobius_add_bwd(int n, const float* x, const float* y, float c, const float* sv, const float* gout,
                             float* gx, float* gy, float* gc) {
  MVAE_UN(N);
  constexpr int CN = Cap<N>::n;
  constexpr int CD = Cap<N>::d;
  (void)CN; (void)CD;
  float x2 = sv[0], y2 = sv[1], xy = sv[2], den = sv[3];
  float A = 1.f + 2.f * c * xy + c * y2;
  float Bc = 1.f - c * x2;
  float dc = fmaxf(den, kPMin);
  float g_A = 0.f, g_B = 0.f, g_dc = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      float num = A * x[j] + Bc * y[j];
      float gn = gout[j] / dc;
      g_dc += -gout[j] * num / (dc * dc);
      g_A += gn * x[j];
      g_B += gn * y[j];
      if (gx) gx[j] += gn * A;
      if (gy) gy[j] += gn * Bc;
    }
  float g_den = g_dc * cmin_d(den, kPMin);
  float g_xy = g_A * 2.f * c + g_den * 2.f * c;
  float g_y2 = g_A * c + g_den * c * c * x2;
  float g_x2 = -g_B * c + g_den * c * c * y2;
  *gc += g_A * (2.f * xy + y2) - g_B * x2 + g_den * (2.f * xy + 2.f * c * x2 * y2);
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      if (gx) gx[j] += g_xy * y[j] + g_x2 * 2.f * x[j];
      if (gy) gy[j] += g_xy * x[j] + g_y2 * 2.f * y[j];
    }
}

// poincare_to_lorentz (poincare.py:167-170): [R(R^2+|y|^2), 2R^2 y] / (R^2 - |y|^2), |y|^2 = norm(y)**2
template <int N>
MVAE_DEV void p2l(int n, const float* y, float R, float* out) {
  MVAE_UN(N);
  constexpr int CN = Cap<N>::n;
  constexpr int CD = Cap<N>::d;
  (void)CN; (void)CD;
  float s = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) s += y[j] * y[j];
  float nrm = sqrtf(s);
  float nn = nrm * nrm;
  float den = R * R - nn;
  out[0] = R * (R * R + nn) / den;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) out[j + 1] = 2.f * (R * R) * y[j] / den;
}
template <int N>
MVAE_DEV void p2l_bwd(int n, const float* y, float R, const float* gout, float* gy, float* gR) {
  MVAE_UN(N);
  constexpr int CN = Cap<N>::n;
  constexpr int CD = Cap<N>::d;
  (void)CN; (void)CD;
  float s = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) s += y[j] * y[j];
  float nrm = sqrtf(s);
  float nn = nrm * nrm;
  float den = R * R - nn;
  float num0 = R * (R * R + nn);
  float g_den = -gout[0] * num0 / (den * den);
  float g_nn = gout[0] * R / den;
  *gR += gout[0] * (3.f * R * R + nn) / den;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      float numj = 2.f * (R * R) * y[j];
      g_den += -gout[j + 1] * numj / (den * den);
      if (gy) gy[j] += gout[j + 1] * 2.f * (R * R) / den;
      *gR += gout[j + 1] * 4.f * R * y[j] / den;
    }
  *gR += g_den * 2.f * R;
  g_nn += -g_den;
  if (gy && nrm > 0.f) {
    MVAE_UNROLL
    for (int j = 0; j < CN; ++j)
      if (j < n) gy[j] += g_nn * 2.f * nrm * (y[j] / nrm);
  }
}

// H.inverse_exp_map (hyperbolics.py:124-128) on explicit ambient vectors; sv = {alpha, coef, acosh_z, sq}
template <int N>
MVAE_DEV void h_inv_exp(int d, const float* x, const float* at, float R, float* w, float* sv) {
  MVAE_UN(N);
  constexpr int CN = Cap<N>::n;
  constexpr int CD = Cap<N>::d;
  (void)CN; (void)CD;
  float lpz = 0.f;
  MVAE_UNROLL
  for (int k = 0; k < CD; ++k)
    if (k < d) lpz += at[k] * x[k];
  lpz = lpz - 2.f * (at[0] * x[0]);
  float alpha = -lpz / (R * R);
  float zz;
  float ach = acosh_g(alpha, &zz);
  float sq = sqrt_g(alpha * alpha - 1.f);
  float coef = ach / sq;
  MVAE_UNROLL
  for (int k = 0; k < CD; ++k)
    if (k < d) w[k] = coef * (x[k] - alpha * at[k]);
  sv[0] = alpha;
  sv[1] = coef;
  sv[2] = zz;
  sv[3] = sq;
}
template <int N>
MVAE_DEV void h_inv_exp_bwd(int d, const float* x, const float* at, float R, const float* sv, const float* gw,
                            float* gx, float* gat, float* gR) {
  MVAE_UN(N);
  constexpr int CN = Cap<N>::n;
  constexpr int CD = Cap<N>::d;
  (void)CN; (void)CD;
  float alpha = sv[0], coef = sv[1], zz = sv[2], sq = sv[3];
  float ach = coef * sq;
  float g_coef = 0.f, g_alpha = 0.f;
  MVAE_UNROLL
  for (int k = 0; k < CD; ++k)
    if (k < d) {
      g_coef += gw[k] * (x[k] - alpha * at[k]);
      gx[k] += gw[k] * coef;
      g_alpha += -gw[k] * coef * at[k];
      gat[k] += -gw[k] * coef * alpha;
    }
  float g_ach = g_coef / sq;
  float g_sq = -g_coef * ach / (sq * sq);
  g_alpha += g_sq * sqrt_g_d(alpha * alpha - 1.f, sq) * 2.f * alpha;
  g_alpha += g_ach / zz;
  float g_lp = -g_alpha / (R * R);
  *gR += g_alpha * (-2.f * alpha / R);
  MVAE_UNROLL
  for (int k = 0; k < CD; ++k)
    if (k < d) {
      float sgn = (k == 0) ? -1.f : 1.f;
      gx[k] += g_lp * sgn * at[k];
      gat[k] += g_lp * sgn * x[k];
    }
}

// PoincareBall.logdet (poincare.py:84-89): H._logdet(H.inverse_exp_map(p2l(z), p2l(mu))).
// With BWD accumulates g * d(logdet) into gzp, gmu (gmu may be nullptr) and *gR.
template <int N, bool BWD>
MVAE_DEV float p_logdet(int n, const float* z, const float* mu, float R, float g, float* gzp, float* gmu, float* gR) {
  MVAE_UN(N);
  constexpr int CN = Cap<N>::n;
  constexpr int CD = Cap<N>::d;
  const int d = n + 1;
  float zs[CD], ms[CD], uu[CD], sv[4];
  p2l<N>(n, z, R, zs);
  p2l<N>(n, mu, R, ms);
  h_inv_exp<N>(d, zs, ms, R, uu, sv);
  float pr = 0.f;
  MVAE_UNROLL
  for (int k = 0; k < CD; ++k)
    if (k < d) pr += uu[k] * uu[k];
  pr = pr - 2.f * (uu[0] * uu[0]);
  if (!BWD) return h_logdet_pr<false>(n, pr, R, 0.f, nullptr, nullptr);
  float g_pr;
  float ld = h_logdet_pr<true>(n, pr, R, g, &g_pr, gR);
  float guu[CD], gzs[CD], gms[CD];
  MVAE_UNROLL
  for (int k = 0; k < CD; ++k)
    if (k < d) {
      guu[k] = g_pr * (k == 0 ? -2.f : 2.f) * uu[k];
      gzs[k] = 0.f;
      gms[k] = 0.f;
    }
  h_inv_exp_bwd<N>(d, zs, ms, R, sv, guu, gzs, gms, gR);
  p2l_bwd<N>(n, z, R, gzs, gzp, gR);
  p2l_bwd<N>(n, mu, R, gms, gmu, gR);
  return ld;
}

template <int N, bool BWD>
MVAE_DEV void comp_p(int n, int l_n, const float* m, const float* l, const float* e, float R, CompOut<N>& o,
                     const float* gz, float gkl, float* gm, float* gl, float* gR_out) {
  MVAE_UN(N);
  constexpr int CN = Cap<N>::n;
  float c = 1.f / (R * R);     // _c(radius) = 1 / radius**2 (poincare.py:108-109)
  float sc = powf(c, 0.5f);    // c ** 0.5
  // ---- encode: expmap0 ----
  float nm2 = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) nm2 += m[j] * m[j];
  float nm = sqrtf(nm2);
  float un_ = fmaxf(nm, kPMin);
  float ta = sc * un_;
  float th = tanh_c(ta);
  float* mu = o.mu;
  float mu2 = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      mu[j] = th * m[j] / (sc * un_);
      mu2 += mu[j] * mu[j];
    }
  float* sg = o.sigma;
  load_sigma<N>(n, l_n, l, sg);
  float v[CN];
  // ---- sample_projection_mu0 (poincare.py:152-157): v_ = v / lambda_mu ; z = expmap_mu(v_) ----
  float lden = 1.f - c * mu2;
  float ldc = fmaxf(lden, kPMin);
  float lam = 2.f / ldc;
  float* vv = o.u;  // data[0] = v_
  float vn2 = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      v[j] = e[j] * sg[j];
      vv[j] = v[j] / lam;
      vn2 += vv[j] * vv[j];
    }
  float vn = sqrtf(vn2);
  float vnc = fmaxf(vn, kPMin);
  float tb = sc / 2.f * lam * vnc;
  float thb = tanh_c(tb);
  float sec[CN];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) sec[j] = thb * vv[j] / (sc * vnc);
  float svm[4];
  float* z = o.z;
  mobius_add<N>(n, mu, sec, c, z, svm);
  // ---- log q ----
  float nlp = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) nlp += -(v[j] * v[j]) / (2.f * (sg[j] * sg[j])) - logf(sg[j]) - kHalfLn2Pi;
  float ld = p_logdet<N, false>(n, z, mu, R, 0.f, nullptr, nullptr, nullptr);
  o.logq = nlp - ld;
  // ---- log p: loc = 0, scale = 1 (poincare.py:160-164: logmap(0, z) * lambda_0) ----
  float zero[CN], sub[CN];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j) zero[j] = 0.f;
  float svs[4];
  mobius_add<N>(n, zero, z, c, sub, svs);
  float sn2 = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) sn2 += sub[j] * sub[j];
  float sn = sqrtf(sn2);
  float snc = fmaxf(sn, kPMin);
  const float lam0 = 2.f;
  float atx;
  float at = artanh_go(sc * snc, &atx);
  float k0 = 2.f / sc / lam0 * at;
  float x0[CN];
  float nlp0 = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      x0[j] = (k0 * sub[j] / snc) * lam0;
      nlp0 += -(x0[j] * x0[j]) / 2.f - kHalfLn2Pi;
    }
  float ld0 = p_logdet<N, false>(n, z, zero, R, 0.f, nullptr, nullptr, nullptr);
  o.logp = nlp0 - ld0;
  o.kl = o.logq - o.logp;
  if (!BWD) return;

  // ================================ reverse sweep ================================
  float gR = 0.f, g_c = 0.f, g_sc = 0.f;
  const float g_logq = gkl, g_logp = -gkl;
  float gzt[CN], g_mu[CN];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j) {
    gzt[j] = (j < n) ? gz[j] : 0.f;
    g_mu[j] = 0.f;
  }
  // logp = nlp0 - ld0
  (void)p_logdet<N, true>(n, z, zero, R, -g_logp, gzt, nullptr, &gR);
  float g_sub[CN];
  float g_k0 = 0.f, g_snc = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      float g_x0 = g_logp * (-x0[j]);
      g_k0 += g_x0 * lam0 * sub[j] / snc;
      g_sub[j] = g_x0 * lam0 * k0 / snc;
      g_snc += -g_x0 * lam0 * k0 * sub[j] / (snc * snc);
    }
  float g_at = g_k0 * 2.f / sc / lam0;
  g_sc += -g_k0 * k0 / sc;
  float g_arg = g_at / (1.f - atx * atx);
  g_sc += g_arg * snc;
  g_snc += g_arg * sc;
  float g_sn = g_snc * cmin_d(sn, kPMin);
  if (sn > 0.f) {
    float k = g_sn / sn;
    MVAE_UNROLL
    for (int j = 0; j < CN; ++j)
      if (j < n) g_sub[j] += k * sub[j];
  }
  mobius_add_bwd<N>(n, zero, z, c, svs, g_sub, nullptr, gzt, &g_c);
  // logq = nlp - ld
  (void)p_logdet<N, true>(n, z, mu, R, -g_logq, gzt, g_mu, &gR);
  const float g_nlp = g_logq;
  // z = mobius_add(mu, sec)
  float g_sec[CN];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j) g_sec[j] = 0.f;
  mobius_add_bwd<N>(n, mu, sec, c, svm, gzt, g_mu, g_sec, &g_c);
  // sec = thb * vv / (sc*vnc)
  float g_thb = 0.f, g_q = 0.f;
  float g_vv[CN];
  float q = sc * vnc;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      g_thb += g_sec[j] * vv[j] / q;
      g_vv[j] = g_sec[j] * thb / q;
      g_q += -g_sec[j] * thb * vv[j] / (q * q);
    }
  g_sc += g_q * vnc;
  float g_vnc = g_q * sc;
  float g_tb = g_thb * tanh_c_d(tb, thb);
  g_sc += g_tb * lam * vnc / 2.f;
  float g_lam = g_tb * sc / 2.f * vnc;
  g_vnc += g_tb * sc / 2.f * lam;
  float g_vn = g_vnc * cmin_d(vn, kPMin);
  if (vn > 0.f) {
    float k = g_vn / vn;
    MVAE_UNROLL
    for (int j = 0; j < CN; ++j)
      if (j < n) g_vv[j] += k * vv[j];
  }
  float g_v[CN];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      g_v[j] = g_vv[j] / lam;
      g_lam += -g_vv[j] * v[j] / (lam * lam);
    }
  float g_lden = (-g_lam * 2.f / (ldc * ldc)) * cmin_d(lden, kPMin);
  g_c += -g_lden * mu2;
  float g_mu2 = -g_lden * c;
  float g_s[CN];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      g_mu[j] += g_mu2 * 2.f * mu[j];
      float s = sg[j];
      g_v[j] += g_nlp * (-v[j] / (s * s));
      g_s[j] = g_nlp * (v[j] * v[j] / (s * s * s) - 1.f / s) + g_v[j] * e[j];
    }
  store_gl<N>(n, l_n, l, g_s, gl);
  // mu = th * m / (sc*un_)
  float p_ = sc * un_;
  float g_th = 0.f, g_p = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      g_th += g_mu[j] * m[j] / p_;
      gm[j] = g_mu[j] * th / p_;
      g_p += -g_mu[j] * th * m[j] / (p_ * p_);
    }
  g_sc += g_p * un_;
  float g_un = g_p * sc;
  float g_ta = g_th * tanh_c_d(ta, th);
  g_sc += g_ta * un_;
  g_un += g_ta * sc;
  float g_nm = g_un * cmin_d(nm, kPMin);
  if (nm > 0.f) {
    float k = g_nm / nm;
    MVAE_UNROLL
    for (int j = 0; j < CN; ++j)
      if (j < n) gm[j] += k * m[j];
  }
  g_c += g_sc * 0.5f / sc;
  gR += g_c * (-2.f / (R * R * R));
  *gR_out = gR;
}

}  // namespace mvae
