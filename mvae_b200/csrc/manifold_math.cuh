// manifold_math.cuh — per-sample device arithmetic of the mixed-curvature latent components (fp32).
//
// One `comp_*<N, BWD>` function per manifold runs, for ONE sample of ONE component, the whole chain
//   Component.encode -> reparametrize -> q_z.rsample_with_parts -> kl_loss (log q - log p)
// of the reference (paths relative to the reference root):
//   mt/mvae/components/component.py:63-75, mt/mvae/sampling/sampling_procedures.py:91-116,145-155,
//   mt/mvae/distributions/wrapped_normal.py:70-103, mt/mvae/ops/{hyperbolics,spherical,euclidean,poincare}.py,
//   guarded scalar math mt/mvae/ops/common.py:28-147 (LeakyClamp / Atanh / Acosh custom backward rules),
//   geoopt==0.1.0 poincare math for the Poincare ball (un-vendored third party; constants as in DESIGN.md).
// With BWD the function also runs the reverse sweep of that chain by recomputation (hand-derived).
//
// Formulation.  The reference evaluates the chain through ambient-space vector algebra (Lorentz products,
// z - alpha*mu0, 1 - alpha^2, Poincare->Lorentz conversions) that cancels catastrophically in float32; its own
// float32 run is only 1e-3 accurate on these terms.  The kernels evaluate the SAME functions
// in the closed forms those expressions reduce to on the manifold (parallel transport is an isometry:
// |u| = |v|; the prior's tangent vector has norm dist(mu0, z); Poincare and hyperboloid log-dets coincide), which
// are well conditioned, so the float32 kernels track the reference's default float64 path to ~1e-6.  Guards of
// the reference are kept where they act on these closed forms: +-85 clamp of cosh/sinh arguments, sqrt clamp at
// 1e-9 with its leaky (1e-8) gradient, plain clamps (zero gradient) in the sphere log-det, geoopt's tanh (+-15)
// and artanh (1-1e-5) clamps and MIN_NORM.  Derivations: DESIGN.md section 4.
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

// ---- elementary functions ----------------------------------------------------------------------------------
// The kernels are instruction-issue bound, so exp / log / division / sqrt go through the SFU
// approximations (MUFU.EX2 / LG2 / RCP / RSQ, <= 2 ulp) wrapped so that the well-conditioned closed forms above keep
// their accuracy: log1p is compensated, sinh switches to its series near 0, sin / cos stay on the accurate sincosf.
#ifdef __CUDA_ARCH__
MVAE_DEV float f_exp(float x) { return __expf(x); }
MVAE_DEV float f_log(float x) { return __logf(x); }
MVAE_DEV float f_div(float a, float b) { return __fdividef(a, b); }
MVAE_DEV float f_sqrt(float x) { return x > 0.f ? x * rsqrtf(x) : 0.f; }
MVAE_DEV float f_rsqrt(float x) { return rsqrtf(x); }
#else
MVAE_DEV float f_exp(float x) { return expf(x); }
MVAE_DEV float f_log(float x) { return logf(x); }
MVAE_DEV float f_div(float a, float b) { return a / b; }
MVAE_DEV float f_sqrt(float x) { return sqrtf(x); }
MVAE_DEV float f_rsqrt(float x) { return 1.f / sqrtf(x); }
#endif
// log(1 + t), t >= 0, accurate for tiny t: log(u) * t / (u - 1) with u = fl(1 + t) cancels the rounding of u
MVAE_DEV float f_log1p(float t) {
  const float u = 1.f + t;
  return u == 1.f ? t : f_log(u) * f_div(t, u - 1.f);
}

// ---- guarded scalar math, ops/common.py --------------------------------------------------------------------
MVAE_DEV float lclamp(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
MVAE_DEV float lclamp_d(float x, float lo, float hi) { return (x >= lo && x <= hi) ? 1.f : 1e-8f; }
MVAE_DEV float lclamp_lo(float x, float lo) { return x < lo ? lo : x; }
MVAE_DEV float lclamp_lo_d(float x, float lo) { return x >= lo ? 1.f : 1e-8f; }
// sqrt (common.py:117-119)
MVAE_DEV float sqrt_g(float x) { return f_sqrt(lclamp_lo(x, 1e-9f)); }
MVAE_DEV float sqrt_g_d(float x, float y) { return lclamp_lo_d(x, 1e-9f) * f_div(0.5f, y); }
// cosh & sinh of a clamped argument in one go: e = exp(|x|)
MVAE_DEV void coshsinh_g(float x, float* ch, float* sh) {
  const float xc = lclamp(x, -kMaxHyp, kMaxHyp);
  const float ax = fabsf(xc);
  const float E = f_exp(ax), Ei = f_div(1.f, E);
  *ch = 0.5f * (E + Ei);
  const float x2 = ax * ax;
  const float s_small = ax * (1.f + x2 * (1.f / 6.f + x2 * (1.f / 120.f + x2 * (1.f / 5040.f))));  // |x| < 0.35: < 1e-9 rel
  const float s = ax < 0.35f ? s_small : 0.5f * (E - Ei);
  *sh = xc < 0.f ? -s : s;
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
MVAE_DEV float softplus(float x) { return x > 20.f ? x : fmaxf(x, 0.f) + f_log1p(f_exp(-fabsf(x))); }
MVAE_DEV float softplus_d(float x) {
  if (x > 20.f) return 1.f;
  const float e = f_exp(-fabsf(x));
  const float inv = f_div(1.f, 1.f + e);
  return x >= 0.f ? inv : e * inv;
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
  float kl;
};

// F(x) = log(sinh(x) / x) = logsinh(x) - log(x)  (hyperbolics.py:58-65 with common.py:122-128), x > 0.
// 1 - e^{-2x} is taken from expm1 so that small x does not cancel.
MVAE_DEV float log_sinhc(float x) {
  if (x < 0.5f) {  // x^2/6 - x^4/180 + x^6/2835 - x^8/37800
    const float x2 = x * x;
    return x2 * (1.f / 6.f - x2 * (1.f / 180.f - x2 * (1.f / 2835.f - x2 * (1.f / 37800.f))));
  }
  return x + f_log(f_div(1.f - f_exp(-2.f * x), 2.f * x));
}
// F'(x) = coth(x) - 1/x
MVAE_DEV float log_sinhc_d(float x) {
  if (x < 0.25f) {
    float x2 = x * x;
    return x * (1.f / 3.f - x2 * (1.f / 45.f - x2 * (2.f / 945.f - x2 * (1.f / 4725.f))));
  }
  const float E = f_exp(-2.f * x);
  return f_div(1.f + E, 1.f - E) - f_div(1.f, x);
}
// G(x) = log clamp(|sin x|, 1e-5) - log clamp(x, 1e-5)  (spherical.py:58-67, plain clamps: zero gradient outside)
MVAE_DEV float log_sinc_abs(float x, float sn) {
  return f_log(f_div(fmaxf(fabsf(sn), 1e-5f), fmaxf(x, 1e-5f)));
}
MVAE_DEV float log_sinc_abs_d(float x, float sn, float cs) {
  float as = fabsf(sn);
  float d = 0.f;
  if (as >= 1e-5f) d += f_div(sn > 0.f ? cs : -cs, as);
  if (x >= 1e-5f) d -= f_div(1.f, x);
  return d;
}

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
  float kl = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      float mu = 0.5f * m[j];
      float s = o.sigma[j];
      o.mu[j] = mu;
      o.z[j] = mu + e[j] * s;
      float var_ratio = s * s;
      kl += 0.5f * (var_ratio + mu * mu - 1.f) - f_log(s);
    }
  o.kl = kl;
  if (!BWD) return;
  float g_s[CN];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      float mu = o.mu[j], s = o.sigma[j];
      gm[j] = 0.5f * (gz[j] + gkl * mu);
      g_s[j] = gz[j] * e[j] + gkl * (s - f_div(1.f, s));
    }
  store_gl<N>(n, l_n, l, g_s, gl);
}

// ---- geoopt 0.1.0 guards used by the Poincare ball (poincare.py) ----
// poincare.py + geoopt 0.1.0 poincare math (MIN_NORM 1e-15, tanh clamp +-15, artanh clamp 1-1e-5).
constexpr float kPMin = 1e-15f;
MVAE_DEV float cmin_d(float x, float lo) { return x >= lo ? 1.f : 0.f; }
MVAE_DEV float tanh_c(float x) { return tanhf(fminf(fmaxf(x, -15.f), 15.f)); }
// sech^2 of the clamped argument = 1 - tanh_c(x)^2, from one exponential (exact to rounding for large |x|)
MVAE_DEV float sech2_c(float x) {
  float e = f_exp(-2.f * fminf(fabsf(x), 15.f));
  float d = 1.f + e;
  return f_div(4.f * e, d * d);
}
MVAE_DEV float tanh_c_d(float x, float) { return (x >= -15.f && x <= 15.f) ? sech2_c(x) : 0.f; }
MVAE_DEV float artanh_go(float x, float* xc_out) {
  float xc = fminf(fmaxf(x, -1.0f + 1e-5f), 1.0f - 1e-5f);
  *xc_out = xc;
  return (log1pf(xc) - log1pf(-xc)) * 0.5f;
}

// ============================= HYPERBOLOID, SPHERE and POINCARE BALL: one geodesic triangle =============================
// Reference chain (H: hyperbolics.py, S: spherical.py), with a = |m|/R, mh = m/max(|m|,1e-12), v = eps*sigma,
// p = <mh, v>, t = |v|/R:
//   exp_map_mu0 (:114-121 / :94-101)      mu = [R C(a), R S(a) mh]                       C,S = cosh,sinh | cos,sin
//   parallel_transport_mu0 (:87-93 / :74-77)
//                                         u = [sg S(a) p,  v + (C(a) - 1) p mh]          sg = +1 (H) | -1 (S);  |u| = |v|
//   exp_map (:106-111 / :86-91)           z = C(t) mu + S(t)/t u
//     => z0 = R C(t) C(a) + sg A S(a) p,  z_tail = A v + Bc mh,  A = S(t)/t,  Bc = R C(t) S(a) + A (C(a)-1) p
//   log q (wrapped_normal.py:84-97)       sum_j logN(v_j; 0, sigma_j) - (n-1)(log R + Fq(t))
//   log p (wrapped_normal.py:99-103, inverse_exp_map :124-128 / :104-109, inverse PT :96-103 / :80-83)
//                                         -r^2 R^2/2 - n ln sqrt(2pi) - (n-1)(log R + Fq(r)),  r = dist(mu0, z)/R
//     H: r = acosh(z0/R) = asinh(|z_tail|/R), Fq = log(sinh x / x)       (logdet :58-65)
//     S: r = acos(z0/R)  = atan2(|z_tail|/R, z0/R), Fq = log clamp|sin x| - log clamp x   (logdet :58-67)
//   KL = log q - log p = -sum eps^2/2 - sum log sigma + R^2 r^2/2 - (n-1)(Fq(t) - Fq(r))
// Poincare ball (poincare.py + geoopt 0.1.0; d = n): the ball of radius R is the hyperboloid seen through
// lorentz_to_poincare (hyperbolics.py:151-152).  exp_map_mu0 (:132-137) gives mu = R tanh(a) mh, i.e. the hyperboloid
// point at distance 2|m|; sample_projection_mu0 (:152-157: v_ = v/lambda_mu, expmap_mu(v_)) is the point at geodesic
// distance lambda_mu |v_| = |v| from mu in the (conformal) direction of v.  So z_P = R Z_tail / (R + Z_0) with Z the
// hyperboloid sample above evaluated at a -> 2a.  log q: PoincareBall.logdet (:84-89) maps to the Lorentz model and
// takes H._logdet of the log map, whose norm is dist(mu, z) = |v|: the same Fq(t).  log p (:160-164): |logmap_0(z)|
// lambda_0 = 2R artanh(|z|/R) = R r with geoopt's artanh clamp (|z|/R <= 1 - 1e-5, i.e. r <= 12.206), while the
// log-det sees the unclamped r.  geoopt's tanh clamp (+-15) bounds a and t/2.
enum { kHyp = 0, kSph = 1, kPoi = 2 };

template <int N, bool BWD, int KIND>
MVAE_DEV void comp_hsp(int n, int l_n, const float* m, const float* l, const float* e, float R, CompOut<N>& o,
                       const float* gz, float gkl, float* gm, float* gl, float* gR_out) {
  MVAE_UN(N);
  constexpr int CN = Cap<N>::n;
  constexpr bool HYP = KIND != kSph;  // hyperbolic trigonometry
  constexpr bool POI = KIND == kPoi;
  const float sgn = HYP ? 1.f : -1.f;
  // ---- encode ----
  float nm2 = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) nm2 += m[j] * m[j];
  // the sphere's log-det is singular at |v| = pi R: its angles keep correctly rounded sqrt / division
  const float nm = HYP ? f_sqrt(nm2) : sqrtf(nm2);
  const float nmin = POI ? kPMin : 1e-12f;  // geoopt MIN_NORM | F.normalize eps
  const float dn = fmaxf(nm, nmin);
  const float iR = f_div(1.f, R);
  const float idn = f_div(1.f, dn);
  const float a = HYP ? (POI ? dn : nm) * iR : nm / R;
  const bool a_sat = POI && a > 15.f;       // geoopt tanh clamp (plain: zero gradient beyond)
  const float aa = POI ? 2.f * fminf(a, 15.f) : a;
  float ca, sa;
  if (HYP) {
    coshsinh_g(aa, &ca, &sa);
  } else {
    sincosf(aa, &sa, &ca);
  }
  // C(a) - 1 without cancellation: H: S^2/(C+1);  S: -S^2/(1+C) (falls back to C-1 near a = pi)
  const float cam1 = HYP ? f_div(sa * sa, ca + 1.f) : (ca > -0.5f ? -f_div(sa * sa, 1.f + ca) : ca - 1.f);
  float* sg = o.sigma;
  load_sigma<N>(n, l_n, l, sg);
  float mh[CN], v[CN];
  float Sv = 0.f, p = 0.f, se2 = 0.f, slog = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      mh[j] = m[j] * idn;
      v[j] = e[j] * sg[j];
      Sv += v[j] * v[j];
      p += mh[j] * v[j];
      se2 += e[j] * e[j];
      slog += f_log(sg[j]);
    }
  // ---- sample ----
  float ln, t, ct, st;
  bool t_sat = false;
  if (HYP) {
    ln = sqrt_g(Sv);
    t = ln * iR;
    t_sat = POI && t > 30.f;
    coshsinh_g(POI ? fminf(t, 30.f) : t, &ct, &st);
  } else {
    ln = sqrtf(Sv);
    t = ln / R;
    sincosf(t, &st, &ct);
  }
  const float it = t > 0.f ? f_div(1.f, t) : 0.f;
  const float A = t > 0.f ? st * it : 1.f;
  const float z0 = R * ct * ca + sgn * A * sa * p;
  const float Bc = R * ct * sa + A * cam1 * p;
  float* z = o.z;
  float* mu = o.mu;
  float zt[CN];
  float zt2 = 0.f;
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      zt[j] = A * v[j] + Bc * mh[j];
      zt2 += zt[j] * zt[j];
    }
  const float iRz = POI ? f_div(1.f, R + z0) : 0.f;
  const float pj = POI ? R * iRz : 1.f;  // lorentz_to_poincare
  if (POI) {
    const float Ta = f_div(sa, ca + 1.f);  // tanh(a) = sinh(2a) / (cosh(2a) + 1)
    MVAE_UNROLL
    for (int j = 0; j < CN; ++j)
      if (j < n) {
        z[j] = pj * zt[j];
        mu[j] = R * Ta * mh[j];
      }
  } else {
    z[0] = z0;
    mu[0] = R * ca;
    MVAE_UNROLL
    for (int j = 0; j < CN; ++j)
      if (j < n) {
        z[j + 1] = zt[j];
        mu[j + 1] = R * sa * mh[j];
      }
  }
  // ---- prior distance r = dist(mu0, z)/R from the tail norm (well conditioned everywhere) ----
  const float s2 = zt2 * (iR * iR);
  const float s = f_sqrt(s2);
  float r, Fr, Ft, alpha = 0.f, as_ = 0.f, snr = 0.f, csr = 0.f;
  bool r_clamped = false, at_clamped = false;
  const float kAtMax = 12.206062f;  // 2 artanh(1 - 1e-5)
  float rq;                         // distance entering the Gaussian term of log p
  if (HYP) {
    as_ = f_sqrt(1.f + s2);
    r = s > 0.5f ? f_log(s + as_) : f_log1p(s + f_div(s2, 1.f + as_));  // asinh(s)
    // H._logdet applies sqrt() (clamp 1e-9) to the squared Lorentz norm R^2 r^2 of the prior's tangent vector
    r_clamped = (R * R) * (r * r) < 1e-9f;
    const float rl = r_clamped ? 3.1622776e-5f * iR : r;
    Fr = log_sinhc(rl);
    Ft = log_sinhc(t);
    at_clamped = POI && r > kAtMax;
    rq = at_clamped ? kAtMax : r;
  } else {
    alpha = z0 * iR;
    r = atan2f(s, alpha);
    const float inv_q = f_rsqrt(alpha * alpha + s2);  // (alpha, s) is a unit vector up to rounding
    snr = s * inv_q;                                  // sin r and cos r without going through r
    csr = alpha * inv_q;
    Fr = log_sinc_abs(r, snr);
    Ft = log_sinc_abs(t, st);
    rq = r;
  }
  const float nm1 = (float)(n - 1);
  o.kl = -0.5f * se2 - slog + 0.5f * (R * R) * (rq * rq) - nm1 * (Ft - Fr);
  if (!BWD) return;

  // ================================ reverse sweep ================================
  float gR = 0.f;
  // KL -> r
  float g_r = gkl * (R * R) * rq;
  gR += gkl * R * (rq * rq);
  if (at_clamped) {
    // geoopt Artanh.backward = g / (1 - x'^2) on the clamped argument x' = 1 - 1e-5; d(rho)/d(r) = sech^2(r/2) / 2
    g_r *= sech2_c(0.5f * r) * (1.f / (1e-5f * (2.f - 1e-5f)));
  }
  float g_s, g_alpha = 0.f;
  if (HYP) {
    if (!r_clamped) {
      g_r += gkl * nm1 * log_sinhc_d(r);
    } else {
      const float rl = 3.1622776e-5f * iR;
      gR += -(gkl * nm1 * log_sinhc_d(rl)) * rl * iR;  // leaky clamp: the 1e-8 * g path into r is dropped
    }
    g_s = f_div(g_r, as_);
  } else {
    g_r += gkl * nm1 * log_sinc_abs_d(r, snr, csr);
    // r = atan2(s, alpha): any smooth extension off the constraint alpha^2 + s^2 = 1 has the same total derivative
    const float iq2 = f_div(1.f, alpha * alpha + s2);
    g_s = g_r * alpha * iq2;
    g_alpha = -g_r * s * iq2;
  }
  // s = |z_tail| / R ; alpha = z0 / R
  const float k_zt = s > 0.f ? f_div(g_s * (iR * iR), s) : 0.f;
  gR += -g_s * s * iR;
  float g_z0 = 0.f;
  if (POI) {
    // z_j = R Z_j / (R + Z_0)
    MVAE_UNROLL
    for (int j = 0; j < CN; ++j)
      if (j < n) {
        g_z0 += -gz[j] * z[j] * iRz;
        gR += gz[j] * z[j] * (iR - iRz);
      }
  } else {
    g_z0 = gz[0];
    if (!HYP) {
      g_z0 += g_alpha * iR;
      gR += -g_alpha * alpha * iR;
    }
  }
  // z0 = R ct ca + sgn A sa p ;  Bc = R ct sa + A cam1 p ; z_tail = A v + Bc mh
  float g_A = 0.f, g_Bc = 0.f;
  float g_v[CN], g_mh[CN];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      const float G = (POI ? pj * gz[j] : gz[j + 1]) + k_zt * zt[j];
      g_A += G * v[j];
      g_Bc += G * mh[j];
      g_v[j] = A * G;
      g_mh[j] = Bc * G;
    }
  float g_ct = g_z0 * R * ca + g_Bc * R * sa;
  float g_ca = g_z0 * R * ct + g_Bc * A * p;  // d(cam1)/d(ca) = 1
  float g_sa = g_z0 * sgn * A * p + g_Bc * R * ct;
  float g_p = g_z0 * sgn * A * sa + g_Bc * A * cam1;
  gR += g_z0 * ct * ca + g_Bc * ct * sa;
  g_A += g_z0 * sgn * sa * p + g_Bc * cam1 * p;
  // A = st / t ; ct, st functions of t ; KL has -(n-1) Fq(t)
  float g_t = 0.f;
  if (t > 0.f) {
    const float g_st = g_A * it;
    g_t += -g_A * A * it;
    if (HYP) {
      const float dclamp = POI ? (t_sat ? 0.f : 1.f) : lclamp_d(t, -kMaxHyp, kMaxHyp);
      g_t += (g_ct * st + g_st * ct) * dclamp;
      g_t += -gkl * nm1 * log_sinhc_d(t);
    } else {
      g_t += -g_ct * st + g_st * ct;
      g_t += -gkl * nm1 * log_sinc_abs_d(t, st, ct);
    }
  }
  // t = ln / R ; ln = sqrt(Sv) ; Sv = <v, v>
  gR += -g_t * t * iR;
  const float g_ln = g_t * iR;
  const float g_Sv = HYP ? g_ln * sqrt_g_d(Sv, ln) : (ln > 0.f ? g_ln * f_div(0.5f, ln) : 0.f);
  // a : ca, sa
  float g_a;
  if (POI) g_a = a_sat ? 0.f : 2.f * (g_ca * sa + g_sa * ca);
  else if (HYP) g_a = (g_ca * sa + g_sa * ca) * lclamp_d(a, -kMaxHyp, kMaxHyp);
  else g_a = -g_ca * sa + g_sa * ca;
  gR += -g_a * a * iR;
  float g_nm = POI ? 0.f : g_a * iR;   // P: a = max(|m|, MIN_NORM) / R
  float g_dn = POI ? g_a * iR : 0.f;
  // v = eps * sigma ; p = <mh, v> ; mh = m / dn
  float g_s_[CN];
  MVAE_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      const float gv = g_v[j] + 2.f * g_Sv * v[j] + g_p * mh[j];
      g_s_[j] = gv * e[j] - f_div(gkl, sg[j]);
      const float gmh = g_mh[j] + g_p * v[j];
      gm[j] = gmh * idn;
      g_dn += -gmh * mh[j] * idn;
    }
  store_gl<N>(n, l_n, l, g_s_, gl);
  if (nm >= nmin) g_nm += g_dn;
  if (nm > 0.f) {
    const float k = f_div(g_nm, nm);
    MVAE_UNROLL
    for (int j = 0; j < CN; ++j)
      if (j < n) gm[j] += k * m[j];
  }
  *gR_out = gR;
}

template <int N, bool BWD>
MVAE_DEV void comp_h(int n, int l_n, const float* m, const float* l, const float* e, float R, CompOut<N>& o,
                     const float* gz, float gkl, float* gm, float* gl, float* gR_out) {
  comp_hsp<N, BWD, kHyp>(n, l_n, m, l, e, R, o, gz, gkl, gm, gl, gR_out);
}
template <int N, bool BWD>
MVAE_DEV void comp_s(int n, int l_n, const float* m, const float* l, const float* e, float R, CompOut<N>& o,
                     const float* gz, float gkl, float* gm, float* gl, float* gR_out) {
  comp_hsp<N, BWD, kSph>(n, l_n, m, l, e, R, o, gz, gkl, gm, gl, gR_out);
}
template <int N, bool BWD>
MVAE_DEV void comp_p(int n, int l_n, const float* m, const float* l, const float* e, float R, CompOut<N>& o,
                     const float* gz, float gkl, float* gm, float* gl, float* gR_out) {
  comp_hsp<N, BWD, kPoi>(n, l_n, m, l, e, R, o, gz, gkl, gm, gl, gR_out);
}

// H._logdet (hyperbolics.py:58-65) on the Lorentz squared norm `pr` of u (standalone ops)
template <bool BWD>
MVAE_DEV float h_logdet_pr(int n, float pr, float R, float g, float* g_pr, float* gR) {
  float s = sqrt_g(pr);
  float r = s / R;
  return (float)(n - 1) * (logf(R) + log_sinhc(r));
}
// S._logdet (spherical.py:58-67) on nu = ||u||_2 (standalone ops)
template <bool BWD>
MVAE_DEV float s_logdet_nu(int n, float nu, float R, float g, float* g_nu, float* gR) {
  float r = nu / R;
  return (float)(n - 1) * (logf(R) + log_sinc_abs(r, sinf(r)));
}


}  // namespace mvae
