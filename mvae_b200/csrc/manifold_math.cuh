// manifold_math.cuh — guarded scalar math of the reference for the STANDALONE manifold / Wrapped-Normal operators
// (manifold_ops.cu): mt/mvae/ops/common.py:28-147 (leaky clamps, guarded sqrt / acosh, logsinh), the geoopt 0.1.0
// guards of the Poincare ball (MIN_NORM, tanh / artanh clamps), the log-det kernels of hyperbolics.py:58-65 and
// spherical.py:58-67, and the radius clamp of manifold.py:73-75.  The standalone operators restate the reference op by
// op; the fused training kernels use the closed forms of pm_math.cuh instead.
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
