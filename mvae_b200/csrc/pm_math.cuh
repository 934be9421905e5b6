// pm_math.cuh — per-sample fp32 arithmetic of the fused product-manifold kernels (pm_kernels_impl.cuh).
//
// One `comp_*<N, BWD, ...>` function per manifold family runs, for ONE sample of ONE component, the whole chain
//   Component.encode -> reparametrize -> q_z.rsample_with_parts -> kl_loss (log q - log p)
// of the reference (paths relative to the reference root):
//   mt/mvae/components/component.py:63-75, mt/mvae/sampling/sampling_procedures.py:91-116,145-155,
//   mt/mvae/distributions/wrapped_normal.py:70-103, mt/mvae/ops/{hyperbolics,spherical,euclidean,poincare}.py,
//   guarded scalar math mt/mvae/ops/common.py:28-147 (LeakyClamp / Atanh / Acosh custom backward rules),
//   geoopt==0.1.0 poincare math for the Poincare ball (un-vendored third party; constants as in DESIGN.md §4).
// With BWD the function also runs the hand-derived reverse sweep of that chain by recomputation.
//
// Formulation: the closed forms of DESIGN.md §4 (geodesic triangle mu0 - mu - z; parallel transport is an isometry;
// the prior's tangent vector has norm dist(mu0, z); Poincare and hyperboloid log-dets coincide).  They are well
// conditioned in float32 where the reference's ambient-space algebra cancels, so the kernels track the reference's
// default float64 path to ~1e-6.  Guards of the reference are kept where they act on these forms.
//
// Instruction economy: the kernels are instruction-issue bound (5.2 SM-cycles per sample at 100 % of HBM for
// h2,s2,e2), so every transcendental is ONE MUFU op (ex2 / lg2 / rcp / rsqrt / sqrt .approx.ftz through inline PTX; the
// CUDA intrinsics without -ftz carry denormal fix-ups), wrapped in compensated forms where conditioning needs it
// (log1p, sinh near 0), sin / cos / atan2 are short Cody-Waite + minimax polynomials (~1.5 ulp), and the two branches
// of log(sinh x / x) are blended instead of branched.
//
// The header also compiles as plain C++ (no CUDA) with libm stand-ins for the MUFU ops: tests/test_pm_math_host.py
// builds it with g++ to check the algebra against the oracle without a GPU (test infrastructure, never shipped).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PM_DEV __device__ __forceinline__
#else
#define PM_DEV static inline
#endif
#if defined(__CUDACC__)
#pragma nv_diag_suppress 128  // "loop is not reachable" in the forward-only instantiations
#endif

namespace mvae {
namespace pm {

constexpr int kDynMaxN = 160;  // widest single component served (tangent dimension)

template <int N>
struct Cap {
  static constexpr int n = N > 0 ? N : kDynMaxN;
  static constexpr int d = n + 1;
};

// Loops over coordinates are fully unrolled when the dimension is static and left rolled for the dynamic path;
// every function using PM_UNROLL defines `constexpr int UN`.
#define PM_UNROLL _Pragma("unroll UN")
#define PM_UN(N) constexpr int UN = (N) > 0 ? (N) + 1 : 1

constexpr float kLn2 = 0.6931471805599453f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kPi = 3.14159265358979323846f;
constexpr float kHalfPi = 1.57079632679489661923f;
constexpr float kMaxHyp = 85.f;        // common.py:107-114 clamp of cosh/sinh arguments
constexpr float kPMin = 1e-15f;        // geoopt MIN_NORM
constexpr float kAtMax = 12.206062f;   // 2 artanh(1 - 1e-5): geoopt's artanh clamp seen as a distance
constexpr float kSqrtClamp = 3.1622776e-5f;  // sqrt(1e-9): common.py:117-119

// ---- one-instruction transcendentals ---------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
PM_DEV float f_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
PM_DEV float f_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
PM_DEV float f_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
PM_DEV float f_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
PM_DEV float f_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
PM_DEV int f_as_int(float x) { return __float_as_int(x); }
#else
PM_DEV float f_ex2(float x) { return exp2f(x); }
PM_DEV float f_lg2(float x) { return log2f(x); }
PM_DEV float f_rcp(float x) { return 1.f / x; }
PM_DEV float f_rsqrt(float x) { return 1.f / sqrtf(x); }
PM_DEV float f_sqrt(float x) { return sqrtf(x); }
PM_DEV int f_as_int(float x) { int i; memcpy(&i, &x, 4); return i; }
#endif
PM_DEV float f_exp(float x) { return f_ex2(x * kLog2e); }
PM_DEV float f_log(float x) { return f_lg2(x) * kLn2; }
// log(1 + t), t >= 0, accurate for tiny t: log(u) * t / (u - 1) with u = fl(1 + t) cancels the rounding of u
PM_DEV float f_log1p(float t) {
  const float u = 1.f + t;
  const float q = (f_lg2(u) * kLn2) * (t * f_rcp(u - 1.f));
  return u == 1.f ? t : q;
}

// cosh & sinh of x >= 0 clamped at 85 (common.py:107-114) from one exponential; sinh switches to its series near 0
PM_DEV void coshsinh_pos(float x, float* ch, float* sh) {
  const float xc = fminf(x, kMaxHyp);
  const float E = f_ex2(xc * kLog2e);
  const float hE = 0.5f * E, hEi = f_rcp(E + E);
  *ch = hE + hEi;
  const float x2 = xc * xc;
  float p = fmaf(x2, 1.f / 5040.f, 1.f / 120.f);
  p = fmaf(p, x2, 1.f / 6.f);
  p = p * x2;
  const float s_small = fmaf(p, xc, xc);  // x < 0.35: < 1e-9 relative
  *sh = xc < 0.35f ? s_small : hE - hEi;
}

// sin & cos, Cody-Waite reduction by pi/2 (three constants) + minimax polynomials on [-pi/4, pi/4]: ~1.5 ulp for
// |x| < 2^15; larger arguments (a sample 10^4 radii away from mu0) go to the library routine.
PM_DEV void sincos_cw(float x, float* sn, float* cs) {
  if (fabsf(x) > 32768.f) {
#if defined(__CUDA_ARCH__)
    sincosf(x, sn, cs);
#else
    *sn = sinf(x);
    *cs = cosf(x);
#endif
    return;
  }
  float j = fmaf(x, 0.636619747f, 12582912.f);
  const int q = f_as_int(j);
  j -= 12582912.f;
  float a = fmaf(j, -0x1.921fb0p+00f, x);
  a = fmaf(j, -0x1.5110b4p-22f, a);
  a = fmaf(j, -0x1.846988p-48f, a);
  const float s2 = a * a;
  float c = 2.44677067e-5f;
  c = fmaf(c, s2, -1.38877297e-3f);
  c = fmaf(c, s2, 4.16666567e-2f);
  c = fmaf(c, s2, -5.00000000e-1f);
  c = fmaf(c, s2, 1.f);
  float s = 2.86567956e-6f;
  s = fmaf(s, s2, -1.98559923e-4f);
  s = fmaf(s, s2, 8.33338592e-3f);
  s = fmaf(s, s2, -1.66666672e-1f);
  s = fmaf(s, a * s2, a);
  const float rs = (q & 1) ? c : s;
  const float rc = (q & 1) ? s : c;
  *sn = (q & 2) ? -rs : rs;
  *cs = ((q + 1) & 2) ? -rc : rc;
}

// atan2(y, x) for y >= 0 -> [0, pi]: odd minimax polynomial of min/max on [0, 1] (~2 ulp)
PM_DEV float atan2_pos(float y, float x) {
  const float ax = fabsf(x);
  const float mx = fmaxf(ax, y), mn = fminf(ax, y);
  const float q = mx > 0.f ? mn * f_rcp(mx) : 0.f;
  const float s = q * q;
  float r = 0.00282363896258175373077393f;
  r = fmaf(r, s, -0.0159569028764963150024414f);
  r = fmaf(r, s, 0.0425049886107444763183594f);
  r = fmaf(r, s, -0.0748900920152664184570312f);
  r = fmaf(r, s, 0.106347933411598205566406f);
  r = fmaf(r, s, -0.142027363181114196777344f);
  r = fmaf(r, s, 0.199926957488059997558594f);
  r = fmaf(r, s, -0.333331018686294555664062f);
  r = fmaf(r * s, q, q);
  r = y > ax ? kHalfPi - r : r;
  return x < 0.f ? kPi - r : r;
}

// F.softplus (beta 1, threshold 20); e = exp(-|x|) is returned for the derivative
PM_DEV float softplus(float x) {
  const float e = f_ex2(fabsf(x) * -kLog2e);
  const float y = fmaxf(x, 0.f) + f_log1p(e);
  return x > 20.f ? x : y;
}
PM_DEV float softplus_d(float x) {
  const float e = f_ex2(fabsf(x) * -kLog2e);
  const float inv = f_rcp(1.f + e);
  const float d = x >= 0.f ? inv : e * inv;
  return x > 20.f ? 1.f : d;
}

// F(x) = log(sinh(x) / x) = logsinh(x) - log(x)  (hyperbolics.py:58-65 with common.py:122-128), x > 0, ix = 1/x.
// Both forms are evaluated and blended (samples of one warp straddle the switch point all the time).
PM_DEV float log_sinhc(float x, float ix) {
  const float x2 = x * x;
  float s = fmaf(x2, -1.f / 2835.f, 1.f / 180.f);
  s = fmaf(-s, x2, 1.f / 6.f);
  s = s * x2;  // x^2/6 - x^4/180 + x^6/2835
  const float E = f_ex2(x * (-2.f * kLog2e));
  const float g = (1.f - E) * (0.5f * ix);
  const float big = fmaf(kLn2, f_lg2(g), x);
  return x < 0.25f ? s : big;
}
// F'(x) = coth(x) - 1/x
PM_DEV float log_sinhc_d(float x, float ix) {
  const float x2 = x * x;
  float s = fmaf(x2, -1.f / 4725.f, 2.f / 945.f);
  s = fmaf(-s, x2, 1.f / 45.f);
  s = fmaf(-s, x2, 1.f / 3.f);
  s = s * x;
  const float E = f_ex2(x * (-2.f * kLog2e));
  const float big = (1.f + E) * f_rcp(1.f - E) - ix;
  return x < 0.25f ? s : big;
}

// radius = clamp(relu(R_param), 1e-8, 1e8) (manifold.py:73-75), plain clamp
PM_DEV float radius_of(float rp) {
  float r = rp > 0.f ? rp : 0.f;
  return r < 1e-8f ? 1e-8f : (r > 1e8f ? 1e8f : r);
}
PM_DEV float radius_d(float rp) {
  if (!(rp > 0.f)) return 0.f;
  return (rp >= 1e-8f && rp <= 1e8f) ? 1.f : 0.f;
}

// Per-component constants derived from the radius parameter (the curvature is -+1/R^2).  Only R and 1/R are kept
// (registers are the scarce resource of the persistent kernels); R^2, 1/R^2 are one multiply away.
struct CompConst {
  float R, iR;
};
PM_DEV CompConst make_const(float rp) {
  CompConst k;
  k.R = radius_of(rp);
  k.iR = 1.f / k.R;
  return k;
}

// Per-sample result of one component.
template <int N>
struct CompOut {
  float mu[Cap<N>::d];
  float sigma[Cap<N>::n];
  float z[Cap<N>::d];
  float kl;
};

// sigma_j = softplus(l_j) + 1e-5 (component.py:69-72; scalar parametrization repeats one value, wrapped_normal.py:46-49)
template <int N>
PM_DEV void load_sigma(int n, int l_n, const float* l, float* sg) {
  PM_UN(N);
  if (N != 1 && l_n == 1) {  // (for n = 1 the two parametrizations coincide)
    const float s = softplus(l[0]) + 1e-5f;
    PM_UNROLL
    for (int j = 0; j < Cap<N>::n; ++j)
      if (j < n) sg[j] = s;
  } else {
    PM_UNROLL
    for (int j = 0; j < Cap<N>::n; ++j)
      if (j < n) sg[j] = softplus(l[j]) + 1e-5f;
  }
}
// sum_j log sigma_j with one logarithm per four factors (sigma >= 1e-5: a product of four stays normal)
template <int N>
PM_DEV float sum_log(int n, const float* sg) {
  PM_UN(N);
  float acc = 0.f, prod = 1.f;
  PM_UNROLL
  for (int j = 0; j < Cap<N>::n; ++j)
    if (j < n) {
      prod *= sg[j];
      if ((j & 3) == 3 || j == n - 1) {
        acc += f_lg2(prod);
        prod = 1.f;
      }
    }
  return acc * kLn2;
}
// d(loss)/d l from d(loss)/d sigma
template <int N>
PM_DEV void store_gl(int n, int l_n, const float* l, const float* g_s, float* gl) {
  PM_UN(N);
  if (N != 1 && l_n == 1) {
    float acc = 0.f;
    PM_UNROLL
    for (int j = 0; j < Cap<N>::n; ++j)
      if (j < n) acc += g_s[j];
    gl[0] = acc * softplus_d(l[0]);
  } else {
    PM_UNROLL
    for (int j = 0; j < Cap<N>::n; ++j)
      if (j < n) gl[j] = g_s[j] * softplus_d(l[j]);
  }
}

// ================================================ EUCLIDEAN ================================================
// euclidean.py:78-79 (mu = m/2); EuclideanNormalProcedure (sampling_procedures.py:145-155):
// z = mu + eps*sigma (wrapped_distributions.py:25-27), KL(N(mu,sigma)||N(0,1)).sum(-1).
template <int N, bool BWD>
PM_DEV void comp_e(int n, int l_n, const float* m, const float* l, const float* e, CompOut<N>& o, const float* gz,
                   float gkl, float* gm, float* gl) {
  PM_UN(N);
  constexpr int CN = Cap<N>::n;
  load_sigma<N>(n, l_n, l, o.sigma);
  float acc = 0.f;
  PM_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      const float mu = 0.5f * m[j];
      const float s = o.sigma[j];
      o.mu[j] = mu;
      o.z[j] = fmaf(e[j], s, mu);
      acc = fmaf(s, s, acc);
      acc = fmaf(mu, mu, acc);
    }
  o.kl = 0.5f * (acc - (float)n) - sum_log<N>(n, o.sigma);
  if (!BWD) return;
  float g_s[CN];
  PM_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      const float mu = o.mu[j], s = o.sigma[j];
      gm[j] = 0.5f * fmaf(gkl, mu, gz[j]);
      g_s[j] = fmaf(gz[j], e[j], gkl * (s - f_rcp(s)));
    }
  store_gl<N>(n, l_n, l, g_s, gl);
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
// Stereographically projected sphere ('d', spherical_projected.py; d = n): the same construction over the sphere.
// exp_map_mu0 (:150-154) gives mu = R tan(a) mh = the sphere point at distance 2|m| seen through spherical_to_projected
// (spherical.py:132-133); sample_projection_mu0 (:176-179, exp_map :141-147 with geoopt's mobius_add at c = -1/R^2)
// lands at geodesic distance |v| from mu in the conformal direction of v: z_D = R Z_tail / (R + Z_0), Z the sphere
// sample at a -> 2a.  logdet (:58-92) maps z and mu back to the sphere and takes S._logdet of the log map, whose norm
// is dist(mu, z) = R acos(cos t) — the WRAPPED angle tw in [0, pi], not t (they differ once |v| > pi R).  log p
// (:182-186, :157-162): |x| = 2R atan(|z|/R) = R r, r = dist(mu0, Z)/R; no clamps besides MIN_NORM on norms.
enum { kHyp = 0, kSph = 1, kPoi = 2, kPsp = 3 };

template <int N, bool BWD, int KIND, bool WANT_MS>
PM_DEV void comp_hsp(int n, int l_n, const float* m, const float* l, const float* e, const CompConst& K,
                     CompOut<N>& o, const float* gz, float gkl, float* gm, float* gl, float* gR_out) {
  PM_UN(N);
  constexpr int CN = Cap<N>::n;
  constexpr bool HYP = KIND == kHyp || KIND == kPoi;  // hyperbolic trigonometry
  constexpr bool POI = KIND == kPoi;
  constexpr bool PSP = KIND == kPsp;
  constexpr bool PROJ = POI || PSP;  // conformal model of H / S: a -> 2a, projected coordinates, MIN_NORM clamps
  const float R = K.R, iR = K.iR;
  // ---- encode ----
  float nm2 = 0.f;
  PM_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) nm2 = fmaf(m[j], m[j], nm2);
  const float nm = f_sqrt(nm2);
  const float nmin = PROJ ? kPMin : 1e-12f;  // geoopt MIN_NORM | F.normalize eps
  const float dn = fmaxf(nm, nmin);
  const float idn = f_rcp(dn);
  const float a = (PROJ ? dn : nm) * iR;
  const bool a_sat = POI && a > 15.f;       // geoopt tanh clamp (plain: zero gradient beyond)
  const float aa = POI ? 2.f * fminf(a, 15.f) : (PSP ? 2.f * a : a);
  float ca, sa;
  if (HYP) coshsinh_pos(aa, &ca, &sa);
  else sincos_cw(aa, &sa, &ca);
  // C(a) - 1 without cancellation: H: S^2/(C+1);  S: -S^2/(1+C) (falls back to C-1 near a = pi)
  const float sa2_c1 = (sa * sa) * f_rcp(ca + 1.f);
  const float cam1 = HYP ? sa2_c1 : (ca > -0.5f ? -sa2_c1 : ca - 1.f);
  float* sg = o.sigma;
  load_sigma<N>(n, l_n, l, sg);
  float v[CN];
  float Sv = 0.f, pm_ = 0.f, se2 = 0.f;
  PM_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      v[j] = e[j] * sg[j];
      Sv = fmaf(v[j], v[j], Sv);
      pm_ = fmaf(m[j], v[j], pm_);
      se2 = fmaf(e[j], e[j], se2);
    }
  const float p = pm_ * idn;  // <mh, v>
  const float slog = sum_log<N>(n, sg);
  // ---- sample ----
  float ln, t, it, ct, st, A;
  bool t_sat = false;
  if (HYP) {
    const float SvC = fmaxf(Sv, 1e-9f);  // sqrt_ clamp (common.py:117-119)
    const float irs = f_rsqrt(SvC);
    ln = SvC * irs;
    t = ln * iR;
    it = irs * R;
    t_sat = POI && t > 30.f;
    coshsinh_pos(POI ? fminf(t, 30.f) : t, &ct, &st);
    A = st * it;
  } else {
    ln = f_sqrt(Sv);
    t = ln * iR;
    sincos_cw(t, &st, &ct);
    it = t > 0.f ? f_rcp(t) : 0.f;
    A = t > 0.f ? st * it : 1.f;
  }
  const float Rct = R * ct, Ap = A * p;
  const float z0 = HYP ? fmaf(Ap, sa, Rct * ca) : fmaf(-Ap, sa, Rct * ca);
  const float Bc = fmaf(Rct, sa, Ap * cam1);
  const float Bm = Bc * idn;
  float* z = o.z;
  float zt[CN];
  float zt2 = 0.f;
  PM_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      zt[j] = fmaf(A, v[j], Bm * m[j]);
      zt2 = fmaf(zt[j], zt[j], zt2);
    }
  const float iRz = PROJ ? f_rcp(R + z0) : 0.f;
  const float pj = PROJ ? R * iRz : 1.f;  // lorentz_to_poincare | spherical_to_projected
  if (PROJ) {
    PM_UNROLL
    for (int j = 0; j < CN; ++j)
      if (j < n) z[j] = pj * zt[j];
  } else {
    z[0] = z0;
    PM_UNROLL
    for (int j = 0; j < CN; ++j)
      if (j < n) z[j + 1] = zt[j];
  }
  if (WANT_MS) {
    float* mu = o.mu;
    if (PROJ) {
      const float Ta = R * sa * f_rcp(ca + 1.f) * idn;  // tanh(a) = sinh(2a) / (cosh(2a) + 1), tan(a) likewise
      PM_UNROLL
      for (int j = 0; j < CN; ++j)
        if (j < n) mu[j] = Ta * m[j];
    } else {
      mu[0] = R * ca;
      const float Rs = R * sa * idn;
      PM_UNROLL
      for (int j = 0; j < CN; ++j)
        if (j < n) mu[j + 1] = Rs * m[j];
    }
  }
  // ---- prior distance r = dist(mu0, z)/R from the tail norm (well conditioned everywhere) ----
  const float R2 = R * R, iR2 = iR * iR;
  const float rl_min = kSqrtClamp * iR;
  const float s2 = zt2 * iR2;
  const float s = f_sqrt(s2);
  float r, rq, as_ = 0.f, alpha = 0.f, snr = 0.f, csr = 0.f, irl = 0.f, tw = 0.f, D;
  bool r_clamped = false, at_clamped = false;
  if (HYP) {
    as_ = f_sqrt(1.f + s2);
    r = f_log1p(fmaf(s2, f_rcp(1.f + as_), s));  // asinh(s)
    // H._logdet applies sqrt() (clamp 1e-9) to the squared Lorentz norm R^2 r^2 of the prior's tangent vector
    r_clamped = r < rl_min;
    const float rl = fmaxf(r, rl_min);
    irl = f_rcp(rl);
    D = log_sinhc(t, it) - log_sinhc(rl, irl);
    at_clamped = POI && r > kAtMax;
    rq = POI ? fminf(r, kAtMax) : r;
  } else {
    alpha = z0 * iR;
    r = atan2_pos(s, alpha);
    const float inv_q = f_rsqrt(fmaf(alpha, alpha, s2));  // (alpha, s) is a unit vector up to rounding
    snr = s * inv_q;                                       // sin r and cos r without going through r
    csr = alpha * inv_q;
    // G(x) = log clamp(|sin x|, 1e-5) - log clamp(x, 1e-5) (spherical.py:58-67); G(t) - G(r) under one logarithm
    // 'd' measures the posterior's log-det at the wrapped angle tw = acos(cos t) (|sin| is the same)
    tw = PSP ? atan2_pos(fabsf(st), ct) : t;
    const float num = fmaxf(fabsf(st), 1e-5f) * fmaxf(r, 1e-5f);
    const float den = fmaxf(fabsf(snr), 1e-5f) * fmaxf(tw, 1e-5f);
    D = f_lg2(num * f_rcp(den)) * kLn2;
    rq = r;
  }
  const float nm1 = (float)(n - 1);
  o.kl = fmaf(0.5f * R2, rq * rq, fmaf(-0.5f, se2, -slog)) - nm1 * D;
  if (!BWD) return;

  // ================================ reverse sweep ================================
  float gR = 0.f;
  // KL -> r
  float g_r = gkl * R2 * rq;
  gR += gkl * R * (rq * rq);
  if (at_clamped) {
    // geoopt Artanh.backward = g / (1 - x'^2) on the clamped argument x' = 1 - 1e-5; d(rho)/d(r) = sech^2(r/2) / 2
    const float eh = f_ex2(-kLog2e * fminf(r, 30.f));  // sech^2(r/2) = 4 e^-r / (1 + e^-r)^2
    const float dh = 1.f + eh;
    g_r *= 4.f * eh * f_rcp(dh * dh) * (1.f / (1e-5f * (2.f - 1e-5f)));
  }
  float g_s, g_alpha = 0.f;
  if (HYP) {
    if (!r_clamped) {
      g_r = fmaf(gkl * nm1, log_sinhc_d(r, irl), g_r);
    } else {
      const float rl = rl_min;
      gR += -(gkl * nm1 * log_sinhc_d(rl, irl)) * rl * iR;  // leaky clamp: the 1e-8 * g path into r is dropped
    }
    g_s = g_r * f_rcp(as_);
  } else {
    // G'(x) = [|sin x| >= 1e-5] cos x / sin x - [x >= 1e-5] / x
    float dG = 0.f;
    if (fabsf(snr) >= 1e-5f) dG = csr * f_rcp(snr);
    if (r >= 1e-5f) dG -= f_rcp(r);
    g_r = fmaf(gkl * nm1, dG, g_r);
    // r = atan2(s, alpha): any smooth extension off the constraint alpha^2 + s^2 = 1 has the same total derivative
    const float iq2 = f_rcp(fmaf(alpha, alpha, s2));
    g_s = g_r * alpha * iq2;
    g_alpha = -g_r * s * iq2;
  }
  // s = |z_tail| / R ; alpha = z0 / R
  const float k_zt = s > 0.f ? g_s * iR2 * f_rcp(s) : 0.f;
  gR += -g_s * s * iR;
  float g_z0 = 0.f;
  if (PROJ) {
    // z_j = R Z_j / (R + Z_0)
    PM_UNROLL
    for (int j = 0; j < CN; ++j)
      if (j < n) {
        const float gzz = gz[j] * z[j];
        g_z0 = fmaf(-gzz, iRz, g_z0);
        gR = fmaf(gzz, iR - iRz, gR);
      }
  } else {
    g_z0 = gz[0];
  }
  if (!HYP) {
    g_z0 = fmaf(g_alpha, iR, g_z0);
    gR += -g_alpha * alpha * iR;
  }
  // z0 = R ct ca + sgn A sa p ;  Bc = R ct sa + A cam1 p ; z_tail = A v + Bc mh
  float g_A = 0.f, g_Bm = 0.f;
  float G[CN];
  PM_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      G[j] = fmaf(k_zt, zt[j], PROJ ? pj * gz[j] : gz[j + 1]);
      g_A = fmaf(G[j], v[j], g_A);
      g_Bm = fmaf(G[j], m[j], g_Bm);
    }
  const float g_Bc = g_Bm * idn;  // <G, mh>
  const float sA = HYP ? A : -A;  // sgn * A
  const float g_ct = R * fmaf(g_z0, ca, g_Bc * sa);
  const float g_ca = fmaf(g_z0, Rct, g_Bc * Ap);  // d(cam1)/d(ca) = 1
  const float g_sa = fmaf(g_z0, sA * p, g_Bc * Rct);
  const float g_p = fmaf(g_z0 * sA, sa, g_Bc * A * cam1);
  gR += ct * fmaf(g_z0, ca, g_Bc * sa);
  g_A += p * fmaf(g_z0, HYP ? sa : -sa, g_Bc * cam1);
  // A = st / t ; ct, st functions of t ; KL has -(n-1) Fq(t)
  float g_t = 0.f;
  if (t > 0.f) {
    const float g_st = g_A * it;
    g_t = -g_A * A * it;
    if (HYP) {
      const float dclamp = POI ? (t_sat ? 0.f : 1.f) : (t <= kMaxHyp ? 1.f : 1e-8f);
      g_t = fmaf(fmaf(g_ct, st, g_st * ct), dclamp, g_t);
      g_t = fmaf(-gkl * nm1, log_sinhc_d(t, it), g_t);
    } else {
      g_t += fmaf(-g_ct, st, g_st * ct);
      float dG = 0.f;
      if (fabsf(st) >= 1e-5f) dG = ct * f_rcp(st);
      if (PSP) {
        if (tw >= 1e-5f) dG -= (st >= 0.f ? 1.f : -1.f) * f_rcp(tw);  // d(tw)/dt = sign(sin t)
      } else if (t >= 1e-5f) {
        dG -= it;
      }
      g_t = fmaf(-gkl * nm1, dG, g_t);
    }
  }
  // t = ln / R ; ln = sqrt(Sv) ; Sv = <v, v>
  gR += -g_t * t * iR;
  const float g_ln = g_t * iR;
  float g_Sv;
  if (HYP) g_Sv = g_ln * (Sv >= 1e-9f ? 1.f : 1e-8f) * (0.5f * f_rcp(ln));  // leaky sqrt clamp
  else g_Sv = ln > 0.f ? g_ln * (0.5f * f_rcp(ln)) : 0.f;
  // a : ca, sa
  float g_a;
  if (POI) g_a = a_sat ? 0.f : 2.f * fmaf(g_ca, sa, g_sa * ca);
  else if (HYP) g_a = fmaf(g_ca, sa, g_sa * ca) * (a <= kMaxHyp ? 1.f : 1e-8f);
  else g_a = (PSP ? 2.f : 1.f) * fmaf(-g_ca, sa, g_sa * ca);
  gR += -g_a * a * iR;
  float g_nm = PROJ ? 0.f : g_a * iR;   // P, D: a = max(|m|, MIN_NORM) / R
  float g_dn = PROJ ? g_a * iR : 0.f;
  // v = eps * sigma ; p = <mh, v> ; mh = m / dn ; z_tail = A v + Bc mh
  float g_s_[CN];
  const float g_pm = g_p * idn, twoSv = 2.f * g_Sv;
  float gmh_m = 0.f;  // <g_mh, m>
  PM_UNROLL
  for (int j = 0; j < CN; ++j)
    if (j < n) {
      const float gv = fmaf(A, G[j], fmaf(twoSv, v[j], g_pm * m[j]));
      g_s_[j] = fmaf(gv, e[j], -gkl * f_rcp(sg[j]));
      const float gmh = fmaf(Bc, G[j], g_p * v[j]);
      gm[j] = gmh * idn;
      gmh_m = fmaf(gmh, m[j], gmh_m);
    }
  g_dn += -gmh_m * idn * idn;  // d(mh_j)/d(dn) = -m_j / dn^2
  store_gl<N>(n, l_n, l, g_s_, gl);
  if (nm >= nmin) g_nm += g_dn;
  if (nm > 0.f) {
    const float k = g_nm * f_rcp(nm);
    PM_UNROLL
    for (int j = 0; j < CN; ++j)
      if (j < n) gm[j] = fmaf(k, m[j], gm[j]);
  }
  *gR_out = gR;
}

}  // namespace pm
}  // namespace mvae
