// manifold_ops.cu — standalone per-sample vector ops of the reference's `Manifold` interface
// (mt/mvae/ops/manifold.py:22-60) and the Wrapped-Normal pieces (mt/mvae/distributions/wrapped_normal.py:70-103)
// for ONE component, behind mvae_manifold_op / mvae_wn_*.  These are not on the train_step path (the fused
// product-manifold kernel is); they keep the Manifold / WrappedNormal API complete on the device.  One thread per
// sample, runtime dimension (rolled loops); rows are short (d <= 161 floats) so accesses stay within a few sectors.
#include <math.h>

#include "manifold_math.cuh"

namespace mvae {

constexpr int kOpMaxD = kDynMaxN + 1;

struct OpParams {
  int op, manifold, n, d;
  int64_t B;
  const float* x;
  const float* y;
  const float* radius;
  float* out;
  int x_w, y_w, out_w;  // row widths
};

__device__ __forceinline__ float lorentz(int d, const float* a, const float* b) {
  float s = 0.f;
  for (int k = 0; k < d; ++k) s += a[k] * b[k];
  return s - 2.f * (a[0] * b[0]);
}
__device__ __forceinline__ float dot(int d, const float* a, const float* b) {
  float s = 0.f;
  for (int k = 0; k < d; ++k) s += a[k] * b[k];
  return s;
}
// common.py:46-63 atanh with the leaky clamp at +-(1 - 4e-8)
__device__ __forceinline__ float atanh_g(float x) {
  const float hi = (float)(1.0 - 4e-8);
  float xc = lclamp(x, -hi, hi);
  return (logf(1.f + xc) - logf(1.f - xc)) * 0.5f;
}
__device__ __forceinline__ void mobius_add_rt(int n, const float* x, const float* y, float c, float* out,
                                              float xsign) {
  float x2 = 0.f, y2 = 0.f, xy = 0.f;
  for (int j = 0; j < n; ++j) {
    float xj = xsign * x[j];
    x2 += xj * xj;
    y2 += y[j] * y[j];
    xy += xj * y[j];
  }
  float A = 1.f + 2.f * c * xy + c * y2;
  float Bc = 1.f - c * x2;
  float den = fmaxf(1.f + 2.f * c * xy + c * c * x2 * y2, kPMin);
  for (int j = 0; j < n; ++j) out[j] = (A * (xsign * x[j]) + Bc * y[j]) / den;
}
__device__ __forceinline__ float lambda_x(int n, const float* x, float c) {
  float s = 0.f;
  for (int j = 0; j < n; ++j) s += x[j] * x[j];
  return 2.f / fmaxf(1.f - c * s, kPMin);
}
// geoopt expmap(x, u): x (+) tanh(sqrt_c/2 * lambda_x * |u|) u / (sqrt_c |u|)
__device__ __forceinline__ void p_expmap(int n, const float* at, const float* u, float c, float* out) {
  float sc = powf(c, 0.5f);
  float un = fmaxf(sqrtf(dot(n, u, u)), kPMin);
  float k = tanh_c(sc / 2.f * lambda_x(n, at, c) * un);
  float sec[kOpMaxD];
  for (int j = 0; j < n; ++j) sec[j] = k * u[j] / (sc * un);
  mobius_add_rt(n, at, sec, c, out, 1.f);
}
// geoopt logmap(x, y)
__device__ __forceinline__ void p_logmap(int n, const float* at, const float* y, float c, float* out) {
  float sub[kOpMaxD];
  mobius_add_rt(n, at, y, c, sub, -1.f);
  float sn = fmaxf(sqrtf(dot(n, sub, sub)), kPMin);
  float lam = lambda_x(n, at, c);
  float sc = powf(c, 0.5f);
  float xc;
  float k = 2.f / sc / lam * artanh_go(sc * sn, &xc);
  for (int j = 0; j < n; ++j) out[j] = k * sub[j] / sn;
}
__device__ __forceinline__ void h_inv_exp_rt(int d, const float* x, const float* at, float R, float* w) {
  float alpha = -lorentz(d, at, x) / (R * R);
  float zz;
  float coef = acosh_g(alpha, &zz) / sqrt_g(alpha * alpha - 1.f);
  for (int k = 0; k < d; ++k) w[k] = coef * (x[k] - alpha * at[k]);
}
__device__ __forceinline__ void s_inv_exp_rt(int d, const float* x, const float* at, float R, float* w) {
  float alpha = dot(d, at, x) / (R * R);
  float coef = acosf(fminf(fmaxf(alpha, -1.f), 1.f)) / sqrt_g(1.f - alpha * alpha);
  for (int k = 0; k < d; ++k) w[k] = coef * (x[k] - alpha * at[k]);
}
__device__ __forceinline__ void p2l_rt(int n, const float* y, float R, float* out) {
  float nrm = sqrtf(dot(n, y, y));
  float nn = nrm * nrm;
  float den = R * R - nn;
  out[0] = R * (R * R + nn) / den;
  for (int j = 0; j < n; ++j) out[j + 1] = 2.f * (R * R) * y[j] / den;
}
__device__ __forceinline__ float h_logdet_rt(int d, const float* u, float R) {
  return h_logdet_pr<false>(d - 1, lorentz(d, u, u), R, 0.f, nullptr, nullptr);
}
__device__ __forceinline__ float s_logdet_rt(int d, const float* u, float R) {
  return s_logdet_nu<false>(d - 1, sqrtf(dot(d, u, u)), R, 0.f, nullptr, nullptr);
}
// PoincareBall.logdet (poincare.py:84-89)
__device__ __forceinline__ float p_logdet_rt(int n, const float* z, const float* mu, float R) {
  float zs[kOpMaxD], ms[kOpMaxD], uu[kOpMaxD];
  p2l_rt(n, z, R, zs);
  p2l_rt(n, mu, R, ms);
  h_inv_exp_rt(n + 1, zs, ms, R, uu);
  return h_logdet_rt(n + 1, uu, R);
}

// ---- stereographically projected sphere 'd' (spherical_projected.py; geoopt mobius_add at c = -1/R^2, :107-113) ----
__device__ __forceinline__ float d_lambda(int n, const float* x, float c) {  // lambda_x_c :123-124
  return 2.f / fmaxf(1.f + c * dot(n, x, x), kPMin);
}
__device__ __forceinline__ void d_expmap(int n, const float* at, const float* x, float R, float* out) {  // :141-147
  const float c = 1.f / (R * R);
  const float r = fmaxf(sqrtf(dot(n, x, x)), kPMin) / R;
  const float k = tanf(r * d_lambda(n, at, c) / 2.f);
  float rhs[kOpMaxD];
  for (int j = 0; j < n; ++j) rhs[j] = k * x[j] / r;
  mobius_add_rt(n, at, rhs, -c, out, 1.f);
}
__device__ __forceinline__ void d_logmap(int n, const float* at, const float* x, float R, float* out) {  // :157-162
  const float c = 1.f / (R * R);
  float sub[kOpMaxD];
  mobius_add_rt(n, at, x, -c, sub, -1.f);
  const float nm = fmaxf(sqrtf(dot(n, sub, sub)), kPMin) / R;
  const float k = 2.f / d_lambda(n, at, c) * atanf(nm);
  for (int j = 0; j < n; ++j) out[j] = k * (sub[j] / nm);
}
__device__ __forceinline__ void d2s_rt(int n, const float* y, float R, float* out) {  // projected_to_spherical :190-195
  const float nrm = sqrtf(dot(n, y, y));
  const float yn2 = nrm * nrm, r2 = R * R;
  out[0] = R * (r2 - yn2) / (yn2 + r2);
  for (int j = 0; j < n; ++j) out[j + 1] = 2.f * r2 * y[j] / (yn2 + r2);
}
// StereographicallyProjectedSphere.logdet (:58-92)
__device__ __forceinline__ float d_logdet_rt(int n, const float* z, const float* mu, float R) {
  float zs[kOpMaxD], ms[kOpMaxD], uu[kOpMaxD];
  d2s_rt(n, z, R, zs);
  d2s_rt(n, mu, R, ms);
  s_inv_exp_rt(n + 1, zs, ms, R, uu);
  return s_logdet_rt(n + 1, uu, R);
}

// sample_projection_mu0(v, at) -> z, u   (hyperbolics.py:138-142, spherical.py:119-123, poincare.py:152-157, euclidean.py:90-93)
__device__ __forceinline__ void sample_projection(int man, int n, const float* v, const float* at, float R, float* z,
                                                  float* u) {
  const int d = (man == MVAE_HYPERBOLOID || man == MVAE_SPHERE) ? n + 1 : n;
  if (man == MVAE_HYPERBOLOID || man == MVAE_SPHERE) {
    const bool hyp = man == MVAE_HYPERBOLOID;
    float lp = 0.f;
    for (int j = 0; j < n; ++j) lp += at[j + 1] * v[j];
    float coef = lp / (R * (R + at[0]));
    if (!hyp) coef = -coef;
    u[0] = hyp ? coef * (at[0] + R) : 0.f + coef * (at[0] + R);
    for (int j = 0; j < n; ++j) u[j + 1] = v[j] + coef * at[j + 1];
    float t, c1, s1;
    if (hyp) {
      t = sqrt_g(lorentz(d, u, u)) / R;
      coshsinh_g(t, &c1, &s1);
    } else {
      t = sqrtf(dot(d, u, u)) / R;
      c1 = cosf(t);
      s1 = sinf(t);
    }
    for (int k = 0; k < d; ++k) z[k] = c1 * at[k] + s1 * (u[k] / t);
  } else if (man == MVAE_POINCARE) {
    float c = 1.f / (R * R);
    float lam = lambda_x(n, at, c);
    for (int j = 0; j < n; ++j) u[j] = v[j] / lam;
    p_expmap(n, at, u, c, z);
  } else if (man == MVAE_PROJ_SPHERE) {  // spherical_projected.py:176-179
    float lam = d_lambda(n, at, 1.f / (R * R));
    for (int j = 0; j < n; ++j) u[j] = v[j] / lam;
    d_expmap(n, at, u, R, z);
  } else {
    for (int j = 0; j < n; ++j) {
      u[j] = v[j];
      z[j] = at[j] + v[j] / 2.f;
    }
  }
}

// inverse_sample_projection_mu0(z, at) -> u, v
__device__ __forceinline__ void inv_sample_projection(int man, int n, const float* z, const float* at, float R,
                                                      float* u, float* v) {
  const int d = (man == MVAE_HYPERBOLOID || man == MVAE_SPHERE) ? n + 1 : n;
  if (man == MVAE_HYPERBOLOID) {
    h_inv_exp_rt(d, z, at, R, u);
    float coef = -u[0] / (R + at[0]);
    for (int j = 0; j < n; ++j) v[j] = u[j + 1] + coef * at[j + 1];
  } else if (man == MVAE_SPHERE) {
    s_inv_exp_rt(d, z, at, R, u);
    float coef = u[0] / (R + at[0]);
    for (int j = 0; j < n; ++j) v[j] = u[j + 1] - coef * at[j + 1];
  } else if (man == MVAE_POINCARE) {
    float c = 1.f / (R * R);
    p_logmap(n, at, z, c, u);
    float lam = lambda_x(n, at, c);
    for (int j = 0; j < n; ++j) v[j] = u[j] * lam;
  } else if (man == MVAE_PROJ_SPHERE) {  // spherical_projected.py:182-186
    d_logmap(n, at, z, R, u);
    float lam = d_lambda(n, at, 1.f / (R * R));
    for (int j = 0; j < n; ++j) v[j] = u[j] * lam;
  } else {
    for (int j = 0; j < n; ++j) {
      u[j] = 2.f * (z[j] - at[j]);
      v[j] = u[j];
    }
  }
}

__device__ __forceinline__ float normal_lp(int n, const float* v, const float* sg) {
  float acc = 0.f;
  for (int j = 0; j < n; ++j) acc += -(v[j] * v[j]) / (2.f * (sg[j] * sg[j])) - logf(sg[j]) - kHalfLn2Pi;
  return acc;
}

__global__ void __launch_bounds__(128) manifold_op_kernel(const OpParams p) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  const int n = p.n, d = p.d, man = p.manifold;
  const bool amb = man == MVAE_HYPERBOLOID || man == MVAE_SPHERE;
  const float R = (p.radius && man != MVAE_EUCLIDEAN) ? radius_of(__ldg(p.radius)) : 1.f;
  float x[kOpMaxD], y[kOpMaxD], o[kOpMaxD];
  for (int k = 0; k < p.x_w; ++k) x[k] = p.x[b * p.x_w + k];
  if (p.y)
    for (int k = 0; k < p.y_w; ++k) y[k] = p.y[b * p.y_w + k];
  const float c = 1.f / (R * R);
  switch (p.op) {
    case MVAE_OP_EXP_MAP_MU0: {
      if (amb) {
        float nm = sqrtf(dot(n, x, x));
        float a = nm / R, dn = fmaxf(nm, 1e-12f), c1, s1;
        if (man == MVAE_HYPERBOLOID) coshsinh_g(a, &c1, &s1);
        else { c1 = cosf(a); s1 = sinf(a); }
        o[0] = c1 * R;
        for (int j = 0; j < n; ++j) o[j + 1] = s1 * ((x[j] / dn) * R);
      } else if (man == MVAE_POINCARE) {
        float sc = powf(c, 0.5f);
        float un = fmaxf(sqrtf(dot(n, x, x)), kPMin);
        float th = tanh_c(sc * un);
        for (int j = 0; j < n; ++j) o[j] = th * x[j] / (sc * un);
      } else if (man == MVAE_PROJ_SPHERE) {  // spherical_projected.py:150-154
        float r = fmaxf(sqrtf(dot(n, x, x)), kPMin) / R;
        float tn = tanf(r);
        for (int j = 0; j < n; ++j) o[j] = tn * x[j] / r;
      } else {
        for (int j = 0; j < n; ++j) o[j] = x[j] / 2.f;
      }
    } break;
    case MVAE_OP_INV_EXP_MAP_MU0: {
      if (amb) {
        float alpha = x[0] / R;
        float coef;
        if (man == MVAE_HYPERBOLOID) {
          float zz;
          coef = acosh_g(alpha, &zz) / sqrt_g(alpha * alpha - 1.f);
        } else {
          coef = acosf(fminf(fmaxf(alpha, -1.f), 1.f)) / sqrt_g(1.f - alpha * alpha);
        }
        o[0] = coef * (x[0] - alpha * R);
        for (int k = 1; k < d; ++k) o[k] = coef * x[k];
      } else if (man == MVAE_POINCARE) {
        float sc = powf(c, 0.5f);
        float yn = fmaxf(sqrtf(dot(n, x, x)), kPMin);
        float xc;
        float at = artanh_go(sc * yn, &xc);
        for (int j = 0; j < n; ++j) o[j] = x[j] / yn / sc * at;
      } else if (man == MVAE_PROJ_SPHERE) {  // spherical_projected.py:165-168
        float nx = fmaxf(sqrtf(dot(n, x, x)), kPMin) / R;
        float at = atanf(nx);
        for (int j = 0; j < n; ++j) o[j] = at * (x[j] / nx);
      } else {
        for (int j = 0; j < n; ++j) o[j] = 2.f * x[j];
      }
    } break;
    case MVAE_OP_EXP_MAP: {
      if (amb) {
        float t, c1, s1;
        if (man == MVAE_HYPERBOLOID) {
          t = sqrt_g(lorentz(d, x, x)) / R;
          coshsinh_g(t, &c1, &s1);
        } else {
          t = sqrtf(dot(d, x, x)) / R;
          c1 = cosf(t);
          s1 = sinf(t);
        }
        for (int k = 0; k < d; ++k) o[k] = c1 * y[k] + s1 * (x[k] / t);
      } else if (man == MVAE_POINCARE) {
        p_expmap(n, y, x, c, o);
      } else if (man == MVAE_PROJ_SPHERE) {
        d_expmap(n, y, x, R, o);
      } else {
        for (int j = 0; j < n; ++j) o[j] = y[j] + x[j] / 2.f;
      }
    } break;
    case MVAE_OP_INV_EXP_MAP: {
      if (man == MVAE_HYPERBOLOID) h_inv_exp_rt(d, x, y, R, o);
      else if (man == MVAE_SPHERE) s_inv_exp_rt(d, x, y, R, o);
      else if (man == MVAE_POINCARE) p_logmap(n, y, x, c, o);
      else if (man == MVAE_PROJ_SPHERE) d_logmap(n, y, x, R, o);
      else
        for (int j = 0; j < n; ++j) o[j] = 2.f * (x[j] - y[j]);
    } break;
    case MVAE_OP_PT_MU0: {
      if (man == MVAE_HYPERBOLOID) {
        float coef = lorentz(d, y, x) / (R * (R + y[0]));
        o[0] = x[0] + coef * (y[0] + R);
        for (int k = 1; k < d; ++k) o[k] = x[k] + coef * y[k];
      } else if (man == MVAE_SPHERE) {
        float coef = dot(d, y, x) / (R * (R + y[0]));
        o[0] = x[0] - coef * (y[0] + R);
        for (int k = 1; k < d; ++k) o[k] = x[k] - coef * y[k];
      } else if (man == MVAE_POINCARE) {
        float f = fmaxf(1.f - c * dot(n, y, y), kPMin);
        for (int j = 0; j < n; ++j) o[j] = x[j] * f;
      } else if (man == MVAE_PROJ_SPHERE) {  // (2 / lambda_dst) x, spherical_projected.py:133-134
        float f = 2.f / d_lambda(n, y, c);
        for (int j = 0; j < n; ++j) o[j] = f * x[j];
      } else {
        for (int j = 0; j < n; ++j) o[j] = x[j];
      }
    } break;
    case MVAE_OP_INV_PT_MU0: {
      if (man == MVAE_HYPERBOLOID) {
        float coef = -x[0] / (R + y[0]);
        o[0] = x[0] + coef * (y[0] + R);
        for (int k = 1; k < d; ++k) o[k] = x[k] + coef * y[k];
      } else if (man == MVAE_SPHERE) {
        float coef = x[0] / (R + y[0]);
        o[0] = x[0] - coef * (y[0] + R);
        for (int k = 1; k < d; ++k) o[k] = x[k] - coef * y[k];
      } else if (man == MVAE_POINCARE) {
        float f = fmaxf(1.f - c * dot(n, y, y), kPMin);
        for (int j = 0; j < n; ++j) o[j] = x[j] / f;
      } else if (man == MVAE_PROJ_SPHERE) {  // (lambda_src / 2) x, spherical_projected.py:137-138
        float f = d_lambda(n, y, c) / 2.f;
        for (int j = 0; j < n; ++j) o[j] = f * x[j];
      } else {
        for (int j = 0; j < n; ++j) o[j] = x[j];
      }
    } break;
    case MVAE_OP_DISTANCE: {
      if (man == MVAE_HYPERBOLOID) {
        float zz;
        o[0] = R * acosh_g(-lorentz(d, x, y) / (R * R), &zz);
      } else if (man == MVAE_SPHERE) {
        o[0] = R * acosf(fminf(fmaxf(dot(d, x, y) / (R * R), -1.f), 1.f));
      } else if (man == MVAE_POINCARE) {
        float sc = sqrt_g(c);
        float sub[kOpMaxD];
        mobius_add_rt(n, x, y, c, sub, -1.f);
        o[0] = atanh_g(sc * sqrtf(dot(n, sub, sub))) * 2.f / sc;
      } else if (man == MVAE_PROJ_SPHERE) {  // spherical_projected_distance, spherical_projected.py:95-102 (K = c)
        float dd = 0.f;
        for (int j = 0; j < n; ++j) dd += (x[j] - y[j]) * (x[j] - y[j]);
        float arg = 1.f - 2.f * c * dd / ((1.f + c * dot(n, x, x)) * (1.f + c * dot(n, y, y)));
        o[0] = 1.f / sqrt_g(c) * acosf(fminf(arg, 1.f));
      } else {
        float s = 0.f;
        for (int j = 0; j < n; ++j) s += (x[j] - y[j]) * (x[j] - y[j]);
        o[0] = 2.f * sqrtf(s);
      }
    } break;
    case MVAE_OP_MOBIUS_ADD: mobius_add_rt(n, x, y, man == MVAE_PROJ_SPHERE ? -c : c, o, 1.f); break;
    case MVAE_OP_MOBIUS_SCALAR_MUL: {
      // r (x)_c x = tanh(r artanh(sqrt_c |x|)) x / (sqrt_c |x|)   (no reference call site: parity unpinned)
      float sc = powf(c, 0.5f);
      float xn = fmaxf(sqrtf(dot(n, x, x)), kPMin);
      float xc;
      float k = tanh_c(y[0] * artanh_go(sc * xn, &xc));
      for (int j = 0; j < n; ++j) o[j] = k * x[j] / (xn * sc);
    } break;
    case MVAE_OP_LOGDET:
      o[0] = man == MVAE_HYPERBOLOID ? h_logdet_rt(d, x, R)
             : man == MVAE_SPHERE    ? s_logdet_rt(d, x, R)
                                     : d_logdet_rt(n, x, y, R);  // 'd': x = z, y = mu
      break;
    case MVAE_OP_TO_POINCARE:
      for (int j = 0; j < n; ++j) o[j] = R * x[j + 1] / (R + x[0]);
      break;
    default:  // MVAE_OP_FROM_POINCARE
      if (man == MVAE_PROJ_SPHERE) d2s_rt(n, x, R, o);
      else p2l_rt(n, x, R, o);
      break;
  }
  for (int k = 0; k < p.out_w; ++k) p.out[b * p.out_w + k] = o[k];
}

struct WnParams {
  int mode;  // 0 rsample, 1 log_prob_from_parts, 2 log_prob
  int manifold, n, d;
  int64_t B;
  const float* loc;
  const float* scale;
  const float* eps;
  const float* z_in;
  const float* u_in;
  const float* v_in;
  const float* radius;
  float* z;
  float* u;
  float* v;
  float* logp;
};

__global__ void __launch_bounds__(128) wn_kernel(const WnParams p) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  const int n = p.n, d = p.d, man = p.manifold;
  const float R = (p.radius && man != MVAE_EUCLIDEAN) ? radius_of(__ldg(p.radius)) : 1.f;
  float loc[kOpMaxD], sg[kOpMaxD], z[kOpMaxD], u[kOpMaxD], v[kOpMaxD];
  for (int k = 0; k < d; ++k) loc[k] = p.loc[b * d + k];
  for (int j = 0; j < n; ++j) sg[j] = p.scale[b * n + j];
  if (p.mode == 0) {
    for (int j = 0; j < n; ++j) v[j] = p.eps[b * n + j] * sg[j];
    sample_projection(man, n, v, loc, R, z, u);
    for (int k = 0; k < d; ++k) {
      p.z[b * d + k] = z[k];
      if (p.u) p.u[b * d + k] = u[k];
    }
    if (p.v)
      for (int j = 0; j < n; ++j) p.v[b * n + j] = v[j];
    return;
  }
  for (int k = 0; k < d; ++k) z[k] = p.z_in[b * d + k];
  if (p.mode == 1) {
    for (int k = 0; k < d; ++k) u[k] = p.u_in[b * d + k];
    for (int j = 0; j < n; ++j) v[j] = p.v_in[b * n + j];
  } else {
    inv_sample_projection(man, n, z, loc, R, u, v);
  }
  float ld = 0.f;
  if (man == MVAE_HYPERBOLOID) ld = h_logdet_rt(d, u, R);
  else if (man == MVAE_SPHERE) ld = s_logdet_rt(d, u, R);
  else if (man == MVAE_POINCARE) ld = p_logdet_rt(n, z, loc, R);
  else if (man == MVAE_PROJ_SPHERE) ld = d_logdet_rt(n, z, loc, R);
  p.logp[b] = normal_lp(n, v, sg) - ld;
}

static int amb_dim(int man, int n) { return (man == MVAE_HYPERBOLOID || man == MVAE_SPHERE) ? n + 1 : n; }

static int check_manifold(int man, int n) {
  if (man < MVAE_EUCLIDEAN || man > MVAE_PROJ_SPHERE || n < 1) return MVAE_ERR_INVALID_ARGUMENT;
  if (n > kDynMaxN) return MVAE_ERR_UNSUPPORTED;
  return MVAE_OK;
}

}  // namespace mvae

using namespace mvae;

extern "C" int mvae_manifold_op(int32_t op, int32_t manifold, int32_t n, int64_t B, const float* x, const float* y,
                                const float* radius, float* out, void* stream) {
  int rc = check_manifold(manifold, n);
  if (rc != MVAE_OK) return rc;
  if (op < MVAE_OP_EXP_MAP_MU0 || op > MVAE_OP_FROM_POINCARE || B < 0) return MVAE_ERR_INVALID_ARGUMENT;
  if (B == 0) return MVAE_OK;
  if (!x || !out) return MVAE_ERR_INVALID_ARGUMENT;
  const int d = amb_dim(manifold, n);
  const bool amb = d != n;
  OpParams p = {};
  p.op = op;
  p.manifold = manifold;
  p.n = n;
  p.d = d;
  p.B = B;
  p.x = x;
  p.y = y;
  p.radius = radius;
  p.out = out;
  bool need_y = false;
  switch (op) {
    case MVAE_OP_EXP_MAP_MU0: p.x_w = n; p.out_w = d; break;
    case MVAE_OP_INV_EXP_MAP_MU0: p.x_w = d; p.out_w = d; break;
    case MVAE_OP_EXP_MAP:
    case MVAE_OP_INV_EXP_MAP:
    case MVAE_OP_PT_MU0:
    case MVAE_OP_INV_PT_MU0: p.x_w = d; p.y_w = d; p.out_w = d; need_y = true; break;
    case MVAE_OP_DISTANCE: p.x_w = d; p.y_w = d; p.out_w = 1; need_y = true; break;
    case MVAE_OP_MOBIUS_ADD:
      if (manifold != MVAE_POINCARE && manifold != MVAE_PROJ_SPHERE) return MVAE_ERR_UNSUPPORTED;
      p.x_w = d; p.y_w = d; p.out_w = d; need_y = true;
      break;
    case MVAE_OP_MOBIUS_SCALAR_MUL:
      if (manifold != MVAE_POINCARE) return MVAE_ERR_UNSUPPORTED;
      p.x_w = d; p.y_w = 1; p.out_w = d; need_y = true;
      break;
    case MVAE_OP_LOGDET:
      if (manifold == MVAE_PROJ_SPHERE) {  // logdet(mu, z) through the sphere: x = z, y = mu
        p.x_w = d; p.y_w = d; p.out_w = 1; need_y = true;
        break;
      }
      if (!amb) return MVAE_ERR_UNSUPPORTED;
      p.x_w = d; p.out_w = 1;
      break;
    case MVAE_OP_TO_POINCARE:
      if (!amb) return MVAE_ERR_UNSUPPORTED;
      p.x_w = d; p.out_w = n;
      break;
    default:
      if (manifold != MVAE_POINCARE && manifold != MVAE_PROJ_SPHERE) return MVAE_ERR_UNSUPPORTED;
      p.x_w = n; p.out_w = n + 1;
      break;
  }
  if (need_y && !y) return MVAE_ERR_INVALID_ARGUMENT;
  if (!need_y) p.y = nullptr;
  const int64_t blocks = (B + 127) / 128;
  if (blocks > 0x7fffffff) return MVAE_ERR_UNSUPPORTED;
  manifold_op_kernel<<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(p);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

static int launch_wn(WnParams& p, void* stream) {
  const int64_t blocks = (p.B + 127) / 128;
  if (blocks > 0x7fffffff) return MVAE_ERR_UNSUPPORTED;
  wn_kernel<<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(p);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

extern "C" int mvae_wn_rsample(int32_t manifold, int32_t n, int64_t B, const float* loc, const float* scale,
                               const float* eps, const float* radius, float* z, float* u, float* v, void* stream) {
  int rc = check_manifold(manifold, n);
  if (rc != MVAE_OK) return rc;
  if (B < 0) return MVAE_ERR_INVALID_ARGUMENT;
  if (B == 0) return MVAE_OK;
  if (!loc || !scale || !eps || !z) return MVAE_ERR_INVALID_ARGUMENT;
  WnParams p = {};
  p.mode = 0; p.manifold = manifold; p.n = n; p.d = amb_dim(manifold, n); p.B = B;
  p.loc = loc; p.scale = scale; p.eps = eps; p.radius = radius; p.z = z; p.u = u; p.v = v;
  return launch_wn(p, stream);
}

extern "C" int mvae_wn_log_prob_from_parts(int32_t manifold, int32_t n, int64_t B, const float* loc,
                                           const float* scale, const float* z, const float* u, const float* v,
                                           const float* radius, float* logp, void* stream) {
  int rc = check_manifold(manifold, n);
  if (rc != MVAE_OK) return rc;
  if (B < 0) return MVAE_ERR_INVALID_ARGUMENT;
  if (B == 0) return MVAE_OK;
  if (!loc || !scale || !z || !u || !v || !logp) return MVAE_ERR_INVALID_ARGUMENT;
  WnParams p = {};
  p.mode = 1; p.manifold = manifold; p.n = n; p.d = amb_dim(manifold, n); p.B = B;
  p.loc = loc; p.scale = scale; p.z_in = z; p.u_in = u; p.v_in = v; p.radius = radius; p.logp = logp;
  return launch_wn(p, stream);
}

extern "C" int mvae_wn_log_prob(int32_t manifold, int32_t n, int64_t B, const float* loc, const float* scale,
                                const float* z, const float* radius, float* logp, void* stream) {
  int rc = check_manifold(manifold, n);
  if (rc != MVAE_OK) return rc;
  if (B < 0) return MVAE_ERR_INVALID_ARGUMENT;
  if (B == 0) return MVAE_OK;
  if (!loc || !scale || !z || !logp) return MVAE_ERR_INVALID_ARGUMENT;
  WnParams p = {};
  p.mode = 2; p.manifold = manifold; p.n = n; p.d = amb_dim(manifold, n); p.B = B;
  p.loc = loc; p.scale = scale; p.z_in = z; p.radius = radius; p.logp = logp;
  return launch_wn(p, stream);
}
