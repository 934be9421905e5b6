/*
 * mvae_oracle.c — CPU oracle for the mvae hot path (TEST INFRASTRUCTURE, see mvae_oracle_impl.h).
 * Builds liboracle.so exporting oracle_*_f64 and oracle_*_f32 (same code, two arithmetic types).
 * Build: make -C oracle   (gcc -O2 -fopenmp; no fast-math so float64 results are reproducible).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mvae_b200.h"

#define MATHFN(sfx, T, name, cname) static inline T name##sfx(T x) { return cname(x); }

/* ---- float64 ---- */
MATHFN(_f64, double, msqrt, sqrt)
MATHFN(_f64, double, mcosh, cosh)
MATHFN(_f64, double, msinh, sinh)
MATHFN(_f64, double, mtanh, tanh)
MATHFN(_f64, double, mlog, log)
MATHFN(_f64, double, mlog1p, log1p)
MATHFN(_f64, double, mexp, exp)
MATHFN(_f64, double, msin, sin)
MATHFN(_f64, double, mcos, cos)
MATHFN(_f64, double, macos, acos)
MATHFN(_f64, double, mtan, tan)
MATHFN(_f64, double, matan, atan)
static inline double mpow_f64(double x, double y) { return pow(x, y); }
#define real double
#define SFX(n) n##_f64
#include "mvae_oracle_impl.h"
#undef real
#undef SFX

/* ---- float32 ---- */
MATHFN(_f32, float, msqrt, sqrtf)
MATHFN(_f32, float, mcosh, coshf)
MATHFN(_f32, float, msinh, sinhf)
MATHFN(_f32, float, mtanh, tanhf)
MATHFN(_f32, float, mlog, logf)
MATHFN(_f32, float, mlog1p, log1pf)
MATHFN(_f32, float, mexp, expf)
MATHFN(_f32, float, msin, sinf)
MATHFN(_f32, float, mcos, cosf)
MATHFN(_f32, float, macos, acosf)
MATHFN(_f32, float, mtan, tanf)
MATHFN(_f32, float, matan, atanf)
static inline float mpow_f32(float x, float y) { return powf(x, y); }
#define real float
#define SFX(n) n##_f32
#include "mvae_oracle_impl.h"
#undef real
#undef SFX

/* Packed descriptor layout shared with the product's mvae_pm_desc_init (include/mvae_b200.h):
 * ml = [m_0|l_0|m_1|l_1|...]; eps/sigma and z/mu concatenated in component order (vae.py:78). */
int oracle_pm_desc_init(mvae_pm_desc* D, int32_t C, const int32_t* types, const int32_t* dims, int32_t scalar) {
  if (!D || C < 1 || C > MVAE_MAX_COMPONENTS) return -1;
  memset(D, 0, sizeof(*D));
  int ml = 0, e = 0, z = 0;
  for (int i = 0; i < C; ++i) {
    mvae_component* c = &D->comp[i];
    if (dims[i] < 1) return -1;
    if (types[i] < MVAE_EUCLIDEAN || types[i] > MVAE_UNIVERSAL) return -1;
    c->type = types[i];
    c->n = dims[i];
    c->d = (types[i] == MVAE_HYPERBOLOID || types[i] == MVAE_SPHERE) ? dims[i] + 1 : dims[i];
    c->l_n = scalar ? 1 : dims[i];
    c->m_off = ml; ml += c->n;
    c->l_off = ml; ml += c->l_n;
    c->eps_off = e; e += c->n;
    c->z_off = z; z += c->d;
  }
  D->C = C; D->ld_ml = ml; D->ld_eps = e; D->ld_z = z;
  return 0;
}
int oracle_sizeof_pm_desc(void) { return (int)sizeof(mvae_pm_desc); }
