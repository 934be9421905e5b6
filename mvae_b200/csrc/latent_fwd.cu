// latent_fwd.cu — forward instantiations + C-ABI entry of the fused latent block (see latent_impl.cuh).
#define MVAE_LAT_BWD 0
#include "latent_impl.cuh"

using namespace mvae;

static bool lat_aligned(const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

extern "C" int mvae_latent_forward_ex(const mvae_pm_desc* desc, int64_t B, int32_t H, const float* h, int64_t ld_h,
                                      const float* Wh, const float* bh, float* eps, const float* radius,
                                      const float* Wd0, const float* bd0, float* ml, float* z, float* kl,
                                      const mvae_planes* dd_out, uint32_t* nonfinite_flag,
                                      const mvae_latent_prologue* pro, void* stream) {
  if (!desc || desc->C < 1 || desc->C > MVAE_MAX_COMPONENTS || B < 0 || H < 8 || !h || !Wh || !bh || !eps ||
      !Wd0 || !bd0 || !ml || !z || !kl || !dd_out || !dd_out->base)
    return MVAE_ERR_INVALID_ARGUMENT;
  if (desc->ld_ml > 64 || desc->ld_z > 64 || (H & 7)) return MVAE_ERR_UNSUPPORTED;
  if (dd_out->planes < 1 || dd_out->planes > 3 || ld_h < H || dd_out->ld < H || dd_out->rows < B)
    return MVAE_ERR_INVALID_ARGUMENT;
  if ((ld_h & 3) || ld_h > 0x7fffffff || (dd_out->ld & 1) || (dd_out->plane_stride & 1) || !lat_aligned(h, 16) ||
      !lat_aligned(dd_out->base, 4) || !lat_aligned(Wh, 16))
    return MVAE_ERR_ALIGNMENT;
  if (pro && (pro->n_zero < 0 || pro->n_zero > 4)) return MVAE_ERR_INVALID_ARGUMENT;
  if (B == 0) {
    // no rows, no launch: the zero fills still have to happen
    for (int s = 0; pro && s < pro->n_zero; ++s)
      if (pro->zero_n[s] > 0) {
        if (!pro->zero_ptr[s]) return MVAE_ERR_INVALID_ARGUMENT;
        MVAE_CUDA_TRY(cudaMemsetAsync(pro->zero_ptr[s], 0, (size_t)pro->zero_n[s] * 4, as_stream(stream)));
      }
    return MVAE_OK;
  }
  LatParams p;
  memset(&p, 0, sizeof(p));
  if (pro) {
    p.draw_eps = pro->draw_eps ? 1 : 0;
    p.seed = (unsigned long long)pro->seed;
    p.counter_dev = reinterpret_cast<const unsigned long long*>(pro->counter_dev);
    for (int s = 0; s < pro->n_zero; ++s) {
      if (pro->zero_n[s] < 0 || (pro->zero_n[s] > 0 && !pro->zero_ptr[s])) return MVAE_ERR_INVALID_ARGUMENT;
      if (pro->zero_n[s] == 0) continue;
      p.zptr[p.n_zero] = pro->zero_ptr[s];
      p.zn[p.n_zero] = pro->zero_n[s];
      ++p.n_zero;
    }
  }
  p.desc = *desc;
  p.B = B;
  p.H = H;
  p.h = h;
  p.h_ld = (int)ld_h;
  p.Wh = Wh;
  p.bh = bh;
  p.Wd0 = Wd0;
  p.bd0 = bd0;
  p.eps = eps;
  p.radius = radius;
  p.ml = ml;
  p.z = z;
  p.kl = kl;
  p.dd = dd_out->base;
  p.dd_stride = dd_out->planes > 1 ? dd_out->plane_stride : 0;
  p.dd_ld = dd_out->ld;
  p.dd_planes = dd_out->planes;
  p.flag = nonfinite_flag;
  return launch_latent(p, stream);
}

extern "C" int mvae_latent_forward(const mvae_pm_desc* desc, int64_t B, int32_t H, const float* h, int64_t ld_h, const float* Wh,
                                   const float* bh, const float* eps, const float* radius, const float* Wd0,
                                   const float* bd0, float* ml, float* z, float* kl, const mvae_planes* dd_out,
                                   uint32_t* nonfinite_flag, void* stream) {
  return mvae_latent_forward_ex(desc, B, H, h, ld_h, Wh, bh, const_cast<float*>(eps), radius, Wd0, bd0, ml, z, kl, dd_out,
                                nonfinite_flag, nullptr, stream);
}
