// latent_bwd.cu — backward instantiations + C-ABI entry of the fused latent block (see latent_impl.cuh).
#define MVAE_LAT_BWD 1
#include "latent_impl.cuh"

using namespace mvae;

static bool lat_aligned(const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

extern "C" int mvae_latent_backward(const mvae_pm_desc* desc, int64_t B, int32_t H, const float* gdd, int64_t ld_gdd,
                                    const float* h, int64_t ld_h, const float* Wh, const float* Wd0, const float* ml,
                                    const float* eps, const float* radius, const float* z, float gkl_scalar,
                                    const mvae_planes* gh_out, float* gWd0, float* gbd0, float* gWh, float* gbh,
                                    float* gradius, void* stream) {
  if (!desc || desc->C < 1 || desc->C > MVAE_MAX_COMPONENTS || B < 0 || H < 8 || !gdd || !h ||
      !Wh || !Wd0 || !ml || !eps || !z || !gh_out || !gh_out->base || !gWd0 || !gWh)
    return MVAE_ERR_INVALID_ARGUMENT;
  if (desc->ld_ml > 64 || desc->ld_z > 64 || (H & 7)) return MVAE_ERR_UNSUPPORTED;
  if (gh_out->planes < 1 || gh_out->planes > 3 || ld_gdd < H || ld_h < H || gh_out->ld < H || gh_out->rows < B)
    return MVAE_ERR_INVALID_ARGUMENT;
  if ((ld_gdd & 3) || (ld_h & 3) || ld_gdd > 0x7fffffff || ld_h > 0x7fffffff || (gh_out->ld & 1) ||
      (gh_out->plane_stride & 1) || !lat_aligned(gdd, 16) || !lat_aligned(h, 16) ||
      !lat_aligned(gh_out->base, 4) || !lat_aligned(Wh, 8) || !lat_aligned(gWh, 8) || !lat_aligned(gbd0, 8) ||
      !lat_aligned(Wd0, 16) || !lat_aligned(gWd0, 16))
    return MVAE_ERR_ALIGNMENT;
  if (B == 0) return MVAE_OK;
  LatParams p;
  memset(&p, 0, sizeof(p));
  p.desc = *desc;
  p.B = B;
  p.H = H;
  p.h = h;
  p.h_ld = (int)ld_h;
  p.gdd = gdd;
  p.gdd_ld = (int)ld_gdd;
  p.Wh = Wh;
  p.Wd0 = Wd0;
  p.ml_in = ml;
  p.eps = eps;
  p.radius = radius;
  p.z_in = z;
  p.gkl = gkl_scalar;
  p.gh = gh_out->base;
  p.gh_stride = gh_out->planes > 1 ? gh_out->plane_stride : 0;
  p.gh_ld = gh_out->ld;
  p.gh_planes = gh_out->planes;
  p.gWd0 = gWd0;
  p.gbd0 = gbd0;
  p.gWh = gWh;
  p.gbh = gbh;
  p.gradius = gradius;
  return launch_latent(p, stream);
}
