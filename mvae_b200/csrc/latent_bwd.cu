// latent_bwd.cu — backward instantiations + C-ABI entry of the fused latent block (see latent_impl.cuh).
#define MVAE_LAT_BWD 1
#include "latent_impl.cuh"

using namespace mvae;

static bool lat_aligned(const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

extern "C" int mvae_latent_backward(const mvae_pm_desc* desc, int64_t B, int32_t H, const mvae_planes* gdd,
                                    const mvae_planes* h, const float* Wh, const float* Wd0, const float* ml,
                                    const float* eps, const float* radius, const float* z, float gkl_scalar,
                                    const mvae_planes* gh_out, float* gWd0, float* gbd0, float* gWh, float* gbh,
                                    float* gradius, void* stream) {
  if (!desc || desc->C < 1 || desc->C > MVAE_MAX_COMPONENTS || B < 0 || H < 8 || !gdd || !gdd->base || !h || !h->base ||
      !Wh || !Wd0 || !ml || !eps || !z || !gh_out || !gh_out->base || !gWd0 || !gWh)
    return MVAE_ERR_INVALID_ARGUMENT;
  if (desc->ld_ml > 64 || desc->ld_z > 64 || (H & 7)) return MVAE_ERR_UNSUPPORTED;
  if (gdd->planes < 1 || gdd->planes > 3 || h->planes < 1 || h->planes > 3 || gh_out->planes < 1 || gh_out->planes > 3 ||
      gdd->ld < H || h->ld < H || gh_out->ld < H || gdd->rows < B || h->rows < B || gh_out->rows < B)
    return MVAE_ERR_INVALID_ARGUMENT;
  if ((gdd->ld & 7) || (gdd->plane_stride & 7) || (h->ld & 7) || (h->plane_stride & 7) || (gh_out->ld & 1) ||
      (gh_out->plane_stride & 1) || !lat_aligned(gdd->base, 16) || !lat_aligned(h->base, 16) ||
      !lat_aligned(gh_out->base, 4) || !lat_aligned(Wh, 8) || !lat_aligned(gWh, 8) || !lat_aligned(gbd0, 8) ||
      !lat_aligned(Wd0, 16) || !lat_aligned(gWd0, 16))
    return MVAE_ERR_ALIGNMENT;
  if (B == 0) return MVAE_OK;
  LatParams p;
  memset(&p, 0, sizeof(p));
  p.desc = *desc;
  p.B = B;
  p.H = H;
  p.h = h->base;
  p.h_stride = h->planes > 1 ? h->plane_stride : 0;
  p.h_ld = h->ld;
  p.h_planes = h->planes;
  p.gdd = gdd->base;
  p.gdd_stride = gdd->planes > 1 ? gdd->plane_stride : 0;
  p.gdd_ld = gdd->ld;
  p.gdd_planes = gdd->planes;
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
