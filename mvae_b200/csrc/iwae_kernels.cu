// iwae_kernels.cu — the latent part of ModelVAE.log_likelihood (mt/mvae/models/vae.py:82-123), the importance-weighted
// estimate the reference evaluates with n = 500 samples per input row:
//
//   mvae_iwae_latent    for every (sample s, row b): z and sum_c (log q_c - log p_c) for ALL components from ONE set of
//                       head pre-activations (rsample_log_probs, sampling_procedures.py:47-50,106-110).  The reference
//                       materialises [n, B, d] tensors per component and walks ~60 torch ops per component; here the
//                       rows' head pre-activations stay in shared memory for the whole sample loop, each sample reads
//                       its eps tile and writes its z tile once, and sum_s z (for cov_norm) accumulates in registers.
//   mvae_iwae_reduce    the two logsumexp's over the sample axis (vae.py:111-117), streaming (online max / sum).
//   mvae_iwae_cov_norm  vae.py:119-121 without the [n, B, D] repeat: the mean over samples commutes with the product.
//
// The arithmetic of a (component, sample) item is pm_math.cuh through dispatch_item (the training kernels' code), so the
// log q - log p of a wrapped-normal component is bit-identical to the training path's KL term.  Euclidean components
// differ: training uses the analytic KL (sampling_procedures.py:153-155), the likelihood estimate the Monte-Carlo
// difference log N(z; mu, sigma) - log N(z; 0, 1)  (wrapped_distributions.py:39-42).
#include "pm_item.cuh"

namespace mvae {

constexpr int kIwRows = 32;      // rows of the batch per CTA (lane = row)
constexpr int kIwThreads = 128;  // 4 warps: components are dealt round-robin to warps
constexpr int kIwMaxAcc = 16;    // z-sum accumulators per thread: 32 * ld_z / 128 <= 16  <=>  ld_z <= 64

struct IwParams {
  mvae_pm_desc desc;
  int64_t B;
  int ns;           // samples in this launch
  int s_per_cta;    // samples per CTA (grid.y slices of the sample axis)
  const float* ml;  // [B, P]
  const float* eps; // [ns, B, Sn]
  const float* radius;
  float* z;         // [ns, B, Sd]
  float* diff;      // [ns, B]
  float* zsum;      // [B, Sd] or nullptr
};

__device__ __forceinline__ int odd_stride(int x) { return x | 1; }  // lane = row: an odd stride is conflict-free

template <int MAXN>
__global__ void __launch_bounds__(kIwThreads) iwae_latent_kernel(const __grid_constant__ IwParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = p.desc.C, P = p.desc.ld_ml, Sn = p.desc.ld_eps, Sd = p.desc.ld_z;
  const int Pp = odd_stride(P), Snp = odd_stride(Sn), Sdp = odd_stride(Sd), Cp = odd_stride(C);
  ItemInfo* info = reinterpret_cast<ItemInfo*>(smem_raw);
  float* sML = reinterpret_cast<float*>(smem_raw + ((C * (int)sizeof(ItemInfo) + 15) / 16) * 16);
  float* sEPS = sML + kIwRows * Pp;
  float* sZ = sEPS + kIwRows * Snp;
  float* sMU = sZ + kIwRows * Sdp;
  float* sSG = sMU + kIwRows * Sdp;
  float* sKL = sSG + kIwRows * Snp;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b0 = (int64_t)blockIdx.x * kIwRows;
  const int rows = (int)((p.B - b0) < kIwRows ? (p.B - b0) : kIwRows);
  const int s0 = blockIdx.y * p.s_per_cta;
  const int s1 = (s0 + p.s_per_cta) < p.ns ? (s0 + p.s_per_cta) : p.ns;

  stage_items(info, p.desc, p.radius);
  for (int i = tid; i < rows * P; i += kIwThreads) sML[(i / P) * Pp + (i % P)] = __ldg(p.ml + b0 * P + i);
  auto load_eps = [&](int s) {
    const float* g = p.eps + ((int64_t)s * p.B + b0) * Sn;
    for (int i = tid; i < rows * Sn; i += kIwThreads) sEPS[(i / Sn) * Snp + (i % Sn)] = __ldg(g + i);
  };
  if (s0 < s1) load_eps(s0);
  __syncthreads();

  float zacc[kIwMaxAcc];
#pragma unroll
  for (int k = 0; k < kIwMaxAcc; ++k) zacc[k] = 0.f;

  for (int s = s0; s < s1; ++s) {
    // ---- one warp per component, lane = row: z, mu, sigma and log q - log p of this sample ----
    for (int ci = warp; ci < C; ci += kIwThreads / 32) {
      if (lane < rows) {
        const ItemInfo c = info[ci];
        float unused = 0.f;
        float* kl = sKL + lane * Cp + ci;
        dispatch_item<false, MAXN, true>(c, sML + lane * Pp, sEPS + lane * Snp, sZ + lane * Sdp, kl, sMU + lane * Sdp,
                                         sSG + lane * Snp, nullptr, 0.f, nullptr, &unused, false);
        if (c.type == MVAE_EUCLIDEAN) {
          // Monte-Carlo difference instead of the analytic KL:
          //   log N(z; mu, sigma) - log N(z; 0, 1) = sum_j [ -eps_j^2/2 - log sigma_j + z_j^2/2 ]
          const float* e = sEPS + lane * Snp + c.eps_off;
          const float* sg = sSG + lane * Snp + c.eps_off;
          const float* zz = sZ + lane * Sdp + c.z_off;
          float acc = 0.f, prod = 1.f;
          for (int j = 0; j < c.n; ++j) {
            acc = fmaf(zz[j], zz[j], acc);
            acc = fmaf(-e[j], e[j], acc);
            prod *= sg[j];
            if ((j & 3) == 3) {  // one logarithm per four factors (sigma >= 1e-5: no underflow)
              acc -= 2.f * logf(prod);
              prod = 1.f;
            }
          }
          acc -= 2.f * logf(prod);
          *kl = 0.5f * acc;
        }
      }
    }
    __syncthreads();
    // ---- stores of this sample + loads of the next one ----
    {
      float* gz = p.z + ((int64_t)s * p.B + b0) * Sd;
      int k = 0;
      for (int i = tid; i < rows * Sd; i += kIwThreads, ++k) {
        const float v = sZ[(i / Sd) * Sdp + (i % Sd)];
        gz[i] = v;
        if (k < kIwMaxAcc) zacc[k] += v;
      }
      if (tid < rows) {
        float d = 0.f;
        for (int ci = 0; ci < C; ++ci) d += sKL[tid * Cp + ci];
        p.diff[(int64_t)s * p.B + b0 + tid] = d;
      }
      if (s + 1 < s1) load_eps(s + 1);
    }
    __syncthreads();
  }
  if (p.zsum) {
    int k = 0;
    for (int i = tid; i < rows * Sd; i += kIwThreads, ++k)
      if (k < kIwMaxAcc) atomicAdd(p.zsum + b0 * Sd + i, zacc[k]);
  }
}

// ------------------------------------------------------------------------------------------------ logsumexp over s
// blockDim = (32 rows, 8 sample groups): coalesced over b, each thread keeps an online (max, sum) pair per estimate.
__device__ __forceinline__ void lse_push(float& m, float& a, float v) {
  if (v > m) {
    a = a * __expf(m - v) + 1.f;
    m = v;
  } else {
    a += __expf(v - m);
  }
}
__device__ __forceinline__ void lse_merge(float& m, float& a, float m2, float a2) {
  if (a2 == 0.f) return;
  if (a == 0.f) {
    m = m2;
    a = a2;
    return;
  }
  const float mm = fmaxf(m, m2);
  a = a * __expf(m - mm) + a2 * __expf(m2 - mm);
  m = mm;
}

__global__ void __launch_bounds__(256) iwae_reduce_kernel(int n, int64_t B, const float* __restrict__ recon,
                                                          const float* __restrict__ diff, float* __restrict__ ll,
                                                          float* __restrict__ mi) {
  __shared__ float sm[4][8][33];
  const int64_t b = (int64_t)blockIdx.x * 32 + threadIdx.x;
  float m1 = -INFINITY, a1 = 0.f, m2 = -INFINITY, a2 = 0.f;
  if (b < B) {
    for (int s = threadIdx.y; s < n; s += 8) {
      const float d = __ldg(diff + (int64_t)s * B + b);
      const float r = __ldg(recon + (int64_t)s * B + b);
      lse_push(m1, a1, -r - d);  // log p(x|z) + log p(z) - log q(z|x)
      lse_push(m2, a2, d);       // log q(z|x) - log p(z)
    }
  }
  sm[0][threadIdx.y][threadIdx.x] = m1;
  sm[1][threadIdx.y][threadIdx.x] = a1;
  sm[2][threadIdx.y][threadIdx.x] = m2;
  sm[3][threadIdx.y][threadIdx.x] = a2;
  __syncthreads();
  if (threadIdx.y == 0 && b < B) {
    for (int g = 1; g < 8; ++g) {
      lse_merge(m1, a1, sm[0][g][threadIdx.x], sm[1][g][threadIdx.x]);
      lse_merge(m2, a2, sm[2][g][threadIdx.x], sm[3][g][threadIdx.x]);
    }
    const float ln_n = logf((float)n);
    ll[b] = m1 + logf(a1) - ln_n;
    mi[b] = m2 + logf(a2) - ln_n;
  }
}

// ------------------------------------------------------------------------------------------------ cov_norm
// zc[b, j] = zsum[b, j] / n - mean_b(zsum[., j] / n): one CTA per latent coordinate (two passes over a column).
__global__ void __launch_bounds__(256) iwae_center_kernel(int64_t B, int Sd, float inv_n, const float* __restrict__ zsum,
                                                          float* __restrict__ zc) {
  __shared__ float red[8];
  const int j = blockIdx.x;
  float acc = 0.f;
  for (int64_t b = threadIdx.x; b < B; b += 256) acc += __ldg(zsum + b * Sd + j);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float mean = tot / (float)B;
  for (int64_t b = threadIdx.x; b < B; b += 256) zc[b * Sd + j] = (__ldg(zsum + b * Sd + j) - mean) * inv_n;
}

// cov[d, j] = sum_b x[b, d] zc[b, j]   (the column means of x drop out: sum_b zc[b, j] = 0).
// One thread per input column d (coalesced over d), a tile of rows of zc staged in shared memory; grid.y slices the
// batch, partial sums leave through atomics into the zero-initialised cov.
constexpr int kCovRows = 64;
template <int SMAX>
__global__ void __launch_bounds__(128) iwae_cov_kernel(int64_t B, int D, int Sd, int64_t rows_per_cta,
                                                       const float* __restrict__ x, const float* __restrict__ zc,
                                                       float* __restrict__ cov) {
  __shared__ float sz[kCovRows * SMAX];
  const int d = blockIdx.x * 128 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = (r0 + rows_per_cta) < B ? (r0 + rows_per_cta) : B;
  float acc[SMAX];
#pragma unroll
  for (int j = 0; j < SMAX; ++j) acc[j] = 0.f;
  for (int64_t rb = r0; rb < r1; rb += kCovRows) {
    const int nr = (int)((r1 - rb) < kCovRows ? (r1 - rb) : kCovRows);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * Sd; i += 128) sz[(i / Sd) * SMAX + (i % Sd)] = __ldg(zc + rb * Sd + i);
    __syncthreads();
    if (d < D) {
      for (int r = 0; r < nr; ++r) {
        const float xv = __ldg(x + (rb + r) * D + d);
#pragma unroll
        for (int j = 0; j < SMAX; ++j)
          if (j < Sd) acc[j] = fmaf(xv, sz[r * SMAX + j], acc[j]);
      }
    }
  }
  if (d < D) {
#pragma unroll
    for (int j = 0; j < SMAX; ++j)
      if (j < Sd) atomicAdd(cov + (int64_t)d * Sd + j, acc[j]);
  }
}

__global__ void __launch_bounds__(256) iwae_fro_kernel(int64_t n, const float* __restrict__ v, float* __restrict__ out) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 256) {
    const float t = __ldg(v + i);
    acc = fmaf(t, t, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w];
    out[0] = sqrtf(tot);
  }
}

}  // namespace mvae

using namespace mvae;

extern "C" int mvae_iwae_latent(const mvae_pm_desc* desc, int64_t B, int32_t ns, const float* ml, const float* eps,
                                const float* radius, float* z, float* diff, float* zsum, void* stream) {
  if (!desc || desc->C < 1 || desc->C > MVAE_MAX_COMPONENTS || B < 0 || ns < 0 || !ml || !eps || !z || !diff)
    return MVAE_ERR_INVALID_ARGUMENT;
  if (desc->ld_ml > 64 || desc->ld_z > 64) return MVAE_ERR_UNSUPPORTED;
  bool dyn = false, any_curved = false;
  int maxn = 0;
  for (int i = 0; i < desc->C; ++i) {
    const mvae_component& c = desc->comp[i];
    if (c.type < MVAE_EUCLIDEAN || c.type > MVAE_UNIVERSAL || c.n < 1) return MVAE_ERR_INVALID_ARGUMENT;
    any_curved = any_curved || c.type != MVAE_EUCLIDEAN;
    dyn = dyn || !(c.n >= 1 && c.n <= 8 && c.n != 7);
    maxn = c.n > maxn ? c.n : maxn;
  }
  if (any_curved && !radius) return MVAE_ERR_INVALID_ARGUMENT;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  if (B == 0 || ns == 0) return MVAE_OK;
  IwParams p;
  memset(&p, 0, sizeof(p));
  p.desc = *desc;
  p.B = B;
  p.ns = ns;
  p.ml = ml;
  p.eps = eps;
  p.radius = radius;
  p.z = z;
  p.diff = diff;
  p.zsum = zsum;
  const int64_t row_tiles = (B + kIwRows - 1) / kIwRows;
  if (row_tiles > 0x7fffffff) return MVAE_ERR_UNSUPPORTED;
  // slice the sample axis so that the grid covers the SMs a few times over; each CTA keeps >= 4 samples so that the
  // head pre-activations it staged are reused
  int64_t want = (int64_t)di.sm_count * 8;
  int slices = (int)((want + row_tiles - 1) / row_tiles);
  if (slices > (ns + 3) / 4) slices = (ns + 3) / 4;
  if (slices < 1) slices = 1;
  if (slices > 65535) slices = 65535;
  p.s_per_cta = (ns + slices - 1) / slices;
  slices = (ns + p.s_per_cta - 1) / p.s_per_cta;
  const int P = desc->ld_ml, Sn = desc->ld_eps, Sd = desc->ld_z, C = desc->C;
  const size_t smem = (size_t)((C * (int)sizeof(ItemInfo) + 15) / 16) * 16 +
                      4u * kIwRows * ((P | 1) + 2 * (Sn | 1) + 2 * (Sd | 1) + (C | 1));
  if (smem > (size_t)di.max_smem_optin) return MVAE_ERR_UNSUPPORTED;
  void (*kern)(const IwParams);
  if (dyn) kern = iwae_latent_kernel<0>;
  else if (maxn <= 2) kern = iwae_latent_kernel<2>;
  else kern = iwae_latent_kernel<8>;
  if (smem > 48 * 1024)
    MVAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3((unsigned)row_tiles, (unsigned)slices), dim3(kIwThreads), smem, as_stream(stream)>>>(p);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

extern "C" int mvae_iwae_reduce(int32_t n, int64_t B, const float* recon, const float* diff, float* log_p_x, float* mi,
                                void* stream) {
  if (n < 1 || B < 0 || !recon || !diff || !log_p_x || !mi) return MVAE_ERR_INVALID_ARGUMENT;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  if (B == 0) return MVAE_OK;
  const int64_t grid = (B + 31) / 32;
  if (grid > 0x7fffffff) return MVAE_ERR_UNSUPPORTED;
  iwae_reduce_kernel<<<dim3((unsigned)grid), dim3(32, 8), 0, as_stream(stream)>>>(n, B, recon, diff, log_p_x, mi);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

extern "C" int mvae_iwae_cov_norm(int64_t B, int32_t D, int32_t Sd, int32_t n, const float* x, const float* zsum,
                                  float* work, float* out, void* stream) {
  if (B < 1 || D < 1 || Sd < 1 || n < 1 || !x || !zsum || !work || !out) return MVAE_ERR_INVALID_ARGUMENT;
  if (Sd > 64) return MVAE_ERR_UNSUPPORTED;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  cudaStream_t st = as_stream(stream);
  float* zc = work;
  float* cov = work + B * Sd;
  MVAE_CUDA_TRY(cudaMemsetAsync(cov, 0, sizeof(float) * (size_t)D * Sd, st));
  iwae_center_kernel<<<dim3((unsigned)Sd), dim3(256), 0, st>>>(B, Sd, 1.f / (float)n, zsum, zc);
  MVAE_LAUNCH_CHECK();
  const int col_tiles = (D + 127) / 128;
  int64_t slices = ((int64_t)di.sm_count * 4 + col_tiles - 1) / col_tiles;
  const int64_t max_slices = (B + kCovRows - 1) / kCovRows;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  if (slices > 65535) slices = 65535;
  int64_t rows_per_cta = (B + slices - 1) / slices;
  rows_per_cta = (rows_per_cta + kCovRows - 1) / kCovRows * kCovRows;
  slices = (B + rows_per_cta - 1) / rows_per_cta;
  const dim3 grid((unsigned)col_tiles, (unsigned)slices);
  if (Sd <= 8) iwae_cov_kernel<8><<<grid, dim3(128), 0, st>>>(B, D, Sd, rows_per_cta, x, zc, cov);
  else if (Sd <= 16) iwae_cov_kernel<16><<<grid, dim3(128), 0, st>>>(B, D, Sd, rows_per_cta, x, zc, cov);
  else if (Sd <= 32) iwae_cov_kernel<32><<<grid, dim3(128), 0, st>>>(B, D, Sd, rows_per_cta, x, zc, cov);
  else iwae_cov_kernel<64><<<grid, dim3(128), 0, st>>>(B, D, Sd, rows_per_cta, x, zc, cov);
  MVAE_LAUNCH_CHECK();
  iwae_fro_kernel<<<dim3(1), dim3(256), 0, st>>>((int64_t)D * Sd, cov, out);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}
