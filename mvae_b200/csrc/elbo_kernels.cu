// elbo_kernels.cu — reconstruction losses, ELBO reduction, optimizer steps and the fp32 -> split-bf16 plane
// conversion (K5 / K7 of SURVEY.md §2.2).  All HBM-bound elementwise / reduction kernels: 128-bit coalesced
// accesses, warp-shuffle reductions, one atomic per CTA per output word.
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include "mvae_common.cuh"

namespace mvae {

constexpr float kHalfLn2PiF = 0.9189385332046727f;

// ------------------------------------------------------------------------------------- reconstruction loss
// kind 0: F.binary_cross_entropy_with_logits(reduction='none') (mt/data/image_reconstruction.py:81-82) in ATen's
//         form (1-t)*x - log_sigmoid(x), log_sigmoid(x) = min(x,0) - log1p(exp(-|x|));
// kind 1: -Normal(x_, 1).log_prob(x) (mt/data/synthetic.py:161-162) = (x - x_)^2/2 + ln(sqrt(2 pi)).
// Row sums as in mt/mvae/models/vae.py:131.  One warp per row.
template <int KIND>
__device__ __forceinline__ float recon_elem(float lg, float t, float* g) {
  if (KIND == 0) {
    float ab = fabsf(lg);
    float ex = expf(-ab);
    float mn = fminf(lg, 0.f);
    float loss = (1.f - t) * lg - (mn - log1pf(ex));
    *g = 1.f / (1.f + expf(-lg)) - t;
    return loss;
  } else {
    float dlt = t - lg;
    *g = lg - t;
    return (dlt * dlt) / 2.f + kHalfLn2PiF;
  }
}

template <int KIND, bool GRAD>
__global__ void __launch_bounds__(256) recon_kernel(int64_t B, int D, const float* __restrict__ logits,
                                                    const float* __restrict__ x, float* __restrict__ rowsum,
                                                    float* __restrict__ glogits, int vec) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t b = warp; b < B; b += nwarps) {
    const float* lr = logits + b * D;
    const float* xr = x + b * D;
    float* gr = GRAD ? glogits + b * D : nullptr;
    float acc = 0.f;
    int done = 0;
    if (vec) {
      const int nv = D >> 2;
      for (int i = lane; i < nv; i += 32) {
        float4 l4 = *reinterpret_cast<const float4*>(lr + 4 * i);
        float4 t4 = *reinterpret_cast<const float4*>(xr + 4 * i);
        float4 g4;
        acc += recon_elem<KIND>(l4.x, t4.x, &g4.x);
        acc += recon_elem<KIND>(l4.y, t4.y, &g4.y);
        acc += recon_elem<KIND>(l4.z, t4.z, &g4.z);
        acc += recon_elem<KIND>(l4.w, t4.w, &g4.w);
        if (GRAD) *reinterpret_cast<float4*>(gr + 4 * i) = g4;
      }
      done = nv << 2;
    }
    for (int i = done + lane; i < D; i += 32) {
      float g;
      acc += recon_elem<KIND>(lr[i], xr[i], &g);
      if (GRAD) gr[i] = g;
    }
    acc = warp_sum(acc);
    if (lane == 0) rowsum[b] = acc;
  }
}

// ------------------------------------------------------------------------------------------- ELBO reduce
// BatchStats (mt/mvae/stats.py:144-202): out = [sum bce, sum_b sum_c kl, sum_b(-bce_b - beta*kl_b), sum_b kl_bc (C)].
// out must be zero on entry (the host entry point memsets it); one atomicAdd per CTA per output word.
__global__ void __launch_bounds__(256) elbo_kernel(int64_t B, int C, const float* __restrict__ bce,
                                                   const float* __restrict__ kl, float beta,
                                                   float* __restrict__ out) {
  extern __shared__ float sh[];  // [(3 + C)] block accumulators
  const int nout = 3 + C;
  for (int i = threadIdx.x; i < nout; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  float sb = 0.f, sk = 0.f, se = 0.f;
  const int lane = threadIdx.x & 31;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    float klb = 0.f;
    for (int c = 0; c < C; ++c) klb += kl[b * C + c];
    float bb = bce[b];
    sb += bb;
    sk += klb;
    se += -bb - beta * klb;
  }
  sb = warp_sum(sb);
  sk = warp_sum(sk);
  se = warp_sum(se);
  if (lane == 0) {
    atomicAdd(&sh[0], sb);
    atomicAdd(&sh[1], sk);
    atomicAdd(&sh[2], se);
  }
  // per-component sums: thread t owns column (t % C) of a strided sweep so that loads stay coalesced
  {
    const int64_t total = B * C;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // make the stride a multiple of C so that each thread always sees the same component
    const int64_t step = (stride / C) * C;
    const int64_t start = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (step > 0 && start < step) {
      float acc = 0.f;
      for (int64_t i = start; i < total; i += step) acc += kl[i];
      atomicAdd(&sh[3 + (int)(start % C)], acc);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nout; i += blockDim.x)
    if (sh[i] != 0.f) atomicAdd(out + i, sh[i]);
}

// Small batches (the training configurations): ONE CTA of 1024 threads, block reduction, results written directly —
// no memset, no atomics, deterministic summation order.
__global__ void __launch_bounds__(1024) elbo_small_kernel(int B, int C, const float* __restrict__ bce,
                                                          const float* __restrict__ kl, float beta,
                                                          float* __restrict__ out) {
  extern __shared__ float sh[];  // [32][3 + C] per-warp partials
  pdl_launch_dependents();
  pdl_wait();
  const int nout = 3 + C, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float sb = 0.f, sk = 0.f, se = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float klb = 0.f;
    for (int c = 0; c < C; ++c) klb += kl[(int64_t)b * C + c];
    const float bb = bce[b];
    sb += bb;
    sk += klb;
    se += -bb - beta * klb;
  }
  sb = warp_sum(sb);
  sk = warp_sum(sk);
  se = warp_sum(se);
  if (lane == 0) {
    sh[warp * nout + 0] = sb;
    sh[warp * nout + 1] = sk;
    sh[warp * nout + 2] = se;
  }
  // per-component sums: warp w sweeps components w, w + 32, ...
  for (int c = warp; c < C; c += 32) {
    float acc = 0.f;
    for (int b = lane; b < B; b += 32) acc += kl[(int64_t)b * C + c];
    acc = warp_sum(acc);
    if (lane == 0) sh[32 * nout + c] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float acc = 0.f;
    for (int w = 0; w < 32; ++w) acc += sh[w * nout + threadIdx.x];
    out[threadIdx.x] = acc;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) out[3 + c] = sh[32 * nout + c];
}

// --------------------------------------------------------------------------------------------- optimizers
// torch.optim.Adam single-tensor semantics (no amsgrad / weight decay): exp_avg.lerp_(g, 1-b1);
// exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2); denom = sqrt(v)/sqrt(bc2) + eps; p -= (lr/bc1) * m / denom.
__global__ void __launch_bounds__(256) adam_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, float w1, float b2,
                                                   float w2, float step_size, float inv_bc2_sqrt, float eps,
                                                   float gscale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float mi = m[i];
    mi = mi + (gi - mi) * w1;
    float vi = v[i] * b2 + w2 * (gi * gi);
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

// Same update with the step counter living on the device (CUDA-graph replay: nothing step-dependent is baked into
// the launch parameters).  bump_step_kernel runs first in stream order.
__global__ void bump_step_kernel(int32_t* step) { *step += 1; }
__global__ void __launch_bounds__(256) adam_dev_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v, float lr,
                                                       float b1, float b2, float eps, const int32_t* __restrict__ step,
                                                       float gscale) {
  const double st = (double)(*step);
  const float step_size = (float)((double)lr / (1.0 - pow((double)b1, st)));
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(1.0 - pow((double)b2, st)));
  const float w1 = 1.f - b1, w2 = 1.f - b2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * gscale;
    float mi = m[i];
    mi = mi + (gi - mi) * w1;
    float vi = v[i] * b2 + w2 * (gi * gi);
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

__global__ void __launch_bounds__(256) sgd_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g,
                                                  float lr, float gscale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = p[i] - lr * (g[i] * gscale);
}

// ------------------------------------------------------------------------------------------- split planes
// X [R,K] fp32 -> X_0 = bf16(X), X_1 = bf16(X - X_0), X_2 = bf16(X - X_0 - X_1)  (round-to-nearest-even),
// optionally also the transposed planes [K,R].  32x32 tiles through shared memory keep both sides coalesced.
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ src, int64_t ld_src, int R, int K,
                                                    mvae_planes dst, mvae_planes dstT, int has_dst, int has_T) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, k = k0 + tx;
    float v = (r < R && k < K) ? src[(int64_t)r * ld_src + k] : 0.f;
    tile[ty + 8 * i][tx] = v;
    if (has_dst && r < R && k < K) {
      float rem = v;
      for (int pl = 0; pl < dst.planes; ++pl) {
        __nv_bfloat16 h = __float2bfloat16_rn(rem);
        reinterpret_cast<__nv_bfloat16*>(dst.base)[pl * dst.plane_stride + (int64_t)r * dst.ld + k] = h;
        rem -= __bfloat162float(h);
      }
    }
  }
  if (!has_T) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty + 8 * i, r = r0 + tx;
    if (k < K && r < R) {
      float rem = tile[tx][ty + 8 * i];
      for (int pl = 0; pl < dstT.planes; ++pl) {
        __nv_bfloat16 h = __float2bfloat16_rn(rem);
        reinterpret_cast<__nv_bfloat16*>(dstT.base)[pl * dstT.plane_stride + (int64_t)k * dstT.ld + r] = h;
        rem -= __bfloat162float(h);
      }
    }
  }
}

// Row-major only (no transposed copy), K % 8 == 0: a thread converts 8 consecutive values — two 128-bit loads, one
// 128-bit store per plane.
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ src, int64_t ld_src, int R, int K8,
                                                         mvae_planes dst) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = (int64_t)R * K8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / K8), k = (int)(i - (int64_t)r * K8) * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + (int64_t)r * ld_src + k));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + (int64_t)r * ld_src + k) + 1);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    for (int pl = 0; pl < dst.planes; ++pl) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        w[j] = *reinterpret_cast<const uint32_t*>(&h);
        v[2 * j] -= __uint_as_float(w[j] << 16);
        v[2 * j + 1] -= __uint_as_float(w[j] & 0xFFFF0000u);
      }
      *reinterpret_cast<uint4*>(dst.base + pl * dst.plane_stride + (int64_t)r * dst.ld + k) =
          make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// ------------------------------------------------------------------------------------- fused optimizer step
// Adam over the flat parameter bucket (128-bit accesses) + the split-bf16 planes of the weight matrices that feed the
// tensor-core GEMMs, written from the freshly updated values + the radii's SGD step + the device step counter: what
// used to be five launches (bump, Adam, SGD, two plane splits) at the end of every step.
struct PlaneTarget {
  int64_t begin, end;  // float range [begin, end) of the flat bucket holding a [rows, cols] matrix (both % 4 == 0)
  int cols, ld, planes;
  uint16_t* base;
  int64_t plane_stride;
};
struct OptParams {
  int64_t n4;
  float* p;
  const float* g;
  float* m;
  float* v;
  float lr, b1, b2, eps;
  int32_t* step_dev;
  uint32_t* done;  // zero-initialised counter (last-CTA-done pattern)
  float* radius;
  const float* gradius;
  const float* radius_mask;
  float radius_lr;
  int C;
  int n_targets;
  PlaneTarget t[8];
};
__global__ void __launch_bounds__(256) opt_fused_kernel(const __grid_constant__ OptParams q) {
  pdl_launch_dependents();
  pdl_wait();
  const double st = (double)(*q.step_dev + 1);
  const float step_size = (float)((double)q.lr / (1.0 - pow((double)q.b1, st)));
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(1.0 - pow((double)q.b2, st)));
  const float w1 = 1.f - q.b1, w2 = 1.f - q.b2;
  if (blockIdx.x == 0 && q.radius && q.radius_lr != 0.f)
    for (int t = threadIdx.x; t < q.C; t += blockDim.x)
      q.radius[t] = q.radius[t] - q.radius_lr * (q.gradius[t] * (q.radius_mask ? q.radius_mask[t] : 1.f));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q.n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 g = reinterpret_cast<const float4*>(q.g)[i];
    float4 mi = reinterpret_cast<float4*>(q.m)[i], vi = reinterpret_cast<float4*>(q.v)[i];
    float4 pi = reinterpret_cast<float4*>(q.p)[i];
#define MVAE_ADAM1(c)                                       \
  mi.c = mi.c + (g.c - mi.c) * w1;                          \
  vi.c = vi.c * q.b2 + w2 * (g.c * g.c);                    \
  pi.c = pi.c - step_size * (mi.c / (sqrtf(vi.c) * inv_bc2_sqrt + q.eps));
    MVAE_ADAM1(x) MVAE_ADAM1(y) MVAE_ADAM1(z) MVAE_ADAM1(w)
#undef MVAE_ADAM1
    reinterpret_cast<float4*>(q.m)[i] = mi;
    reinterpret_cast<float4*>(q.v)[i] = vi;
    reinterpret_cast<float4*>(q.p)[i] = pi;
    const int64_t idx = 4 * i;
#pragma unroll
    for (int t = 0; t < 8; ++t)
      if (t < q.n_targets && idx >= q.t[t].begin && idx < q.t[t].end) {
        const int64_t rel = idx - q.t[t].begin;
        const int64_t r = rel / q.t[t].cols;
        const int c = (int)(rel - r * q.t[t].cols);
        float v4[4] = {pi.x, pi.y, pi.z, pi.w};
        for (int pl = 0; pl < q.t[t].planes; ++pl) {
          const __nv_bfloat162 h0 = __floats2bfloat162_rn(v4[0], v4[1]), h1 = __floats2bfloat162_rn(v4[2], v4[3]);
          const uint32_t u0 = *reinterpret_cast<const uint32_t*>(&h0), u1 = *reinterpret_cast<const uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(q.t[t].base + pl * q.t[t].plane_stride + r * q.t[t].ld + c) = make_uint2(u0, u1);
          v4[0] -= __uint_as_float(u0 << 16);
          v4[1] -= __uint_as_float(u0 & 0xFFFF0000u);
          v4[2] -= __uint_as_float(u1 << 16);
          v4[3] -= __uint_as_float(u1 & 0xFFFF0000u);
        }
      }
  }
  // the last CTA to finish bumps the step counter (every CTA has read it by then); a launch without a done counter
  // (an early launch over part of the parameters: the rest follows in the launch that ends the step) leaves it alone
  if (!q.done) return;
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(q.done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *q.step_dev += 1;
    *q.done = 0u;
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int planes_ok(const mvae_planes* p, int rows, int cols) {
  if (!p->base || p->planes < 1 || p->planes > 3) return MVAE_ERR_INVALID_ARGUMENT;
  if (p->rows != rows || p->cols != cols || p->ld < cols) return MVAE_ERR_INVALID_ARGUMENT;
  if ((p->ld & 7) || (p->plane_stride & 7) || !aligned16(p->base)) return MVAE_ERR_ALIGNMENT;
  if (p->planes > 1 && p->plane_stride < (int64_t)rows * p->ld) return MVAE_ERR_INVALID_ARGUMENT;
  return MVAE_OK;
}

}  // namespace mvae

using namespace mvae;

extern "C" int mvae_recon_loss(int32_t kind, int64_t B, int32_t D, const float* logits, const float* x,
                               float* rowsum, float* glogits, void* stream) {
  if ((kind != 0 && kind != 1) || B < 0 || D < 1) return MVAE_ERR_INVALID_ARGUMENT;
  if (B == 0) return MVAE_OK;
  if (!logits || !x || !rowsum) return MVAE_ERR_INVALID_ARGUMENT;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  const int vec = (D % 4 == 0) && aligned16(logits) && aligned16(x) && (!glogits || aligned16(glogits));
  const int64_t want = (B + 7) / 8;
  const int grid = (int)(want < (int64_t)di.sm_count * 8 ? want : (int64_t)di.sm_count * 8);
  cudaStream_t s = as_stream(stream);
  if (kind == 0) {
    if (glogits) recon_kernel<0, true><<<grid, 256, 0, s>>>(B, D, logits, x, rowsum, glogits, vec);
    else recon_kernel<0, false><<<grid, 256, 0, s>>>(B, D, logits, x, rowsum, nullptr, vec);
  } else {
    if (glogits) recon_kernel<1, true><<<grid, 256, 0, s>>>(B, D, logits, x, rowsum, glogits, vec);
    else recon_kernel<1, false><<<grid, 256, 0, s>>>(B, D, logits, x, rowsum, nullptr, vec);
  }
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

extern "C" int mvae_elbo_reduce(int64_t B, int32_t C, const float* bce, const float* kl, float beta, float* out,
                                void* stream) {
  if (B < 0 || C < 1 || C > MVAE_MAX_COMPONENTS || !out) return MVAE_ERR_INVALID_ARGUMENT;
  cudaStream_t s = as_stream(stream);
  if (B > 0 && (!bce || !kl)) return MVAE_ERR_INVALID_ARGUMENT;
  if (B > 0 && B <= 65536) {
    MVAE_CUDA_TRY(launch_pdl(elbo_small_kernel, dim3(1), dim3(1024), sizeof(float) * (33 * (3 + C)), s, (int)B, C, bce, kl,
                             beta, out));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
  }
  MVAE_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(float) * (3 + C), s));
  if (B == 0) return MVAE_OK;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  const int64_t want = (B + 255) / 256;
  const int grid = (int)(want < (int64_t)di.sm_count * 2 ? want : (int64_t)di.sm_count * 2);
  elbo_kernel<<<grid, 256, sizeof(float) * (3 + C), s>>>(B, C, bce, kl, beta, out);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

extern "C" int mvae_adam_step(int64_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr,
                              float beta1, float beta2, float eps, int32_t step, float grad_scale, void* stream) {
  if (n < 0 || step < 1) return MVAE_ERR_INVALID_ARGUMENT;
  if (n == 0) return MVAE_OK;
  if (!param || !grad || !exp_avg || !exp_avg_sq) return MVAE_ERR_INVALID_ARGUMENT;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  const int64_t want = (n + 255) / 256;
  const int grid = (int)(want < (int64_t)di.sm_count * 8 ? want : (int64_t)di.sm_count * 8);
  adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(n, param, grad, exp_avg, exp_avg_sq, 1.f - beta1, beta2,
                                                   1.f - beta2, step_size, inv_bc2_sqrt, eps, grad_scale);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

extern "C" int mvae_adam_step_dev(int64_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                  float lr, float beta1, float beta2, float eps, int32_t* step_dev, float grad_scale,
                                  void* stream) {
  if (n < 0 || !step_dev) return MVAE_ERR_INVALID_ARGUMENT;
  if (n > 0 && (!param || !grad || !exp_avg || !exp_avg_sq)) return MVAE_ERR_INVALID_ARGUMENT;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  bump_step_kernel<<<1, 1, 0, as_stream(stream)>>>(step_dev);
  MVAE_LAUNCH_CHECK();
  if (n == 0) return MVAE_OK;
  const int64_t want = (n + 255) / 256;
  const int grid = (int)(want < (int64_t)di.sm_count * 8 ? want : (int64_t)di.sm_count * 8);
  adam_dev_kernel<<<grid, 256, 0, as_stream(stream)>>>(n, param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps,
                                                       step_dev, grad_scale);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

extern "C" int mvae_sgd_step(int64_t n, float* param, const float* grad, float lr, float grad_scale, void* stream) {
  if (n < 0) return MVAE_ERR_INVALID_ARGUMENT;
  if (n == 0) return MVAE_OK;
  if (!param || !grad) return MVAE_ERR_INVALID_ARGUMENT;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  const int64_t want = (n + 255) / 256;
  const int grid = (int)(want < (int64_t)di.sm_count * 8 ? want : (int64_t)di.sm_count * 8);
  sgd_kernel<<<grid, 256, 0, as_stream(stream)>>>(n, param, grad, lr, grad_scale);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

extern "C" int mvae_split_planes(const float* src, int64_t ld_src, int32_t R, int32_t K, const mvae_planes* dst,
                                 const mvae_planes* dst_transposed, void* stream) {
  if (!src || R < 0 || K < 0 || ld_src < K || (!dst && !dst_transposed)) return MVAE_ERR_INVALID_ARGUMENT;
  if (R == 0 || K == 0) return MVAE_OK;
  mvae_planes d = {}, t = {};
  if (dst) {
    int rc = planes_ok(dst, R, K);
    if (rc != MVAE_OK) return rc;
    d = *dst;
  }
  if (dst_transposed) {
    int rc = planes_ok(dst_transposed, K, R);
    if (rc != MVAE_OK) return rc;
    t = *dst_transposed;
  }
  if (!dst_transposed && (K & 7) == 0 && (ld_src & 3) == 0 && aligned16(src)) {
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc != MVAE_OK) return rc;
    const int64_t want = ((int64_t)R * (K / 8) + 255) / 256;
    const int g1 = (int)(want < (int64_t)di.sm_count * 8 ? want : (int64_t)di.sm_count * 8);
    MVAE_CUDA_TRY(launch_pdl(split_rows_kernel, dim3(g1), dim3(256), 0, as_stream(stream), src, ld_src, R, K / 8, d));
    MVAE_LAUNCH_CHECK();
    return MVAE_OK;
  }
  dim3 grid((K + 31) / 32, (R + 31) / 32);
  if (grid.y > 65535) return MVAE_ERR_UNSUPPORTED;
  split_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, ld_src, R, K, d, t, dst != nullptr,
                                                    dst_transposed != nullptr);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

namespace mvae {
__global__ void __launch_bounds__(128) clip_grad_norm_kernel(int n, float* __restrict__ g, const float* __restrict__ mask,
                                                             float max_norm) {
  __shared__ float red[4];
  __shared__ float coef;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 128) {
    const float v = mask[i] != 0.f ? g[i] : 0.f;
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float norm = sqrtf(red[0] + red[1] + red[2] + red[3]);
    coef = fminf(max_norm / (norm + 1e-6f), 1.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += 128)
    if (mask[i] != 0.f) g[i] *= coef;
}
}  // namespace mvae

extern "C" int mvae_clip_grad_norm(int32_t n, float* grad, const float* mask, float max_norm, void* stream) {
  if (n < 0 || n > 4096 || !(max_norm > 0.f)) return MVAE_ERR_INVALID_ARGUMENT;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  if (n == 0) return MVAE_OK;
  if (!grad || !mask) return MVAE_ERR_INVALID_ARGUMENT;
  clip_grad_norm_kernel<<<1, 128, 0, as_stream(stream)>>>(n, grad, mask, max_norm);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

extern "C" int mvae_opt_step_fused(int64_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                   float lr, float beta1, float beta2, float eps, int32_t* step_dev,
                                   uint32_t* done_counter, float* radius, const float* gradius,
                                   const float* radius_mask, float radius_lr, int32_t C, int32_t n_targets,
                                   const int64_t* target_begin, const int32_t* target_rows,
                                   const mvae_planes* targets, void* stream) {
  if (n < 0 || (n & 3) || !step_dev || n_targets < 0 || n_targets > 8 || C < 0) return MVAE_ERR_INVALID_ARGUMENT;
  if (n > 0 && (!param || !grad || !exp_avg || !exp_avg_sq)) return MVAE_ERR_INVALID_ARGUMENT;
  if (!aligned16(param) || !aligned16(grad) || !aligned16(exp_avg) || !aligned16(exp_avg_sq)) return MVAE_ERR_ALIGNMENT;
  if (radius && radius_lr != 0.f && !gradius) return MVAE_ERR_INVALID_ARGUMENT;
  OptParams q;
  memset(&q, 0, sizeof(q));
  q.n4 = n / 4;
  q.p = param;
  q.g = grad;
  q.m = exp_avg;
  q.v = exp_avg_sq;
  q.lr = lr;
  q.b1 = beta1;
  q.b2 = beta2;
  q.eps = eps;
  q.step_dev = step_dev;
  q.done = done_counter;
  q.radius = radius;
  q.gradius = gradius;
  q.radius_mask = radius_mask;
  q.radius_lr = radius_lr;
  q.C = C;
  q.n_targets = n_targets;
  for (int t = 0; t < n_targets; ++t) {
    if (!target_begin || !target_rows || !targets) return MVAE_ERR_INVALID_ARGUMENT;
    const mvae_planes& pl = targets[t];
    int rc = planes_ok(&pl, target_rows[t], pl.cols);
    if (rc != MVAE_OK) return rc;
    if ((target_begin[t] & 3) || (pl.cols & 3) || target_begin[t] < 0 ||
        target_begin[t] + (int64_t)target_rows[t] * pl.cols > n)
      return MVAE_ERR_ALIGNMENT;
    q.t[t].begin = target_begin[t];
    q.t[t].end = target_begin[t] + (int64_t)target_rows[t] * pl.cols;
    q.t[t].cols = pl.cols;
    q.t[t].ld = pl.ld;
    q.t[t].planes = pl.planes;
    q.t[t].base = pl.base;
    q.t[t].plane_stride = pl.planes > 1 ? pl.plane_stride : 0;
  }
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  const int64_t want = (q.n4 + 255) / 256;
  int grid = (int)(want < (int64_t)di.sm_count * 4 ? want : (int64_t)di.sm_count * 4);
  if (grid < 1) grid = 1;
  MVAE_CUDA_TRY(launch_pdl(opt_fused_kernel, dim3(grid), dim3(256), 0, as_stream(stream), q));
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}
