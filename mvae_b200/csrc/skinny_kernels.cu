// skinny_kernels.cu — the dense layers whose N or K is tiny (the latent heads, N = sum(n)+sum(l_n) ~ 12..60, and the
// first decoder layer, K = sum(d) ~ 8..34: SURVEY.md §2.2 K2 / K4, ffnn_vae.py:56, component.py:64,69) and their
// backward passes.  These are far below a tcgen05 tile (a 128 x N x 16 MMA needs N >= 16 and costs a fixed ~10 us of
// TMEM / TMA / mbarrier set-up per launch), move a few MB and do <= 40 MFLOP: they run on the CUDA cores in exact fp32
// FMA arithmetic, read fp32 or split-bf16 planes directly (no conversion kernels) and are HBM / latency bound.
//
//   rowdot   out[b, n]   = sum_k A[b, k] W(n, k) (+ bias[n])            n <= 16, K large   (heads fwd, dgrad into z)
//   expand   out[b, n]   = act(sum_k A[b, k] W(n, k) + bias[n])         k <= 64, N large   (fc_d0 fwd, dgrad into h)
//   wgrad    out(s, w)  += sum_b small[b, s] wide[b, w]                 s <= 16(+1), W large (wgrad of both layers)
#include <cuda_bf16.h>
#include <string.h>

#include "mvae_common.cuh"

namespace mvae {

constexpr int kSkMaxSmall = 64;  // widest "small" dimension served

struct WideSrc {          // a [B, cols] matrix given either as fp32 or as split-bf16 planes
  const float* f32;       // or nullptr
  int64_t ld_f32;
  const uint16_t* planes; // plane 0
  int64_t plane_stride;
  int ld_planes;
  int nplanes;
};

__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }

__device__ __forceinline__ float wide_load(const WideSrc& a, int64_t row, int col) {
  if (a.f32) return __ldg(a.f32 + row * a.ld_f32 + col);
  float v = 0.f;
  for (int p = 0; p < a.nplanes; ++p)
    v += bf16_bits_to_float(__ldg(a.planes + p * a.plane_stride + row * a.ld_planes + col));
  return v;
}

// ------------------------------------------------------------------------------------------------ rowdot
// One warp per row; lanes own 8-wide chunks of K (one 128-bit load per plane, or two for fp32), W (n <= 16 rows) is
// staged once per CTA in shared memory as [n][Kp] fp32 (zero padded to a multiple of 8), and a shuffle tree reduces the
// n partial sums.  Persistent grid: every warp walks many rows, so the weight staging is amortised.
__device__ __forceinline__ void load8(const WideSrc& a, int64_t row, int k0, float (&v)[8]) {
  if (a.f32) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(a.f32 + row * a.ld_f32 + k0));
    const float4 y = __ldg(reinterpret_cast<const float4*>(a.f32 + row * a.ld_f32 + k0) + 1);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
    return;
  }
  // all plane loads are issued before any is consumed (statically unrolled, predicated on the plane count)
  uint4 q[3];
#pragma unroll
  for (int p = 0; p < 3; ++p)
    q[p] = p < a.nplanes
               ? __ldg(reinterpret_cast<const uint4*>(a.planes + p * a.plane_stride + row * a.ld_planes + k0))
               : make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const uint32_t w[4] = {q[p].x, q[p].y, q[p].z, q[p].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[2 * j] += __uint_as_float(w[j] << 16);
      v[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
    }
  }
}

template <int NMAX>
__global__ void __launch_bounds__(256) rowdot_kernel(int64_t B, int K, int Kp, int N, WideSrc a,
                                                     const float* __restrict__ W, int64_t w_sn, int64_t w_sk,
                                                     const float* __restrict__ bias, float* __restrict__ out,
                                                     int64_t ld_out, int vec_ok) {
  extern __shared__ __align__(16) float sW[];  // [NMAX][Kp]
  for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
    float wv[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) wv[n] = (n < N && k < K) ? __ldg(W + n * w_sn + k * w_sk) : 0.f;  // NMAX loads in flight
#pragma unroll
    for (int n = 0; n < NMAX; ++n) sW[n * Kp + k] = wv[n];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int nchunk = Kp >> 3;
  for (int64_t b = warp; b < B; b += nwarps) {
    float acc[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
    for (int c = lane; c < nchunk; c += 32) {
      float v[8];
      if (vec_ok) {
        load8(a, b, 8 * c, v);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (8 * c + j < K) ? wide_load(a, b, 8 * c + j) : 0.f;
      }
#pragma unroll
      for (int n = 0; n < NMAX; ++n) {
        const float4 w0 = *reinterpret_cast<const float4*>(sW + n * Kp + 8 * c);
        const float4 w1 = *reinterpret_cast<const float4*>(sW + n * Kp + 8 * c + 4);
        float t = acc[n];
        t = fmaf(v[0], w0.x, t); t = fmaf(v[1], w0.y, t); t = fmaf(v[2], w0.z, t); t = fmaf(v[3], w0.w, t);
        t = fmaf(v[4], w1.x, t); t = fmaf(v[5], w1.y, t); t = fmaf(v[6], w1.z, t); t = fmaf(v[7], w1.w, t);
        acc[n] = t;
      }
    }
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n] = warp_sum(acc[n]);
    if (lane == 0) {
#pragma unroll
      for (int n = 0; n < NMAX; ++n)
        if (n < N) out[b * ld_out + n] = acc[n] + (bias ? __ldg(bias + n) : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------ expand
// Thread per PAIR of output columns (32-bit bf16x2 plane stores), CTA per group of rows; the K <= 64 inputs of a row
// are broadcast from shared memory, the two W columns of the thread live in registers; rows are processed sixteen at a
// time so that the mask loads of a group are in flight together.
enum { kActNone = 0, kActRelu = 1, kActMask = 2 };

template <int KMAX>
__global__ void __launch_bounds__(128) expand_kernel(int64_t B, int K, int N, const float* __restrict__ A, int64_t ld_a,
                                                     const float* __restrict__ W, int64_t w_sn, int64_t w_sk,
                                                     const float* __restrict__ bias, int act,
                                                     const uint16_t* __restrict__ mask, int64_t ld_mask,
                                                     uint16_t* __restrict__ op_base, int64_t op_stride, int op_ld,
                                                     int op_planes, float* __restrict__ out_f32, int64_t ld_out,
                                                     int rows_per_cta) {
  __shared__ float sA[32][KMAX];
  const int n0 = 2 * (blockIdx.y * blockDim.x + threadIdx.x);
  const bool ok0 = n0 < N, ok1 = n0 + 1 < N;
  float w0[KMAX], w1[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    w0[k] = (ok0 && k < K) ? __ldg(W + n0 * w_sn + k * w_sk) : 0.f;
    w1[k] = (ok1 && k < K) ? __ldg(W + (n0 + 1) * w_sn + k * w_sk) : 0.f;
  }
  const float b0 = (bias && ok0) ? __ldg(bias + n0) : 0.f;
  const float b1 = (bias && ok1) ? __ldg(bias + n0 + 1) : 0.f;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row1 = min(B, row0 + rows_per_cta);
  const bool pair_store = ok1 && ((op_ld & 1) == 0);
  for (int64_t r0 = row0; r0 < row1; r0 += 32) {
    const int nr = (int)min((int64_t)32, row1 - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * KMAX; i += blockDim.x) {
      const int r = i / KMAX, k = i - r * KMAX;
      sA[r][k] = k < K ? __ldg(A + (r0 + r) * ld_a + k) : 0.f;
    }
    __syncthreads();
    if (!ok0) continue;
    for (int rb = 0; rb < nr; rb += 16) {
      uint32_t mk[16];
      if (act == kActMask) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          mk[j] = 0;
          if (rb + j < nr) {
            const uint16_t* mp = mask + (r0 + rb + j) * ld_mask + n0;
            mk[j] = ok1 ? ((uint32_t)__ldg(mp) | ((uint32_t)__ldg(mp + 1) << 16)) : (uint32_t)__ldg(mp);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int r = rb + j;
        if (r >= nr) break;
        float a0 = b0, a1 = b1;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          const float x = sA[r][k];
          a0 = fmaf(x, w0[k], a0);
          a1 = fmaf(x, w1[k], a1);
        }
        const int64_t row = r0 + r;
        if (act == kActRelu) {
          a0 = fmaxf(a0, 0.f);
          a1 = fmaxf(a1, 0.f);
        } else if (act == kActMask) {
          const uint32_t lo = mk[j] & 0xFFFFu, hi = mk[j] >> 16;
          a0 = (((lo & 0x8000u) == 0) && ((lo & 0x7FFFu) != 0)) ? a0 : 0.f;
          a1 = (((hi & 0x8000u) == 0) && ((hi & 0x7FFFu) != 0)) ? a1 : 0.f;
        }
        if (out_f32) {
          out_f32[row * ld_out + n0] = a0;
          if (ok1) out_f32[row * ld_out + n0 + 1] = a1;
        }
        float r0f = a0, r1f = a1;
        for (int p = 0; p < op_planes; ++p) {
          const __nv_bfloat16 h0 = __float2bfloat16_rn(r0f), h1 = __float2bfloat16_rn(r1f);
          uint16_t* dst = op_base + p * op_stride + row * op_ld + n0;
          const uint16_t u0 = *reinterpret_cast<const uint16_t*>(&h0), u1 = *reinterpret_cast<const uint16_t*>(&h1);
          if (pair_store) {
            *reinterpret_cast<uint32_t*>(dst) = (uint32_t)u0 | ((uint32_t)u1 << 16);
          } else {
            dst[0] = u0;
            if (ok1) dst[1] = u1;
          }
          r0f -= __bfloat162float(h0);
          r1f -= __bfloat162float(h1);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ wgrad
// out(s, w) += sum_b small[b, s] * wide[b, w]; thread per wide column, CTA per (row chunk, column chunk); the small
// rows are broadcast from shared memory and the wide column is read eight rows at a time (loads in flight together);
// `small_ones` appends an implicit all-ones small column whose sums are the bias gradient of the wide side; a ones
// column of the wide side (w == col_split) is diverted to out_col.
template <int SMAX>
__global__ void __launch_bounds__(128) skinny_wgrad_kernel(int64_t B, int S, int Wd, const float* __restrict__ small,
                                                           int64_t ld_small, int small_ones, WideSrc wide,
                                                           float* __restrict__ out, int64_t out_ss, int64_t out_sw,
                                                           float* __restrict__ out_row, float* __restrict__ out_col,
                                                           int col_split, int rows_per_cta) {
  __shared__ float sS[32][SMAX];
  const int w = blockIdx.y * blockDim.x + threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t row1 = min(B, row0 + rows_per_cta);
  float acc[SMAX];
#pragma unroll
  for (int s = 0; s < SMAX; ++s) acc[s] = 0.f;
  float acc_one = 0.f;
  for (int64_t r0 = row0; r0 < row1; r0 += 32) {
    const int nr = (int)min((int64_t)32, row1 - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < nr * SMAX; i += blockDim.x) {
      const int r = i / SMAX, s = i - r * SMAX;
      sS[r][s] = s < S ? __ldg(small + (r0 + r) * ld_small + s) : 0.f;
    }
    __syncthreads();
    if (w >= Wd) continue;
    for (int rb = 0; rb < nr; rb += 8) {
      float v[8];
      if (wide.f32) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (rb + j < nr) ? __ldg(wide.f32 + (r0 + rb + j) * wide.ld_f32 + w) : 0.f;
      } else {
        // 8 rows x up to 3 planes of 16-bit loads, all in flight before the first use
        uint16_t raw[8][3];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int p = 0; p < 3; ++p)
            raw[j][p] = (p < wide.nplanes && rb + j < nr)
                            ? __ldg(wide.planes + p * wide.plane_stride + (r0 + rb + j) * wide.ld_planes + w)
                            : (uint16_t)0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          v[j] = bf16_bits_to_float(raw[j][0]) + bf16_bits_to_float(raw[j][1]) + bf16_bits_to_float(raw[j][2]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = (rb + j < nr) ? rb + j : 0;  // v[j] == 0 for rows past the end
#pragma unroll
        for (int s = 0; s < SMAX; ++s) acc[s] = fmaf(sS[r][s], v[j], acc[s]);
        acc_one += v[j];
      }
    }
  }
  if (w >= Wd) return;
  if (w == col_split) {
    if (out_col) {
#pragma unroll
      for (int s = 0; s < SMAX; ++s)
        if (s < S) atomicAdd(out_col + s, acc[s]);
    }
    return;
  }
#pragma unroll
  for (int s = 0; s < SMAX; ++s)
    if (s < S) atomicAdd(out + s * out_ss + w * out_sw, acc[s]);
  if (small_ones && out_row) atomicAdd(out_row + w, acc_one);
}

static int fill_wide(WideSrc* w, const float* f32, int64_t ld_f32, const mvae_planes* pl) {
  memset(w, 0, sizeof(*w));
  if (f32) {
    w->f32 = f32;
    w->ld_f32 = ld_f32;
    return MVAE_OK;
  }
  if (!pl || !pl->base || pl->planes < 1 || pl->planes > 3) return MVAE_ERR_INVALID_ARGUMENT;
  w->planes = pl->base;
  w->plane_stride = pl->planes > 1 ? pl->plane_stride : 0;
  w->ld_planes = pl->ld;
  w->nplanes = pl->planes;
  return MVAE_OK;
}

}  // namespace mvae

using namespace mvae;

extern "C" int mvae_skinny_rowdot(int64_t B, int32_t K, int32_t N, const float* a_f32, int64_t ld_a,
                                  const mvae_planes* a_planes, const float* W, int64_t w_stride_n, int64_t w_stride_k,
                                  const float* bias, float* out, int64_t ld_out, void* stream) {
  if (B < 0 || K < 1 || N < 1 || N > 64 || !W || !out || ld_out < N) return MVAE_ERR_INVALID_ARGUMENT;
  if (B == 0) return MVAE_OK;
  WideSrc a;
  int rc = fill_wide(&a, a_f32, ld_a, a_planes);
  if (rc != MVAE_OK) return rc;
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  const int64_t want = (B + 7) / 8;
  const int grid = (int)(want < (int64_t)di.sm_count * 2 ? want : (int64_t)di.sm_count * 2);
  cudaStream_t s = as_stream(stream);
  const int Kp = (K + 7) / 8 * 8;
  // 128-bit row loads need 16-byte aligned rows that are readable up to Kp (planes: ld is a multiple of 8 by contract)
  int vec_ok;
  if (a.f32) vec_ok = ((reinterpret_cast<uintptr_t>(a.f32) & 15) == 0) && (a.ld_f32 % 4 == 0) && (a.ld_f32 >= Kp);
  else vec_ok = ((reinterpret_cast<uintptr_t>(a.planes) & 15) == 0) && (a.ld_planes % 8 == 0) && (a.ld_planes >= Kp) &&
                (a.plane_stride % 8 == 0);
  // one launch per group of <= 16 outputs keeps the accumulators in registers
  for (int n0 = 0; n0 < N; n0 += 16) {
    const int nn = N - n0 < 16 ? N - n0 : 16;
    const size_t smem = (size_t)Kp * 16 * sizeof(float);
    if (smem > 200 * 1024) return MVAE_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
      MVAE_CUDA_TRY(cudaFuncSetAttribute(rowdot_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rowdot_kernel<16><<<grid, 256, smem, s>>>(B, K, Kp, nn, a, W + n0 * w_stride_n, w_stride_n, w_stride_k,
                                              bias ? bias + n0 : nullptr, out + n0, ld_out, vec_ok);
    MVAE_LAUNCH_CHECK();
  }
  return MVAE_OK;
}

extern "C" int mvae_skinny_expand(int64_t B, int32_t K, int32_t N, const float* a, int64_t ld_a, const float* W,
                                  int64_t w_stride_n, int64_t w_stride_k, const float* bias, int32_t act,
                                  const uint16_t* mask, int64_t ld_mask, const mvae_planes* out_planes, float* out_f32,
                                  int64_t ld_out, void* stream) {
  if (B < 0 || K < 1 || K > kSkMaxSmall || N < 1 || !a || !W || ld_a < K) return MVAE_ERR_INVALID_ARGUMENT;
  if (act < kActNone || act > kActMask || (act == kActMask && (!mask || ld_mask < N))) return MVAE_ERR_INVALID_ARGUMENT;
  if (!out_planes && !out_f32) return MVAE_ERR_INVALID_ARGUMENT;
  if (out_planes && (!out_planes->base || out_planes->ld < N || out_planes->planes < 1 || out_planes->planes > 3))
    return MVAE_ERR_INVALID_ARGUMENT;
  if (B == 0) return MVAE_OK;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  const int ny = (N + 255) / 256;  // 128 threads x 2 columns
  int64_t chunks = (int64_t)di.sm_count * 4 / ny;
  if (chunks < 1) chunks = 1;
  int rows_per_cta = (int)((B + chunks - 1) / chunks);
  rows_per_cta = (rows_per_cta + 31) / 32 * 32;
  dim3 grid((unsigned)((B + rows_per_cta - 1) / rows_per_cta), ny);
  uint16_t* ob = out_planes ? out_planes->base : nullptr;
  const int64_t os = out_planes && out_planes->planes > 1 ? out_planes->plane_stride : 0;
  const int old = out_planes ? out_planes->ld : 0, opl = out_planes ? out_planes->planes : 0;
  cudaStream_t s = as_stream(stream);
#define MVAE_EXPAND(KM)                                                                                              \
  expand_kernel<KM><<<grid, 128, 0, s>>>(B, K, N, a, ld_a, W, w_stride_n, w_stride_k, bias, act, mask, ld_mask, ob, \
                                         os, old, opl, out_f32, ld_out, rows_per_cta)
  if (K <= 8) MVAE_EXPAND(8);
  else if (K <= 16) MVAE_EXPAND(16);
  else if (K <= 32) MVAE_EXPAND(32);
  else MVAE_EXPAND(64);
#undef MVAE_EXPAND
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

extern "C" int mvae_skinny_wgrad(int64_t B, int32_t S, int32_t Wd, const float* small, int64_t ld_small,
                                 int32_t small_ones, const float* wide_f32, int64_t ld_wide,
                                 const mvae_planes* wide_planes, float* out, int64_t out_stride_s, int64_t out_stride_w,
                                 float* out_row, float* out_col, int32_t col_split, void* stream) {
  if (B < 0 || S < 1 || S > kSkMaxSmall || Wd < 1 || !small || ld_small < S || !out) return MVAE_ERR_INVALID_ARGUMENT;
  if (B == 0) return MVAE_OK;
  WideSrc wide;
  int rc = fill_wide(&wide, wide_f32, ld_wide, wide_planes);
  if (rc != MVAE_OK) return rc;
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  const int ny = (Wd + 127) / 128;
  int64_t chunks = (int64_t)di.sm_count * 2 / ny;
  if (chunks < 1) chunks = 1;
  int rows_per_cta = (int)((B + chunks - 1) / chunks);
  rows_per_cta = (rows_per_cta + 31) / 32 * 32;
  dim3 grid((unsigned)((B + rows_per_cta - 1) / rows_per_cta), ny);
  cudaStream_t s = as_stream(stream);
#define MVAE_WGRAD(SM)                                                                                           \
  skinny_wgrad_kernel<SM><<<grid, 128, 0, s>>>(B, S, Wd, small, ld_small, small_ones, wide, out, out_stride_s,   \
                                               out_stride_w, out_row, out_col, col_split, rows_per_cta)
  if (S <= 8) MVAE_WGRAD(8);
  else if (S <= 16) MVAE_WGRAD(16);
  else if (S <= 32) MVAE_WGRAD(32);
  else MVAE_WGRAD(64);
#undef MVAE_WGRAD
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}
