// conv_kernels.cu — the data movement around the tcgen05 GEMM that turns it into the convolutions of the reference's
// ConvolutionalVAE (mt/mvae/models/conv_vae.py:47-55: nn.Conv2d / nn.ConvTranspose2d, kernel 4, stride 2, padding 1).
//
// Activations live channels-last ("NHWC"): a [B, H, W, C] tensor is the row-major matrix [B*H*W, C], carried like every
// GEMM operand here as split-bf16 planes.  With the filter taps unrolled into the contraction dimension,
//
//   Conv2d            y[B*OH*OW, Co]      = im2col(x)[B*OH*OW, 16 Ci]  .  W[Co, (ky, kx, ci)]^T
//   ConvTranspose2d   y                   = col2im( x[B*H*W, Ci]  .  W[Ci, (ky, kx, co)] )
//
// and their gradients are the same two gathers around the same GEMM (the adjoint of one is the other):
//   mvae_conv_im2col   rows of 16 taps x C channels gathered from the (zero padded) source image — a pure copy of bf16
//                      planes (splitting into planes commutes with gathering), 16 bytes per thread
//   mvae_conv_col2im   every output pixel sums the (at most) four tap columns that reach it, + bias, relu / relu-mask,
//                      written as planes for the next GEMM (and / or fp32)
// plus the two layout changes at the ends of the convolutional stacks (the reference flattens [B, C, H, W] as
// (c, y, x): conv_vae.py:61,65,73,77) and column sums for the bias gradients of the transposed convolutions.
#include <cuda_bf16.h>
#include <string.h>

#include "mvae_common.cuh"

namespace mvae {

struct Im2colParams {
  const uint16_t* src;
  int64_t src_stride;
  int src_ld;
  uint16_t* dst;
  int64_t dst_stride;
  int dst_ld;
  int planes, B, H, W, C, ones_col;
};

// VEC: C % 8 == 0 -> one 16-byte copy per thread; else one element per thread.
template <bool VEC>
__global__ void __launch_bounds__(256) im2col_k4s2_kernel(const Im2colParams p) {
  const int OH = p.H >> 1, OW = p.W >> 1;
  const int64_t rows = (int64_t)p.B * OH * OW;
  const int per_tap = VEC ? (p.C >> 3) : p.C;
  const int64_t per_plane = rows * 16 * per_tap;
  const int64_t total = per_plane * p.planes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pl = (int)(i / per_plane);
    int64_t r = i - (int64_t)pl * per_plane;
    const int ch = (int)(r % per_tap);
    r /= per_tap;
    const int tap = (int)(r & 15);
    const int64_t m = r >> 4;
    const int ox = (int)(m % OW);
    const int64_t t = m / OW;
    const int oy = (int)(t % OH);
    const int64_t b = t / OH;
    const int iy = 2 * oy - 1 + (tap >> 2), ix = 2 * ox - 1 + (tap & 3);
    const bool in = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
    const uint16_t* s = p.src + (int64_t)pl * p.src_stride + ((b * p.H + iy) * p.W + ix) * (int64_t)p.src_ld;
    uint16_t* d = p.dst + (int64_t)pl * p.dst_stride + m * p.dst_ld + tap * p.C;
    if (VEC) {
      const uint4 v = in ? __ldg(reinterpret_cast<const uint4*>(s) + ch) : make_uint4(0u, 0u, 0u, 0u);
      reinterpret_cast<uint4*>(d)[ch] = v;
    } else {
      d[ch] = in ? __ldg(s + ch) : (uint16_t)0;
    }
  }
  if (p.ones_col) {  // column 16 C holds 1.0 (plane 0): read as a wgrad operand it yields the bias gradient
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * p.planes;
         i += (int64_t)gridDim.x * blockDim.x) {
      const int pl = (int)(i / rows);
      const int64_t m = i - (int64_t)pl * rows;
      p.dst[(int64_t)pl * p.dst_stride + m * p.dst_ld + 16 * p.C] = pl == 0 ? (uint16_t)0x3F80u : (uint16_t)0u;
    }
  }
}

struct Col2imParams {
  const float* cols;
  int64_t ld_cols;
  int B, H, W, C;  // H, W: the SMALL (column-side) image; the output is [B, 2H, 2W, C]
  const float* bias;
  int act;  // 0 none, 1 relu, 2 mask by (mask plane 0 > 0)
  const uint16_t* mask;
  int mask_ld;
  uint16_t* out;
  int64_t out_stride;
  int out_ld, out_planes;
  float* out_f32;
  int64_t ld_f32;
};

__device__ __forceinline__ bool bf16_positive(uint16_t v) { return ((v & 0x8000u) == 0) && ((v & 0x7FFFu) != 0); }

template <int V>  // V = 4: C % 4 == 0, float4 loads; V = 1: any C
__global__ void __launch_bounds__(256) col2im_k4s2_kernel(const Col2imParams p) {
  const int OH = 2 * p.H, OW = 2 * p.W;
  const int per_row = p.C / V;
  const int64_t total = (int64_t)p.B * OH * OW * per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % per_row) * V;
    const int64_t mo = i / per_row;
    const int ox = (int)(mo % OW);
    const int64_t t = mo / OW;
    const int oy = (int)(t % OH);
    const int64_t b = t / OH;
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = p.bias ? __ldg(p.bias + c + j) : 0.f;
    // output row oy receives tap ky of input row iy when oy = 2 iy - 1 + ky: ky has the parity of oy + 1
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ky = ((oy + 1) & 1) + 2 * a;
      const int iy = (oy + 1 - ky) >> 1;
      if (iy < 0 || iy >= p.H) continue;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int kx = ((ox + 1) & 1) + 2 * e;
        const int ix = (ox + 1 - kx) >> 1;
        if (ix < 0 || ix >= p.W) continue;
        const float* s = p.cols + ((b * p.H + iy) * p.W + ix) * p.ld_cols + (ky * 4 + kx) * p.C + c;
        if (V == 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(s));
          acc[0] += v.x;
          acc[1 % V] += v.y;
          acc[2 % V] += v.z;
          acc[3 % V] += v.w;
        } else {
          acc[0] += __ldg(s);
        }
      }
    }
    if (p.act == 1) {
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = fmaxf(acc[j], 0.f);
    } else if (p.act == 2) {
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (!bf16_positive(__ldg(p.mask + mo * p.mask_ld + c + j))) acc[j] = 0.f;
    }
    if (p.out_f32) {
#pragma unroll
      for (int j = 0; j < V; ++j) p.out_f32[mo * p.ld_f32 + c + j] = acc[j];
    }
    for (int pl = 0; pl < p.out_planes; ++pl) {
      uint16_t* d = p.out + (int64_t)pl * p.out_stride + mo * p.out_ld + c;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const __nv_bfloat16 h = __float2bfloat16_rn(acc[j]);
        d[j] = *reinterpret_cast<const uint16_t*>(&h);
        acc[j] -= __bfloat162float(h);
      }
    }
  }
}

// [B, C*S] rows in (c, s) order  <->  [B*S, C] rows ("NHWC"); T = element type (uint16_t: planes, float).
template <typename T>
__global__ void __launch_bounds__(256) permute_sc_kernel(const T* src, int64_t src_ld, int64_t src_stride, T* dst,
                                                         int64_t dst_ld, int64_t dst_stride, int planes, int64_t B,
                                                         int S, int C, int to_nhwc) {
  const int64_t per_plane = B * S * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_plane * planes;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int pl = (int)(i / per_plane);
    int64_t r = i - (int64_t)pl * per_plane;
    const T* sp = src + (int64_t)pl * src_stride;
    T* dp = dst + (int64_t)pl * dst_stride;
    if (to_nhwc) {  // destination order: b, s, c
      const int c = (int)(r % C);
      r /= C;
      const int s = (int)(r % S);
      const int64_t b = r / S;
      dp[(b * S + s) * dst_ld + c] = sp[b * src_ld + (int64_t)c * S + s];
    } else {  // destination order: b, c, s
      const int s = (int)(r % S);
      r /= S;
      const int c = (int)(r % C);
      const int64_t b = r / C;
      dp[b * dst_ld + (int64_t)c * S + s] = sp[(b * S + s) * src_ld + c];
    }
  }
}

// out[c] += sum_m src[m, c]   (src fp32, or the sum of its planes); C <= 1024
__global__ void __launch_bounds__(256) colsum_kernel(const float* f32, const uint16_t* planes, int64_t plane_stride,
                                                     int n_planes, int64_t M, int C, int64_t ld, int rows_per_block,
                                                     float* out) {
  __shared__ float s_acc[1024];
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_acc[c] = 0.f;
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  // a thread keeps one column (C <= 256: several row phases per block) so that its partial sum stays in a register
  const int c = threadIdx.x % C, phase = threadIdx.x / C, phases = max(1, (int)blockDim.x / C);
  if (C <= (int)blockDim.x) {
    if (phase < phases) {
      float acc = 0.f;
      for (int64_t r = r0 + phase; r < r1; r += phases) {
        if (f32) {
          acc += __ldg(f32 + r * ld + c);
        } else {
          for (int pl = 0; pl < n_planes; ++pl)
            acc += __uint_as_float((uint32_t)__ldg(planes + pl * plane_stride + r * ld + c) << 16);
        }
      }
      atomicAdd(&s_acc[c], acc);
    }
  } else {
    for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
      float acc = 0.f;
      for (int64_t r = r0; r < r1; ++r) {
        if (f32) {
          acc += __ldg(f32 + r * ld + cc);
        } else {
          for (int pl = 0; pl < n_planes; ++pl)
            acc += __uint_as_float((uint32_t)__ldg(planes + pl * plane_stride + r * ld + cc) << 16);
        }
      }
      s_acc[cc] = acc;
    }
  }
  __syncthreads();
  for (int cc = threadIdx.x; cc < C; cc += blockDim.x) atomicAdd(out + cc, s_acc[cc]);
}

static int grid_for(int64_t work, const DeviceInfo& di) {
  int64_t blocks = (work + 255) / 256;
  const int64_t cap = (int64_t)di.sm_count * 16;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

static bool planes_valid(const mvae_planes* p) {
  return p && p->base && p->planes >= 1 && p->planes <= 3 && p->rows >= 1 && p->cols >= 1 && p->ld >= p->cols &&
         (p->ld & 7) == 0 && (reinterpret_cast<uintptr_t>(p->base) & 15) == 0 &&
         (p->planes == 1 || ((p->plane_stride & 7) == 0 && p->plane_stride >= (int64_t)p->rows * p->ld));
}

}  // namespace mvae

using namespace mvae;

extern "C" int mvae_conv_im2col(const mvae_planes* src, int32_t B, int32_t H, int32_t W, int32_t C,
                                const mvae_planes* dst, int32_t ones_col, void* stream) {
  if (!planes_valid(src) || !planes_valid(dst) || B < 1 || H < 2 || W < 2 || (H & 1) || (W & 1) || C < 1)
    return MVAE_ERR_INVALID_ARGUMENT;
  const int64_t rows_in = (int64_t)B * H * W, rows_out = rows_in / 4;
  if (src->rows != rows_in || src->cols != C || dst->rows != rows_out || dst->cols != 16 * C ||
      dst->ld < 16 * C + (ones_col ? 1 : 0) || dst->planes > src->planes)
    return MVAE_ERR_INVALID_ARGUMENT;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  Im2colParams p;
  memset(&p, 0, sizeof(p));
  p.src = src->base;
  p.src_stride = src->planes > 1 ? src->plane_stride : 0;
  p.src_ld = src->ld;
  p.dst = dst->base;
  p.dst_stride = dst->planes > 1 ? dst->plane_stride : 0;
  p.dst_ld = dst->ld;
  p.planes = dst->planes;
  p.B = B;
  p.H = H;
  p.W = W;
  p.C = C;
  p.ones_col = ones_col ? 1 : 0;
  if (C % 8 == 0) {
    im2col_k4s2_kernel<true><<<grid_for(rows_out * 16 * (C / 8) * p.planes, di), 256, 0, as_stream(stream)>>>(p);
  } else {
    im2col_k4s2_kernel<false><<<grid_for(rows_out * 16 * C * p.planes, di), 256, 0, as_stream(stream)>>>(p);
  }
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

extern "C" int mvae_conv_col2im(const float* cols, int64_t ld_cols, int32_t B, int32_t H, int32_t W, int32_t C,
                                const float* bias, int32_t act, const mvae_planes* mask, const mvae_planes* out_planes,
                                float* out_f32, int64_t ld_out, void* stream) {
  if (!cols || B < 1 || H < 1 || W < 1 || C < 1 || ld_cols < 16 * (int64_t)C || act < 0 || act > 2 ||
      (!out_planes && !out_f32))
    return MVAE_ERR_INVALID_ARGUMENT;
  const int64_t rows_out = (int64_t)B * H * W * 4;
  if (act == 2 && (!planes_valid(mask) || mask->rows != rows_out || mask->cols != C)) return MVAE_ERR_INVALID_ARGUMENT;
  if (out_planes && (!planes_valid(out_planes) || out_planes->rows != rows_out || out_planes->cols != C))
    return MVAE_ERR_INVALID_ARGUMENT;
  if (out_f32 && ld_out < C) return MVAE_ERR_INVALID_ARGUMENT;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  Col2imParams p;
  memset(&p, 0, sizeof(p));
  p.cols = cols;
  p.ld_cols = ld_cols;
  p.B = B;
  p.H = H;
  p.W = W;
  p.C = C;
  p.bias = bias;
  p.act = act;
  if (act == 2) {
    p.mask = mask->base;
    p.mask_ld = mask->ld;
  }
  if (out_planes) {
    p.out = out_planes->base;
    p.out_stride = out_planes->planes > 1 ? out_planes->plane_stride : 0;
    p.out_ld = out_planes->ld;
    p.out_planes = out_planes->planes;
  }
  p.out_f32 = out_f32;
  p.ld_f32 = ld_out;
  const bool vec = (C % 4 == 0) && (ld_cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(cols) & 15) == 0);
  if (vec) col2im_k4s2_kernel<4><<<grid_for(rows_out * (C / 4), di), 256, 0, as_stream(stream)>>>(p);
  else col2im_k4s2_kernel<1><<<grid_for(rows_out * C, di), 256, 0, as_stream(stream)>>>(p);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

extern "C" int mvae_permute_sc(int32_t elem_bytes, const void* src, int64_t src_ld, int64_t src_plane_stride, void* dst,
                               int64_t dst_ld, int64_t dst_plane_stride, int32_t planes, int64_t B, int32_t S, int32_t C,
                               int32_t to_nhwc, void* stream) {
  if ((elem_bytes != 2 && elem_bytes != 4) || !src || !dst || planes < 1 || planes > 3 || B < 1 || S < 1 || C < 1)
    return MVAE_ERR_INVALID_ARGUMENT;
  const int64_t flat_ld = to_nhwc ? src_ld : dst_ld, nhwc_ld = to_nhwc ? dst_ld : src_ld;
  if (flat_ld < (int64_t)C * S || nhwc_ld < C) return MVAE_ERR_INVALID_ARGUMENT;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  const int grid = grid_for(B * S * C * planes, di);
  if (elem_bytes == 2)
    permute_sc_kernel<uint16_t><<<grid, 256, 0, as_stream(stream)>>>(
        static_cast<const uint16_t*>(src), src_ld, src_plane_stride, static_cast<uint16_t*>(dst), dst_ld,
        dst_plane_stride, planes, B, S, C, to_nhwc ? 1 : 0);
  else
    permute_sc_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(
        static_cast<const float*>(src), src_ld, src_plane_stride, static_cast<float*>(dst), dst_ld, dst_plane_stride,
        planes, B, S, C, to_nhwc ? 1 : 0);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

extern "C" int mvae_colsum(const float* src_f32, const mvae_planes* src_planes, int64_t M, int32_t C, int64_t ld,
                           float* out, void* stream) {
  if ((!src_f32 && !src_planes) || !out || M < 0 || C < 1 || C > 1024) return MVAE_ERR_INVALID_ARGUMENT;
  if (src_planes && (!planes_valid(src_planes) || src_planes->rows < M || src_planes->cols < C))
    return MVAE_ERR_INVALID_ARGUMENT;
  if (src_f32 && ld < C) return MVAE_ERR_INVALID_ARGUMENT;
  if (M == 0) return MVAE_OK;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  int rows_per_block = 256;
  while ((M + rows_per_block - 1) / rows_per_block > (int64_t)di.sm_count * 8) rows_per_block *= 2;
  const int grid = (int)((M + rows_per_block - 1) / rows_per_block);
  colsum_kernel<<<grid, 256, 0, as_stream(stream)>>>(
      src_f32, src_planes ? src_planes->base : nullptr,
      src_planes && src_planes->planes > 1 ? src_planes->plane_stride : 0, src_planes ? src_planes->planes : 0, M, C,
      src_f32 ? ld : (int64_t)src_planes->ld, rows_per_block, out);
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}
