// latent_impl.cuh — the whole latent block of the VAE in ONE kernel per direction (mvae_latent_forward /
// mvae_latent_backward, include/mvae_b200.h):
//
//   forward   h --fc_mean/fc_logvar--> ml --Component.encode, reparametrize, rsample, kl_loss--> z, kl --fc_d0, relu--> dd
//             (component.py:63-75, sampling_procedures.py:91-116,145-155, vae.py:69-80, ffnn_vae.py:56)
//   backward  gdd --fc_d0 dgrad--> gz --reverse sweep of the manifold chain--> gml --heads dgrad, relu mask--> gh
//             plus the weight / bias gradients of fc_d0 and of the heads, and dR
//
// The dense layers here are "skinny" (N = sum(n)+sum(l_n) ~ 12..60 outputs, K = sum(d) ~ 8..34 inputs): far below a
// tcgen05 tile, a few MB of traffic.  Done as separate launches they cost six kernels and five round trips of
// [B, P] / [B, Sd] / [B, H] intermediates through L2; fused, the per-sample intermediates (ml on the way back, gz, gml)
// never leave shared memory and the activations h / gdd are read exactly once.
//
// A CTA owns R = 16 (backward) or 8 (forward) consecutive rows.  Their rows of h (or gdd and h; fp32, written by the producing GEMM's epilogue
// for this kernel alone) are contiguous: one bulk asynchronous copy (TMA) lands them in shared memory, where they serve both access
// patterns — warp-per-row dot products (lanes along the hidden dimension, shuffle reduction) and thread-per-column
// outer products (rows unrolled in registers).  The manifold arithmetic is pm_math.cuh through dispatch_item, one warp
// per component, lane = row.  Weight gradients are accumulated per CTA in registers over its rows and leave as
// vector reductions (red.global.add.v2/.v4.f32).  All arithmetic is exact fp32 FMA.
#pragma once
#include <cuda_bf16.h>

#include "pm_item.cuh"

#ifndef MVAE_LAT_BWD
#error "define MVAE_LAT_BWD to 0 (forward) or 1 (backward) before including latent_impl.cuh"
#endif

#if !MVAE_LAT_BWD
#include <curand_kernel.h>
#endif

namespace mvae {

// rows per CTA.  Backward: 16 (8 doubles the per-CTA weight-gradient reductions: measured slower).  Forward has no
// reductions: 8 rows halve the serial chain of a CTA (one row per warp in the heads) and double the CTAs in flight.
#ifndef MVAE_LAT_ROWS_FWD
#define MVAE_LAT_ROWS_FWD 8
#endif
constexpr int kLatRows = MVAE_LAT_BWD ? 16 : MVAE_LAT_ROWS_FWD;
constexpr int kLatThreads = 256;  // 8 warps

struct LatParams {
  mvae_pm_desc desc;
  int64_t B;
  int H;                       // hidden width (multiple of 8)
  // h = relu(fc_e0(x)) as fp32 [B, h_ld] (forward: heads; backward: heads weight gradient and the relu mask)
  const float* h;
  int h_ld;
  const float* Wh;             // [P, H]
  const float* bh;             // [P]
  const float* Wd0;            // [H, Sd]
  const float* bd0;            // [H]
  const float* eps;
  const float* radius;
  // forward outputs
  float* ml;
  float* z;
  float* kl;
  uint16_t* dd;
  int64_t dd_stride;
  int dd_ld, dd_planes;
  uint32_t* flag;
  // backward
  const float* gdd;            // fp32 [B, gdd_ld]
  int gdd_ld;
  const float* ml_in;
  const float* z_in;
  float gkl;
  uint16_t* gh;
  int64_t gh_stride;
  int gh_ld, gh_planes;
  float* gWd0;
  float* gbd0;
  float* gWh;
  float* gbh;
  float* gradius;
  int zero_gml;
  // forward: the head of the train step taken along (mvae_latent_forward_ex): noise drawn here, zero fills
  int draw_eps, n_zero;
  unsigned long long seed;
  const unsigned long long* counter_dev;
  float* zptr[4];
  int64_t zn[4];
  // diagnostics (mvae_debug_latent; null / 0 in production): per-CTA %globaltimer stamps at the phase boundaries, and
  // bit 0 of debug_flags drops the global reductions (timing experiments only: the gradients are then wrong)
  unsigned long long* stamps;
  int debug_flags;
  // shared-memory layout (bytes from the start of dynamic shared memory), host-computed
  int off_a, off_b, off_ml, off_part, off_eps, off_z, off_kl, off_gz, off_gml, off_bar;
};

// diagnostics: set by mvae_debug_latent (api.cu), copied into LatParams by launch_latent
extern unsigned long long* g_lat_stamps;
extern int g_lat_debug_flags;
constexpr int kLatStampSlots = 16;
__device__ __forceinline__ void lat_stamp(const LatParams& p, int i, bool sync) {
  if (p.stamps == nullptr) return;
  if (sync) __syncthreads();
  if (threadIdx.x == 0) {
    // slots 0-7: %globaltimer (comparable across CTAs, ~0.26 us resolution); slots 8-15: the SM's cycle counter
    // (thread 0's own progress inside a phase; slot 15 = reference taken together with slot 2)
    unsigned long long t;
    if (i < 8) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    else t = (unsigned long long)clock64();
    p.stamps[(size_t)blockIdx.x * kLatStampSlots + i] = t;
  }
}

#if !MVAE_LAT_BWD
// one Philox4x32-10 counter block -> four standard normals (Box-Muller), exactly as step_prologue_kernel draws them;
// not inlined: the generator's state would otherwise sit in the registers of the whole kernel
__device__ __noinline__ float4 lat_normal4(unsigned long long seed, unsigned long long block, unsigned long long offset) {
  curandStatePhilox4_32_10_t st;
  curand_init(seed, block, offset, &st);
  return curand_normal4(&st);
}
#endif

// Pull a weight matrix (a few KB every CTA reads in full) towards the SM while the CTA's activation rows are still in
// flight.  In a train step the weights were last touched by the previous step's optimizer: read on demand, the first
// touch of every line is a DRAM round trip inside the dependent chains of the dot products (scripts/step_timeline.py:
// heads 11.0 us in the captured step against 4.6 us with the weights warm in L2 — without this prefetch).
__device__ __forceinline__ void lat_prefetch_l1(const float* base, int bytes) {
  const char* b = reinterpret_cast<const char*>(base);
  for (int off = threadIdx.x * 128; off < bytes; off += kLatThreads * 128)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(b + off));
}

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// 8 consecutive fp32 columns of one row held in shared memory
__device__ __forceinline__ void row_load8(const float* s, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
// split a column pair into bf16 planes and store it (32-bit stores, coalesced across the warp)
__device__ __forceinline__ void planes_store2(uint16_t* dst, int64_t stride, int planes, float a, float b) {
  for (int p = 0; p < planes; ++p) {
    const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
    const uint32_t ua = *reinterpret_cast<const uint16_t*>(&ha), ub = *reinterpret_cast<const uint16_t*>(&hb);
    *reinterpret_cast<uint32_t*>(dst + p * stride) = ua | (ub << 16);
    a -= __bfloat162float(ha);
    b -= __bfloat162float(hb);
  }
}
__device__ __forceinline__ void red_add2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Access pattern of the skinny dot products.  Both kernels were bound by L1 wavefronts, not by arithmetic
// (scripts/latent_phases.py: 16 of the backward kernel's 31 us sat in gz, 8 of the forward kernel's 17 us in the heads):
// with a lane owning 8 consecutive hidden units, a warp's load of W touched 8 (W[n][k]) or 32 (W[h][j]) different
// 128-byte lines per instruction, and every row of the CTA re-read all of W.  Now the lanes of a warp own NEIGHBOURING
// pieces (one float4 of a W[n][.] row / one W[h][.] row each), and a warp takes NR = 2 rows at once so that every
// W value it loads is used twice.

// Sum 32 per-lane values across the warp in one go: returns, in lane L, the warp-wide sum of v[OFF + L].  Every round
// a lane hands one half of what it still carries to its partner and keeps the other: 31 shuffles in 5 dependent
// rounds (32 separate butterflies: 160 shuffles — and, one per conditional block, 32 x 5 exposed shuffle latencies).
template <int HALF>
__device__ __forceinline__ void warp_fold(float (&t)[16], int lane) {
  const bool up = (lane & HALF) != 0;
#pragma unroll
  for (int k = 0; k < HALF; ++k) {
    const float send = up ? t[k] : t[k + HALF], keep = up ? t[k + HALF] : t[k];
    t[k] = keep + __shfl_xor_sync(0xffffffffu, send, HALF);
  }
}
template <int OFF, int TOTAL>
__device__ __forceinline__ float warp_reduce32(const float (&v)[TOTAL]) {
  static_assert(OFF + 32 <= TOTAL, "slice out of range");
  const int lane = threadIdx.x & 31;
  const bool up = (lane & 16) != 0;
  float t[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float send = up ? v[OFF + k] : v[OFF + k + 16], keep = up ? v[OFF + k + 16] : v[OFF + k];
    t[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  warp_fold<8>(t, lane);
  warp_fold<4>(t, lane);
  warp_fold<2>(t, lane);
  warp_fold<1>(t, lane);
  return t[0];
}

// Per-lane partial sums acc[i * NMAX + n] of <row_i[4f..4f+3], W[n][4f..4f+3]> over the float4 columns f in [f0, f1)
// (W row-major [N, H]); the caller reduces them over the warp.
template <int NMAX, int NR>
__device__ __forceinline__ void rows_dot_WnK(const float* const* srow, int f0, int f1, int H,
                                             const float* __restrict__ W, int N, float (&acc)[NR * NMAX]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NR * NMAX; ++i) acc[i] = 0.f;
  for (int f = f0 + lane; f < f1; f += 32) {
    float4 v[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) v[i] = *reinterpret_cast<const float4*>(srow[i] + 4 * f);
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * H) + f);
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          float t = acc[i * NMAX + n];
          t = fmaf(v[i].x, w.x, t); t = fmaf(v[i].y, w.y, t); t = fmaf(v[i].z, w.z, t); t = fmaf(v[i].w, w.w, t);
          acc[i * NMAX + n] = t;
        }
      }
  }
}

// Per-lane partial sums acc[i * JMAX + j] of row_i[h] * W[h][j] over h = lane, lane + 32, ...  (W row-major [H, J], J
// small): the dgrad of fc_d0 into z.  Neighbouring lanes read neighbouring W rows.
template <int JMAX, int NR>
__device__ __forceinline__ void rows_dot_WKn(const float* const* srow, int H, const float* __restrict__ W, int J,
                                             float (&acc)[NR * JMAX]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NR * JMAX; ++i) acc[i] = 0.f;
  const bool vec = (J & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0;
#pragma unroll 2
  for (int h = lane; h < H; h += 32) {
    float g[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) g[i] = srow[i][h];
    const float* wr = W + (int64_t)h * J;
    if (vec) {
#pragma unroll
      for (int j = 0; j < JMAX; j += 4)
        if (j < J) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(wr + j));
#pragma unroll
          for (int i = 0; i < NR; ++i) {
            float* a = acc + i * JMAX + j;
            a[0] = fmaf(g[i], w.x, a[0]);
            a[1] = fmaf(g[i], w.y, a[1]);
            a[2] = fmaf(g[i], w.z, a[2]);
            a[3] = fmaf(g[i], w.w, a[3]);
          }
        }
    } else {
#pragma unroll
      for (int j = 0; j < JMAX; ++j)
        if (j < J) {
          const float w = __ldg(wr + j);
#pragma unroll
          for (int i = 0; i < NR; ++i) acc[i * JMAX + j] = fmaf(g[i], w, acc[i * JMAX + j]);
        }
    }
  }
}

#if !MVAE_LAT_BWD
// ================================================= forward =================================================
template <int MAXN, int SMAX>
__global__ void __launch_bounds__(kLatThreads) latent_forward_kernel(const __grid_constant__ LatParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int R = kLatRows;
  const int C = p.desc.C, P = p.desc.ld_ml, Sn = p.desc.ld_eps, Sd = p.desc.ld_z, H = p.H;
  ItemInfo* info = reinterpret_cast<ItemInfo*>(smem);
  float* sH = reinterpret_cast<float*>(smem + p.off_a);
  float* sML = reinterpret_cast<float*>(smem + p.off_ml);
  float* sEPS = reinterpret_cast<float*>(smem + p.off_eps);
  float* sZ = reinterpret_cast<float*>(smem + p.off_z);
  float* sKL = reinterpret_cast<float*>(smem + p.off_kl);
  const uint32_t bar = pm_smem_u32(smem + p.off_bar);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * R;
  const int rows = (int)min((int64_t)R, p.B - row0);

  if (tid == 0) {
    pm_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  lat_stamp(p, 0, false);
  pdl_wait();  // the radii (previous optimizer step) and h (previous GEMM) are read from here on
  lat_stamp(p, 1, false);
  stage_items(info, p.desc, p.radius);
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)(rows * p.h_ld) * 4u;
    pm_mbar_expect_tx(bar, bytes);
    pm_bulk_g2s(pm_smem_u32(sH), p.h + row0 * p.h_ld, bytes, bar);
  }
  lat_prefetch_l1(p.Wh, P * H * 4);
  lat_prefetch_l1(p.Wd0, H * Sd * 4);
  lat_prefetch_l1(p.bh, P * 4);
  lat_prefetch_l1(p.bd0, H * 4);
  if (p.draw_eps) {
    // eps ~ N(0, I) for this CTA's rows: the Philox stream of step_prologue_kernel (input_kernels.cu) element for
    // element — counter block i serves the flat elements 4i .. 4i+3 of eps [B, Sn]; R * Sn is a multiple of 4, so a
    // CTA owns whole blocks.  Kept in shared memory for the chain below and written out for the backward pass.
    const unsigned long long step = p.counter_dev ? *p.counter_dev : 0ull;
    const int64_t n_eps = p.B * Sn, g0 = row0 * Sn >> 2;
    const int ng = (rows * Sn + 3) >> 2;
    float* eps_out = const_cast<float*>(p.eps);
    for (int i = tid; i < ng; i += blockDim.x) {
      const float4 v = lat_normal4(p.seed, (unsigned long long)(g0 + i), step * 4ull);
      const float t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4 * (g0 + i) + j < n_eps) {
          sEPS[4 * i + j] = t[j];
          eps_out[4 * (g0 + i) + j] = t[j];
        }
    }
  } else {
    for (int i = tid; i < rows * Sn; i += blockDim.x) sEPS[i] = __ldg(p.eps + row0 * Sn + i);
  }
  // zero fills of the step (gradient bucket, reconstruction row sums), spread over the grid
  for (int s = 0; s < p.n_zero; ++s) {
    float* q = p.zptr[s];
    const int64_t n = p.zn[s], gt = (int64_t)blockIdx.x * blockDim.x + tid, nth = (int64_t)gridDim.x * blockDim.x;
    if ((reinterpret_cast<uintptr_t>(q) & 15) == 0) {
      const int64_t n4 = n >> 2;
      for (int64_t i = gt; i < n4; i += nth) reinterpret_cast<float4*>(q)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int64_t i = 4 * n4 + gt; i < n; i += nth) q[i] = 0.f;
    } else {
      for (int64_t i = gt; i < n; i += nth) q[i] = 0.f;
    }
  }
  __syncthreads();
  pm_mbar_wait(bar, 0);
  lat_stamp(p, 2, false);
  lat_stamp(p, 15, false);

  // ---- heads: ml[r][n] = <h[r], Wh[n]> + bh[n] ----
  // A warp takes the row pair (pr, pr + R/2) and, with R = 8, one half of the hidden dimension (8 warps = 4 pairs x
  // 2 halves); the halves meet in shared memory (fixed order: the result does not depend on scheduling).
  {
    constexpr int kWarps = kLatThreads / 32, PAIRS = R / 2, KS = kWarps / PAIRS;
    static_assert(R % 2 == 0 && kWarps % PAIRS == 0 && (KS == 1 || KS == 2), "row pairs x K halves must tile the warps");
    constexpr int NR = SMAX <= 16 ? 2 : 1;  // wider head blocks: one row at a time (registers)
    float* sPart = reinterpret_cast<float*>(smem + p.off_part);
    const int pr = warp % PAIRS, ks = warp / PAIRS;
    const int nf = H >> 2, f0 = ks * (nf / KS), f1 = ks == KS - 1 ? nf : f0 + nf / KS;
    float* dst = ks == 0 ? sML : sPart;
#pragma unroll
    for (int i0 = 0; i0 < 2; i0 += NR) {
      const float* srow[NR];
#pragma unroll
      for (int i = 0; i < NR; ++i) srow[i] = sH + min(pr + (i0 + i) * PAIRS, rows - 1) * p.h_ld;
      float acc[NR * SMAX];
      rows_dot_WnK<SMAX, NR>(srow, f0, f1, H, p.Wh, P, acc);
      lat_stamp(p, 8, false);   // (thread 0's warp) partial dot products done
      // lane L ends up with the warp's sum number L (and 32 + L): one store per lane
#pragma unroll
      for (int off = 0; off < NR * SMAX; off += 32) {
        const float sum = off == 0 ? warp_reduce32<0>(acc) : warp_reduce32<(NR * SMAX > 32 ? 32 : 0)>(acc);
        const int idx = off + lane, n = idx % SMAX, r = pr + (i0 + idx / SMAX) * PAIRS;
        if (n < P && r < rows) dst[r * P + n] = sum + (ks == 0 ? __ldg(p.bh + n) : 0.f);
      }
    }
    lat_stamp(p, 9, false);     // reduced and stored
    if (KS > 1) {
      __syncthreads();
      lat_stamp(p, 10, false);  // all warps there
      for (int i = tid; i < rows * P; i += blockDim.x) sML[i] += sPart[i];
    }
  }
  __syncthreads();
  lat_stamp(p, 3, false);

  // ---- manifold chain: one warp per component, lane = row ----
  bool finite = true;
  const bool check = p.flag != nullptr;
  for (int ci = warp; ci < C; ci += kLatThreads / 32) {
    if (lane < rows) {
      const ItemInfo c = info[ci];
      float unused = 0.f;
      finite &= dispatch_item<false, MAXN, false>(c, sML + lane * P, sEPS + lane * Sn, sZ + lane * Sd,
                                                  sKL + lane * C + ci, nullptr, nullptr, nullptr, 0.f, nullptr,
                                                  &unused, check);
    }
  }
  if (check) {
    const unsigned bad = __ballot_sync(0xffffffffu, !finite);
    if (bad && lane == 0) atomicOr(p.flag, 1u);
  }
  __syncthreads();
  lat_stamp(p, 4, false);
  for (int i = tid; i < rows * P; i += blockDim.x) p.ml[row0 * P + i] = sML[i];
  for (int i = tid; i < rows * Sd; i += blockDim.x) p.z[row0 * Sd + i] = sZ[i];
  for (int i = tid; i < rows * C; i += blockDim.x) p.kl[row0 * C + i] = sKL[i];

  pdl_launch_dependents();
  // ---- fc_d0 + relu -> planes: thread per column pair, the rows unrolled in registers ----
  const bool vecd = (Sd & 3) == 0 && (reinterpret_cast<uintptr_t>(p.Wd0) & 15) == 0;
  for (int n = 2 * tid; n < H; n += 2 * kLatThreads) {
    float a0[R], a1[R];
    const float b0 = __ldg(p.bd0 + n), b1 = __ldg(p.bd0 + n + 1);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      a0[r] = b0;
      a1[r] = b1;
    }
    if (vecd) {
      // Sd % 4 == 0: the weights of the column pair as 128-bit loads, z four coordinates at a time (same order of
      // the fused multiply-adds as the scalar loop)
      for (int k = 0; k < Sd; k += 4) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.Wd0 + (int64_t)n * Sd + k));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.Wd0 + (int64_t)(n + 1) * Sd + k));
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 zz = *reinterpret_cast<const float4*>(sZ + r * Sd + k);
          a0[r] = fmaf(zz.x, w0.x, a0[r]); a0[r] = fmaf(zz.y, w0.y, a0[r]);
          a0[r] = fmaf(zz.z, w0.z, a0[r]); a0[r] = fmaf(zz.w, w0.w, a0[r]);
          a1[r] = fmaf(zz.x, w1.x, a1[r]); a1[r] = fmaf(zz.y, w1.y, a1[r]);
          a1[r] = fmaf(zz.z, w1.z, a1[r]); a1[r] = fmaf(zz.w, w1.w, a1[r]);
        }
      }
    } else {
      for (int k = 0; k < Sd; ++k) {
        const float w0 = __ldg(p.Wd0 + (int64_t)n * Sd + k), w1 = __ldg(p.Wd0 + (int64_t)(n + 1) * Sd + k);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float zz = sZ[r * Sd + k];
          a0[r] = fmaf(zz, w0, a0[r]);
          a1[r] = fmaf(zz, w1, a1[r]);
        }
      }
    }
    if (n == 0) lat_stamp(p, 11, false);  // thread 0: fc_d0 column pair accumulated
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (r < rows)
        planes_store2(p.dd + (row0 + r) * p.dd_ld + n, p.dd_stride, p.dd_planes, fmaxf(a0[r], 0.f), fmaxf(a1[r], 0.f));
  }
  lat_stamp(p, 5, true);
}

#else
// ================================================= backward =================================================
template <int MAXN, int SMAX>
__global__ void __launch_bounds__(kLatThreads) latent_backward_kernel(const __grid_constant__ LatParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int R = kLatRows;
  const int C = p.desc.C, P = p.desc.ld_ml, Sn = p.desc.ld_eps, Sd = p.desc.ld_z, H = p.H;
  ItemInfo* info = reinterpret_cast<ItemInfo*>(smem);
  float* sG = reinterpret_cast<float*>(smem + p.off_a);
  float* sH = reinterpret_cast<float*>(smem + p.off_b);
  float* sML = reinterpret_cast<float*>(smem + p.off_ml);
  float* sEPS = reinterpret_cast<float*>(smem + p.off_eps);
  float* sZ = reinterpret_cast<float*>(smem + p.off_z);
  float* sGZ = reinterpret_cast<float*>(smem + p.off_gz);
  float* sGML = reinterpret_cast<float*>(smem + p.off_gml);
  const uint32_t bar = pm_smem_u32(smem + p.off_bar);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * R;
  const int rows = (int)min((int64_t)R, p.B - row0);

  if (tid == 0) {
    pm_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  lat_stamp(p, 0, false);
  pdl_wait();
  lat_stamp(p, 1, false);
  stage_items(info, p.desc, p.radius);
  if (tid == 0) {
    const uint32_t gb = (uint32_t)(rows * p.gdd_ld) * 4u, hb = (uint32_t)(rows * p.h_ld) * 4u;
    pm_mbar_expect_tx(bar, gb + hb);
    pm_bulk_g2s(pm_smem_u32(sG), p.gdd + row0 * p.gdd_ld, gb, bar);
    pm_bulk_g2s(pm_smem_u32(sH), p.h + row0 * p.h_ld, hb, bar);
  }
  lat_prefetch_l1(p.Wd0, H * Sd * 4);
  lat_prefetch_l1(p.Wh, P * H * 4);
  for (int i = tid; i < R * P; i += blockDim.x) {
    sML[i] = i < rows * P ? __ldg(p.ml_in + row0 * P + i) : 0.f;
    sGML[i] = 0.f;  // rows past the end and columns no component owns contribute nothing below
  }
  for (int i = tid; i < rows * Sn; i += blockDim.x) sEPS[i] = __ldg(p.eps + row0 * Sn + i);
  for (int i = tid; i < R * Sd; i += blockDim.x) sZ[i] = i < rows * Sd ? __ldg(p.z_in + row0 * Sd + i) : 0.f;
  __syncthreads();
  pm_mbar_wait(bar, 0);
  lat_stamp(p, 2, false);
  lat_stamp(p, 15, false);
  const bool nored = (p.debug_flags & 1) != 0;

  // ---- gz[r][j] = sum_h gdd[r][h] Wd0[h][j]; a warp takes the row pair (warp, warp + R/2) ----
  {
    constexpr int kWarps = kLatThreads / 32, PAIRS = R / 2;
    constexpr int NR = SMAX <= 16 ? 2 : 1;
    for (int pr = warp; pr < PAIRS; pr += kWarps) {
#pragma unroll
      for (int i0 = 0; i0 < 2; i0 += NR) {
        const float* srow[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) srow[i] = sG + min(pr + (i0 + i) * PAIRS, rows - 1) * p.gdd_ld;
        float acc[NR * SMAX];
        rows_dot_WKn<SMAX, NR>(srow, H, p.Wd0, Sd, acc);
        lat_stamp(p, 8, false);  // (thread 0's warp) partial dot products done
#pragma unroll
        for (int off = 0; off < NR * SMAX; off += 32) {
          const float sum = off == 0 ? warp_reduce32<0>(acc) : warp_reduce32<(NR * SMAX > 32 ? 32 : 0)>(acc);
          const int idx = off + lane, j = idx % SMAX, r = pr + (i0 + idx / SMAX) * PAIRS;
          if (j < Sd && r < rows) sGZ[r * Sd + j] = sum;
        }
      }
    }
  }
  lat_stamp(p, 3, true);
  // ---- fc_d0 weight / bias gradient: gWd0[n][j] += sum_r gdd[r][n] z[r][j]; thread per column pair ----
  for (int n = 2 * tid; n < H; n += 2 * kLatThreads) {
    float g0[R], g1[R];
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 g = make_float2(0.f, 0.f);
      if (r < rows) g = *reinterpret_cast<const float2*>(sG + r * p.gdd_ld + n);
      g0[r] = g.x;
      g1[r] = g.y;
      s0 += g.x;
      s1 += g.y;
    }
    if (p.gbd0 && !nored) red_add2(p.gbd0 + n, s0, s1);
    if ((Sd & 3) == 0) {
      for (int j = 0; j < Sd; j += 4) {
        float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 zz = *reinterpret_cast<const float4*>(sZ + r * Sd + j);
          a[0] = fmaf(g0[r], zz.x, a[0]); a[1] = fmaf(g0[r], zz.y, a[1]); a[2] = fmaf(g0[r], zz.z, a[2]); a[3] = fmaf(g0[r], zz.w, a[3]);
          b[0] = fmaf(g1[r], zz.x, b[0]); b[1] = fmaf(g1[r], zz.y, b[1]); b[2] = fmaf(g1[r], zz.z, b[2]); b[3] = fmaf(g1[r], zz.w, b[3]);
        }
        if (!nored || a[0] + b[0] == 12345.f) {
          red_add4(p.gWd0 + (int64_t)n * Sd + j, a[0], a[1], a[2], a[3]);
          red_add4(p.gWd0 + (int64_t)(n + 1) * Sd + j, b[0], b[1], b[2], b[3]);
        }
      }
    } else if (Sd <= 8) {
      // narrow latent spaces whose width is not a multiple of 4 (h2: 3, p2: 2 coordinates): the 2 Sd gradients of the
      // column pair (n, n + 1) are contiguous in gWd0 and start on an 8-byte boundary (n is even): Sd two-float
      // reductions (one four-float reduction for Sd = 2) instead of 2 Sd scalar atomics
      float c[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = 0.f, b = 0.f;
        if (j < Sd) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float zz = sZ[r * Sd + j];
            a = fmaf(g0[r], zz, a);
            b = fmaf(g1[r], zz, b);
          }
        }
        c[j] = a;
        c[8 + j] = b;
      }
      float* dst = p.gWd0 + (int64_t)n * Sd;
      if (Sd == 2) {
        red_add4(dst, c[0], c[1], c[8], c[9]);
      } else {
        // c2[i] = i < Sd ? a[i] : b[i - Sd], reduced in pairs (compile-time indices only: no local memory)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < Sd) {
            const int i0 = 2 * i, i1 = 2 * i + 1;
            float v0 = 0.f, v1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (j < Sd) {
                if (i0 == j) v0 = c[j];
                if (i0 == Sd + j) v0 = c[8 + j];
                if (i1 == j) v1 = c[j];
                if (i1 == Sd + j) v1 = c[8 + j];
              }
            }
            red_add2(dst + i0, v0, v1);
          }
        }
      }
    } else {
      for (int j = 0; j < Sd; ++j) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float zz = sZ[r * Sd + j];
          a = fmaf(g0[r], zz, a);
          b = fmaf(g1[r], zz, b);
        }
        atomicAdd(p.gWd0 + (int64_t)n * Sd + j, a);
        atomicAdd(p.gWd0 + (int64_t)(n + 1) * Sd + j, b);
      }
    }
  }
  __syncthreads();
  lat_stamp(p, 4, false);

  // ---- reverse sweep of the manifold chain: one warp per component, lane = row ----
  for (int ci = warp; ci < C; ci += kLatThreads / 32) {
    float gR = 0.f;
    const ItemInfo c = info[ci];
    if (lane < rows)
      dispatch_item<true, MAXN, false>(c, sML + lane * P, sEPS + lane * Sn, nullptr, nullptr, nullptr, nullptr,
                                       sGZ + lane * Sd, p.gkl, sGML + lane * P, &gR, false);
    if (p.gradius) {
      gR = warp_sum(gR) * c.dfac;
      if (lane == 0 && gR != 0.f) atomicAdd(p.gradius + ci, gR);
    }
  }
  __syncthreads();
  lat_stamp(p, 5, false);

  pdl_launch_dependents();
  // ---- heads: gWh[q][k] += sum_r gml[r][q] h[r][k];  gh[r][k] = (sum_q gml[r][q] Wh[q][k]) 1[h[r][k] > 0] ----
  if (p.gbh)
    for (int q = tid; q < P; q += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) s += sGML[r * P + q];
      atomicAdd(p.gbh + q, s);
    }
  // (H % 4 == 0 and a 16-byte aligned gWh: lanes 2i, 2i+1 cover 4 consecutive floats; the loop bound is warp-uniform
  // up to the last pair, whose partner then holds zeros)
  const bool pair4 = (H & 3) == 0 && (reinterpret_cast<uintptr_t>(p.gWh) & 15) == 0;
  for (int k0 = 0; k0 < H; k0 += 2 * kLatThreads) {
    const int k = k0 + 2 * tid;
    const bool live = k < H;
    float h0[R], h1[R], y0[R], y1[R];
    uint32_t mask = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float2 hv = make_float2(0.f, 0.f);
      if (r < rows && live) {
        hv = *reinterpret_cast<const float2*>(sH + r * p.h_ld + k);
        mask |= (uint32_t)(hv.x > 0.f) << (2 * r);  // relu'(h)
        mask |= (uint32_t)(hv.y > 0.f) << (2 * r + 1);
      }
      h0[r] = hv.x;
      h1[r] = hv.y;
      y0[r] = 0.f;
      y1[r] = 0.f;
    }
    // four head rows per pass: their Wh loads are issued together (one after the other they were a chain of L2
    // latencies), and with P % 4 == 0 one 128-bit shared-memory load brings a row's four gml values
    const bool q4 = (P & 3) == 0;
    lat_stamp(p, 9, false);  // thread 0: h columns in registers
    for (int q0 = 0; q0 < P; q0 += 4) {
      if (q0 == 4) lat_stamp(p, 10, false);  // thread 0: first pass of four head rows done
      float2 w[4];
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        w[u] = (live && q0 + u < P) ? __ldg(reinterpret_cast<const float2*>(p.Wh + (int64_t)(q0 + u) * H + k))
                                    : make_float2(0.f, 0.f);
        a[u] = 0.f;
        b[u] = 0.f;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float g[4];
        if (q4) {
          const float4 gg = *reinterpret_cast<const float4*>(sGML + r * P + q0);
          g[0] = gg.x; g[1] = gg.y; g[2] = gg.z; g[3] = gg.w;
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) g[u] = q0 + u < P ? sGML[r * P + q0 + u] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          y0[r] = fmaf(g[u], w[u].x, y0[r]);
          y1[r] = fmaf(g[u], w[u].y, y1[r]);
          a[u] = fmaf(g[u], h0[r], a[u]);
          b[u] = fmaf(g[u], h1[r], b[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (q0 + u < P) {
          // adjacent lanes own adjacent column pairs: the even lane issues ONE 128-bit reduction for both
          const float a2 = __shfl_down_sync(0xffffffffu, a[u], 1), b2 = __shfl_down_sync(0xffffffffu, b[u], 1);
          float* dstw = p.gWh + (int64_t)(q0 + u) * H + k;
          if (nored && a[u] + b2 != 12345.f) {
            // timing experiment: no reduction (the comparison keeps the arithmetic alive)
          } else if (pair4) {
            if ((lane & 1) == 0 && live) red_add4(dstw, a[u], b[u], a2, b2);
          } else if (live) {
            red_add2(dstw, a[u], b[u]);
          }
        }
    }
    lat_stamp(p, 6, true);
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (r < rows && live)
        planes_store2(p.gh + (row0 + r) * p.gh_ld + k, p.gh_stride, p.gh_planes, ((mask >> (2 * r)) & 1u) ? y0[r] : 0.f,
                      ((mask >> (2 * r + 1)) & 1u) ? y1[r] : 0.f);
  }
  lat_stamp(p, 7, true);
}
#endif

// ------------------------------------------------------------------------------------------------ host side
static inline int lat_align(int x, int a) { return (x + a - 1) / a * a; }

static int launch_latent(LatParams& p, void* stream) {
  constexpr bool bwd = MVAE_LAT_BWD != 0;
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;
  const mvae_pm_desc& D = p.desc;
  const int P = D.ld_ml, Sn = D.ld_eps, Sd = D.ld_z, C = D.C, R = kLatRows;
  int maxn = 0;
  bool dyn = false;
  for (int i = 0; i < C; ++i) {
    const int n = D.comp[i].n;
    dyn = dyn || !(n >= 1 && n <= 8 && n != 7);
    maxn = n > maxn ? n : maxn;
  }
  // shared-memory layout
  int o = lat_align(C * (int)sizeof(ItemInfo), 128);
  p.off_a = o;
  o += lat_align((bwd ? R * p.gdd_ld : R * p.h_ld) * 4, 128);
  p.off_b = o;
  if (bwd) o += lat_align(R * p.h_ld * 4, 128);
  p.off_ml = o;
  o += lat_align(R * P * 4, 16);
  p.off_part = o;  // forward: the heads' partial sums of the second half of the hidden dimension
  if (!bwd) o += lat_align(R * P * 4, 16);
  p.off_eps = o;
  o += lat_align(R * Sn * 4, 16);
  p.off_z = o;
  o += lat_align(R * Sd * 4, 16);
  p.off_kl = o;
  if (!bwd) o += lat_align(R * C * 4, 16);
  p.off_gz = o;
  if (bwd) o += lat_align(R * Sd * 4, 16);
  p.off_gml = o;
  if (bwd) o += lat_align(R * P * 4, 16);
  p.off_bar = o;
  o += 16;
  const size_t smem = (size_t)o;
  if (smem > (size_t)di.max_smem_optin) return MVAE_ERR_UNSUPPORTED;
  const int small = (bwd ? (Sd > P ? Sd : P) : P);
  void (*kern)(LatParams);
#if MVAE_LAT_BWD
#define MVAE_LAT_K(MN, SM) latent_backward_kernel<MN, SM>
#else
#define MVAE_LAT_K(MN, SM) latent_forward_kernel<MN, SM>
#endif
#define MVAE_LAT_PICK(MN) (small <= 16 ? MVAE_LAT_K(MN, 16) : small <= 32 ? MVAE_LAT_K(MN, 32) : MVAE_LAT_K(MN, 64))
  if (dyn) kern = MVAE_LAT_PICK(0);
  else if (maxn <= 2) kern = MVAE_LAT_PICK(2);
  else kern = MVAE_LAT_PICK(8);
#undef MVAE_LAT_PICK
#undef MVAE_LAT_K
  if (smem > 48 * 1024) MVAE_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t grid = (p.B + R - 1) / R;
  if (grid > 0x7fffffff) return MVAE_ERR_UNSUPPORTED;
  p.stamps = g_lat_stamps;
  p.debug_flags = g_lat_debug_flags;
  if (unsigned long long* region = debug_timeline_region(bwd ? 2 : 1, grid, (int)p.B, p.H, 0)) p.stamps = region;
  MVAE_CUDA_TRY(launch_pdl(kern, dim3((unsigned)grid), dim3(kLatThreads), smem, as_stream(stream), p));
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}

}  // namespace mvae
