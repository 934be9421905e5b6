// gemm_sm100.cu — the dense layers of the VAE as ONE warp-specialised tcgen05 GEMM kernel (sm_100a).
//
//   D[M,N] = sum_k A[m,k] * B[n,k]      A, B: split-bf16 operand planes (include/mvae_b200.h), fp32 accumulate
//
// Per CTA: one 128 x BLOCK_N output tile (BLOCK_N a multiple of 16, <= 256), optionally one split-K slice.
//   warp 0      TMA producer: cp.async.bulk.tensor (3-D maps: [plane][row][col], 128B swizzle) into an
//               mbarrier-guarded ring of shared-memory stages; one k-block = 64 bf16 of K per operand plane.
//   warp 1      allocates TMEM, then one elected lane issues tcgen05.mma (kind::f16, M=128, N=BLOCK_N, K=16) for
//               every plane pair (i,j) with i+j < max(planes); tcgen05.commit releases stages / signals the epilogue.
//   warps 2..9  epilogue: tcgen05.ld the fp32 accumulator (lane = row, column = n) and apply the fused epilogue
//               (bias, relu, relu-mask, BCE / Gaussian-NLL row sums + dloss/dlogits, split planes for the next GEMM).
//               A warp may only read the TMEM lanes 32 (warp % 4) .. +31, so two warps share each lane quadrant and
//               take alternating 16-column chunks: the epilogue of these short-K GEMMs is latency bound (MUFU chains
//               of the BCE, dependent tcgen05.ld waits), and eight warps hide twice as much of it as four.
// Operands may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows], read through an MN-major UMMA
// descriptor) so that dgrad and wgrad read the very same buffers as the forward pass — no transposes in HBM.
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "mvae_common.cuh"

namespace mvae {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;             // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;              // K of one tcgen05.mma kind::f16
constexpr int kEpiThreads = 256;        // 8 epilogue warps: two per TMEM lane quadrant, alternating 16-column chunks
constexpr int kGemmThreads = 64 + kEpiThreads;  // + TMA producer warp + MMA issuer warp
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KiB per plane per stage
constexpr int kMaxStages = 8;
constexpr float kHalfLn2PiG = 0.9189385332046727f;

struct GemmParams {
  int M, N, K;
  int block_n;       // multiple of 16
  int a_major, b_major;
  int a_planes, b_planes;
  int stages;
  int b_tile_bytes;  // per plane per stage
  int kb_total;      // ceil(K / 64)
  int kb_per_split;
  int epilogue;
  int atomic_out;    // split_k > 1
  uint32_t tmem_cols;
  uint32_t idesc;
  const float* bias;
  float* out_f32;
  int64_t ld_out;
  float* out_col;
  int col_split;
  uint16_t* op_base;  // out planes
  int64_t op_stride;
  int op_ld;
  int op_planes;
  const float* aux;
  int64_t ld_aux;
  int64_t aux_rows;  // > 0: targets repeat every aux_rows rows (row m reads x[m % aux_rows])
  const uint16_t* mask;
  int64_t ld_mask;
  float* rowsum;
  int tma_store;     // out planes leave through a shared-memory tile and ONE bulk tensor store per plane
  unsigned long long* stamps;  // diagnostics (mvae_debug_gemm): 16 %globaltimer words per CTA, null in production
  int debug_flags;             // timing experiments: 1 = no plane staging / stores, 2 = no loss math, 4 = no aux prefetch
};

// diagnostics: set by mvae_debug_gemm (api.cu), copied into GemmParams at launch
extern unsigned long long* g_gemm_stamps;
extern int g_gemm_debug_flags;
__device__ __forceinline__ void gemm_stamp(const GemmParams& p, int i) {
  if (p.stamps == nullptr) return;
  // slots 0-7: %globaltimer (comparable across CTAs, ~0.26 us resolution); 8-15: the SM's cycle counter (the first
  // epilogue thread's own progress; slot 15 = reference taken together with slot 5)
  unsigned long long t;
  if (i < 8) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  else t = (unsigned long long)clock64();
  const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  p.stamps[cta * 16 + i] = t;
}

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the same in two halves: the load is in flight while the caller issues independent work (the next chunk's global loads)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// (the registers are in/out operands of the wait so that no use of them can be scheduled ahead of it)
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// UMMA shared-memory descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor: start>>4 @0, LBO>>4 @16, SBO>>4 @32,
// version=1 @46, layout SWIZZLE_128B=2 @61).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// one-MUFU transcendentals (ex2 / lg2 / rcp .approx.ftz: what __expf / __logf / __fdividef reduce to, without the
// scaling multiplies the epilogue folds elsewhere)
__device__ __forceinline__ float exp2f_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// writes 16 consecutive values of row `m`, columns [n, n+16) as split-bf16 planes
// One cvt.rn.bf16x2 per column pair and plane; the residual for the next plane comes from the packed word itself (a
// bf16 is the upper half of an fp32: shift / mask, no second conversion).
__device__ __forceinline__ void store_planes16(const GemmParams& p, int m, int n, const float (&y)[16]) {
  float rem[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) rem[j] = y[j];
  for (int pl = 0; pl < p.op_planes; ++pl) {
    uint16_t* dst = p.op_base + (int64_t)pl * p.op_stride + (int64_t)m * p.op_ld + n;
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      w[k] = pack_bf16x2(rem[2 * k], rem[2 * k + 1]);
      rem[2 * k] -= __uint_as_float(w[k] << 16);
      rem[2 * k + 1] -= __uint_as_float(w[k] & 0xFFFF0000u);
    }
    if (n + 16 <= p.N) {
      *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(dst + 8) = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (n + j < p.N) dst[j] = (uint16_t)((j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xFFFFu));
    }
  }
}

// The same 16 values into the CTA's staging tile (shared memory, [plane][128 rows][block_n] bf16, row-major): the
// epilogue's plane outputs then leave with one cp.async.bulk.tensor store per plane — full rows, clipped by the
// tensor map at M and N — instead of 32-byte pieces of 32 different rows per warp instruction.
__device__ __forceinline__ void stage_planes16(const GemmParams& p, uint8_t* tile, int r, int c, const float (&y)[16]) {
  float rem[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) rem[j] = y[j];
  for (int pl = 0; pl < p.op_planes; ++pl) {
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      w[k] = pack_bf16x2(rem[2 * k], rem[2 * k + 1]);
      rem[2 * k] -= __uint_as_float(w[k] << 16);
      rem[2 * k + 1] -= __uint_as_float(w[k] & 0xFFFF0000u);
    }
    uint4* dst = reinterpret_cast<uint4*>(tile + ((size_t)(pl * kBlockM + r) * p.block_n + c) * 2);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__device__ __forceinline__ void store_f32_16(float* out, int64_t ld, int m, int n, int N, const float (&y)[16],
                                             bool vec) {
  float* dst = out + (int64_t)m * ld + n;
  if (vec && n + 16 <= N) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (n + j < N) dst[j] = y[j];
  }
}

__global__ void __launch_bounds__(kGemmThreads, 2)
    gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const __grid_constant__ CUtensorMap map_out, const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBlockM;
  const int n0 = blockIdx.y * p.block_n;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
  if (kb0 >= kb1) return;  // empty split-K slice (uniform over the CTA)
  const int nkb = kb1 - kb0;
  if (threadIdx.x == 0) gemm_stamp(p, 0);

  // ---- shared memory carve-up: [stages x (A planes | B planes)] | barriers | tmem slot ----
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t stage_bytes = p.a_planes * kATileBytes + p.b_planes * p.b_tile_bytes;
  const uint32_t bar_base = base + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kMaxStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* s_bias = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - raw));  // [block_n <= 256] bias slice

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (threadIdx.x == 0) gemm_stamp(p, 1);
  pdl_wait();  // everything above is independent of the previous kernel; operands and epilogue inputs are not
  if (threadIdx.x == 0) gemm_stamp(p, 2);

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // converged warp, one elected lane issues (see the MMA issuer below for why not `if (lane == 0) { loop }`)
    {
      const uint32_t tx_bytes = stage_bytes;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % p.stages;
        const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        if (elect_one_sync()) {
          mbar_expect_tx(full_bar(s), tx_bytes);
          const int k = (kb0 + i) * kBlockK;
          const uint32_t sa = base + s * stage_bytes;
          const uint32_t sb = sa + p.a_planes * kATileBytes;
          for (int pl = 0; pl < p.a_planes; ++pl) {
            const uint32_t dst = sa + pl * kATileBytes;
            if (p.a_major == MVAE_K_MAJOR) {
              tma_load_3d(dst, &map_a, full_bar(s), k, m0, pl);
            } else {
              tma_load_3d(dst, &map_a, full_bar(s), m0, k, pl);
              tma_load_3d(dst + 8192, &map_a, full_bar(s), m0 + 64, k, pl);
            }
          }
          for (int pl = 0; pl < p.b_planes; ++pl) {
            const uint32_t dst = sb + pl * p.b_tile_bytes;
            if (p.b_major == MVAE_K_MAJOR) {
              tma_load_3d(dst, &map_b, full_bar(s), k, n0, pl);
            } else {
              for (int c = 0; c * 8192 < p.b_tile_bytes; ++c)
                tma_load_3d(dst + c * 8192, &map_b, full_bar(s), n0 + c * 64, k, pl);
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    // The warp stays converged through the k-block loop (every lane polls the barrier and derives the same shared-memory
    // descriptors); the MMAs and commits themselves are issued by ONE lane chosen with elect.sync.  Written as
    // `if (lane == 0) { whole loop }` the compiler cannot prove that a single lane is active and wraps every
    // tcgen05.mma in an ELECT / BRA.U.ANY loop over the active lanes: ~25 dependent instructions per MMA on one warp,
    // i.e. more cycles to ISSUE an MMA than the tensor core needs to execute it (the issuing thread, not the data, was
    // what the epilogue warps were waiting for).
    {
      const int pmax = max(p.a_planes, p.b_planes);
      // K-major: rows of 128 B, 8-row groups 1024 B apart (SBO); k-step = 32 B inside the swizzle row.
      // MN-major: 64-element column chunks 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO); k-step = 16 rows = 2048 B.
      const uint32_t a_lbo = p.a_major == MVAE_K_MAJOR ? 16u : 8192u;
      const uint32_t b_lbo = p.b_major == MVAE_K_MAJOR ? 16u : 8192u;
      const uint32_t a_kstep = p.a_major == MVAE_K_MAJOR ? 32u : 2048u;
      const uint32_t b_kstep = p.b_major == MVAE_K_MAJOR ? 32u : 2048u;
      // descriptors differ only in their 14-bit start-address field: build the constant part once
      const uint64_t da_hi = make_desc(0u, a_lbo, 1024u), db_hi = make_desc(0u, b_lbo, 1024u);
      uint32_t accumulate = 0;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % p.stages;
        const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
        mbar_wait(full_bar(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = base + s * stage_bytes;
        const uint32_t sb = sa + p.a_planes * kATileBytes;
        const int k_left = p.K - (kb0 + i) * kBlockK;
        const int nks = k_left >= kBlockK ? kBlockK / kUmmaK : (k_left + kUmmaK - 1) / kUmmaK;
        if (elect_one_sync()) {
          if (i == 0) gemm_stamp(p, 3);  // first stage landed
          for (int pa = 0; pa < p.a_planes; ++pa)
            for (int pb = 0; pb < p.b_planes; ++pb) {
              if (pa + pb >= pmax) continue;
              const uint32_t ta = (sa + pa * kATileBytes) >> 4;
              const uint32_t tb = (sb + pb * p.b_tile_bytes) >> 4;
              for (int ks = 0; ks < nks; ++ks) {
                const uint64_t da = da_hi | (uint64_t)((ta + ks * (a_kstep >> 4)) & 0x3FFFu);
                const uint64_t db = db_hi | (uint64_t)((tb + ks * (b_kstep >> 4)) & 0x3FFFu);
                umma_bf16(tmem_base, da, db, p.idesc, accumulate);
                accumulate = 1;
              }
            }
          umma_commit(empty_bar(s));  // stage reusable once these MMAs retire
          if (i == nkb - 1) {
            umma_commit(tmem_full_bar);  // accumulator complete
            gemm_stamp(p, 4);            // last MMA issued
          }
        }
        accumulate = 1;
        __syncwarp();
      }
    }
  } else {
    // ===================================== epilogue =====================================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int c_first = ((warp - 2) >> 2) * 16;  // first column chunk of this warp (the quadrant's other warp: +16)
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < p.M;
    const int te = threadIdx.x - 64;  // 0..255 within the epilogue warps
    // stage the bias slice of this tile in shared memory while the main loop runs (split-K: first slice only)
    {
      const float* bias = blockIdx.z == 0 ? p.bias : nullptr;
      for (int i = te; i < p.block_n; i += kEpiThreads) s_bias[i] = (bias && n0 + i < p.N) ? __ldg(bias + n0 + i) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    const bool vec_out = p.out_f32 && ((p.ld_out & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out_f32) & 15) == 0);
    const bool vec_aux = p.aux && ((p.ld_aux & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0);
    const bool is_loss = p.epilogue == MVAE_EPI_BCE_ROWSUM || p.epilogue == MVAE_EPI_NLL_ROWSUM;
    const bool is_mask = p.epilogue == MVAE_EPI_RELU_MASK;
    const int64_t m_aux = p.aux_rows > 0 ? (int64_t)m % p.aux_rows : (int64_t)m;
    // operands of the epilogue that do not depend on the accumulator are fetched one chunk ahead
    float t_nxt[16];
    uint32_t mk_nxt = 0;
    auto prefetch = [&](int c) {
      const int n = n0 + c;
      if (!row_ok || n >= p.N || (p.debug_flags & 4)) return;
      if (is_loss) {
        const float* xr = p.aux + m_aux * p.ld_aux + n;
        if (vec_aux && n + 16 <= p.N) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 v = __ldg(reinterpret_cast<const float4*>(xr) + j);
            t_nxt[4 * j] = v.x;
            t_nxt[4 * j + 1] = v.y;
            t_nxt[4 * j + 2] = v.z;
            t_nxt[4 * j + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) t_nxt[j] = (n + j < p.N) ? __ldg(xr + j) : 0.f;
        }
      } else if (is_mask) {
        const uint16_t* mk = p.mask + (int64_t)m * p.ld_mask + n;
        uint32_t bits = 0;
        if (n + 16 <= p.ld_mask && ((p.ld_mask & 7) == 0)) {
          const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(mk));
          const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(mk) + 1);
          const uint32_t w[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            // bf16 > 0  <=>  sign bit clear and magnitude non-zero
            const uint32_t lo = w[j] & 0xFFFFu, hi = w[j] >> 16;
            bits |= (uint32_t)(((lo & 0x8000u) == 0) && ((lo & 0x7FFFu) != 0)) << (2 * j);
            bits |= (uint32_t)(((hi & 0x8000u) == 0) && ((hi & 0x7FFFu) != 0)) << (2 * j + 1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint16_t bq = (n + j < p.N) ? __ldg(mk + j) : (uint16_t)0;
            bits |= (uint32_t)(((bq & 0x8000u) == 0) && ((bq & 0x7FFFu) != 0)) << j;
          }
        }
        mk_nxt = bits;
      }
    };
    if (c_first < p.block_n) prefetch(c_first);
    mbar_wait(tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    pdl_launch_dependents();  // main loop done: the next kernel's CTAs may take the SM resources this CTA frees soon
    if (te == 0) gemm_stamp(p, 5);  // accumulator complete
    if (te == 0) gemm_stamp(p, 15);
    float row_acc = 0.f;
    // the operand stages are free once the accumulator is complete (every load consumed, every MMA retired): the first
    // of them doubles as the staging tile of the plane outputs
    uint8_t* tile = smem_raw + (base - raw);
    for (int c = c_first; c < p.block_n; c += 32) {
      const int n = n0 + c;
      if (n >= p.N) break;  // warp-uniform
      float t[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) t[j] = t_nxt[j];
      const uint32_t mk_cur = mk_nxt;
      uint32_t r[16];
      tmem_ld16_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);   // in flight under the prefetch below
      if (c + 32 < p.block_n) prefetch(c + 32);
      tmem_ld_wait(r);
      if (te == 0 && c == c_first) gemm_stamp(p, 8);   // first accumulator chunk in registers
      if (!row_ok) continue;
      float y[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = __uint_as_float(r[j]) + s_bias[c + j];
      const bool dbg_nostage = (p.debug_flags & 1) != 0, dbg_nomath = (p.debug_flags & 2) != 0;
      switch (p.epilogue) {
        case MVAE_EPI_STORE: {
          if (p.atomic_out && vec_out && n + 16 <= p.N && !(p.col_split >= n && p.col_split < n + 16)) {
            // split-K partial sums: four 128-bit reductions per row and chunk instead of sixteen scalar atomics
            float* dst = p.out_f32 + (int64_t)m * p.ld_out + n;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + j), "f"(y[j]), "f"(y[j + 1]),
                           "f"(y[j + 2]), "f"(y[j + 3])
                           : "memory");
          } else if (p.atomic_out) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n + j < p.N) {
                if (n + j == p.col_split) {
                  if (p.out_col) atomicAdd(p.out_col + m, y[j]);
                } else if (p.out_f32) {
                  atomicAdd(p.out_f32 + (int64_t)m * p.ld_out + n + j, y[j]);
                }
              }
          } else if (p.col_split >= n && p.col_split < n + 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n + j < p.N) {
                if (n + j == p.col_split) {
                  if (p.out_col) p.out_col[m] = y[j];
                } else if (p.out_f32) {
                  p.out_f32[(int64_t)m * p.ld_out + n + j] = y[j];
                }
              }
          } else if (p.out_f32) {
            store_f32_16(p.out_f32, p.ld_out, m, n, p.N, y, vec_out);
          }
        } break;
        case MVAE_EPI_BIAS_RELU: {
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] = fmaxf(y[j], 0.f);
          if (dbg_nostage && y[0] != 12345.f) break;
          if (p.tma_store) stage_planes16(p, tile, q * 32 + lane, c, y);
          else if (p.op_base) store_planes16(p, m, n, y);
          if (p.out_f32) store_f32_16(p.out_f32, p.ld_out, m, n, p.N, y, vec_out);
        } break;
        case MVAE_EPI_RELU_MASK: {
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] = ((mk_cur >> j) & 1u) ? y[j] : 0.f;
          if (dbg_nostage && y[0] != 12345.f) break;
          if (p.tma_store) stage_planes16(p, tile, q * 32 + lane, c, y);
          else if (p.op_base) store_planes16(p, m, n, y);
          if (p.out_f32) store_f32_16(p.out_f32, p.ld_out, m, n, p.N, y, vec_out);
        } break;
        default: {  // BCE_ROWSUM / NLL_ROWSUM
          if (p.out_f32) store_f32_16(p.out_f32, p.ld_out, m, n, p.N, y, vec_out);
          float g[16];
          // The epilogue of this GEMM is bound by ALU issue (scripts/gemm_phases.py: ~24 instructions per element at
          // two issue cycles each), so the arithmetic is written for instruction count:
          //   loss = (1-t) x - log_sigmoid(x) = max(x, 0) - t x + ln2 lg2(1 + e),  e = exp(-|x|)
          // is summed as two running sums (the lg2 terms take their ln2 once per chunk), the row-bound test of the
          // ragged last chunk is hoisted out of the element loop, and sigmoid - t is one select + one FMA behind the
          // reciprocal.  (log1p(e) as log(1 + e): the rounding of 1 + e costs <= 6e-8 ABSOLUTE per element, i.e. <= 5e-5
          // on a row sum of order 1e2 and nothing measurable on the ELBO.)
          const bool full = n + 16 <= p.N;
          if (dbg_nomath) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              g[j] = y[j] - t[j];
              row_acc += y[j];
            }
          } else if (p.epilogue == MVAE_EPI_BCE_ROWSUM) {
            float lin[16], l2[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float lg = y[j];
              const float ex = exp2f_approx(fabsf(lg) * -1.4426950408889634f);
              const float u = 1.f + ex;
              lin[j] = fmaf(-t[j], lg, fmaxf(lg, 0.f));
              l2[j] = lg2f_approx(u);
              g[j] = fmaf(rcpf_approx(u), lg >= 0.f ? 1.f : ex, -t[j]);
            }
            float s_lin = 0.f, s_lg2 = 0.f;
            if (full) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                s_lin += lin[j];
                s_lg2 += l2[j];
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (n + j < p.N) {
                  s_lin += lin[j];
                  s_lg2 += l2[j];
                }
            }
            row_acc += fmaf(s_lg2, 0.6931471805599453f, s_lin);
          } else {
            float s_sq = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float dlt = t[j] - y[j];
              g[j] = -dlt;
              if (full || n + j < p.N) s_sq = fmaf(dlt, dlt, s_sq);
            }
            const int cnt = full ? 16 : p.N - n;
            row_acc += fmaf(0.5f, s_sq, (float)cnt * kHalfLn2PiG);
          }
          if (te == 0 && c == c_first) gemm_stamp(p, 9);  // first chunk's loss math done
          if (dbg_nostage && g[0] != 12345.f) break;
          if (p.tma_store) stage_planes16(p, tile, q * 32 + lane, c, g);
          else if (p.op_base) store_planes16(p, m, n, g);
        } break;
      }
      if (te == 0 && c == c_first) gemm_stamp(p, 10);  // first chunk complete
    }
    if (is_loss && row_ok && p.rowsum) atomicAdd(p.rowsum + m, row_acc);
    if (te == 0) gemm_stamp(p, 6);  // this warp's chunks done
    if (p.tma_store) {
      // writes to shared memory -> visible to the async proxy, all 8 epilogue warps done, then one thread stores
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (te == 0) gemm_stamp(p, 11);  // all epilogue warps done
      if (te == 0) {
        for (int pl = 0; pl < p.op_planes; ++pl)
          tma_store_3d(&map_out, base + (uint32_t)(pl * kBlockM * p.block_n * 2), n0, m0, pl);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile may go away with the CTA
        gemm_stamp(p, 12);  // bulk stores have read the tile
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) gemm_stamp(p, 7);
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// 3-D map over split planes: dims {cols, rows, planes}; box {64, box_rows, 1}; 128-byte swizzle, zero OOB fill.
static int encode_planes(CUtensorMap* map, const mvae_planes& pl, int cols_bound, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return MVAE_ERR_CUDA;
  cuuint64_t dims[3] = {(cuuint64_t)cols_bound, (cuuint64_t)pl.rows, (cuuint64_t)pl.planes};
  cuuint64_t strides[2] = {(cuuint64_t)pl.ld * 2, (cuuint64_t)(pl.planes > 1 ? pl.plane_stride : (int64_t)pl.rows * pl.ld) * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, pl.base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MVAE_OK : MVAE_ERR_CUDA;
}

// 3-D map over output planes for the epilogue's bulk stores: box {block_n, 128, 1}, no swizzle (the staging tile is
// plain row-major); the hardware clips the box at the tensor's bounds (rows >= M, columns >= N are not written).
static int encode_out_planes(CUtensorMap* map, const mvae_planes& pl, int M, int N, int block_n) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return MVAE_ERR_CUDA;
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)pl.planes};
  cuuint64_t strides[2] = {(cuuint64_t)pl.ld * 2, (cuuint64_t)(pl.planes > 1 ? pl.plane_stride : (int64_t)pl.rows * pl.ld) * 2};
  cuuint32_t box[3] = {(cuuint32_t)block_n, (cuuint32_t)kBlockM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, pl.base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MVAE_OK : MVAE_ERR_CUDA;
}

static int check_planes(const mvae_planes& p) {
  if (!p.base || p.planes < 1 || p.planes > 3 || p.rows < 1 || p.cols < 1 || p.ld < p.cols)
    return MVAE_ERR_INVALID_ARGUMENT;
  if ((p.ld & 7) || (reinterpret_cast<uintptr_t>(p.base) & 15)) return MVAE_ERR_ALIGNMENT;
  if (p.planes > 1 && ((p.plane_stride & 7) || p.plane_stride < (int64_t)p.rows * p.ld)) return MVAE_ERR_ALIGNMENT;
  return MVAE_OK;
}

static int round_up(int x, int m) { return (x + m - 1) / m * m; }

// BLOCK_N: split N into t tiles of equal width (multiple of 16, <= cap) minimising padded columns.
static int pick_block_n(int N, int cap) {
  int best_bn = 0, best_waste = 1 << 30;
  const int t0 = (N + cap - 1) / cap;
  for (int t = t0; t <= t0 + 3; ++t) {
    int bn = round_up((N + t - 1) / t, 16);
    if (bn > cap) continue;
    int waste = ((N + bn - 1) / bn) * bn - N;
    if (waste < best_waste) {
      best_waste = waste;
      best_bn = bn;
    }
  }
  return best_bn;
}

}  // namespace mvae

using namespace mvae;

extern "C" int mvae_gemm(const mvae_gemm_args* a, void* stream) {
  if (!a) return MVAE_ERR_INVALID_ARGUMENT;
  if (a->M < 1 || a->N < 1 || a->K < 1 || a->split_k < 0) return MVAE_ERR_INVALID_ARGUMENT;
  if (a->epilogue < MVAE_EPI_STORE || a->epilogue > MVAE_EPI_NLL_ROWSUM) return MVAE_ERR_INVALID_ARGUMENT;
  if (a->split_k != 1 && a->epilogue != MVAE_EPI_STORE) return MVAE_ERR_INVALID_ARGUMENT;
  if ((a->a_major != MVAE_K_MAJOR && a->a_major != MVAE_MN_MAJOR) ||
      (a->b_major != MVAE_K_MAJOR && a->b_major != MVAE_MN_MAJOR))
    return MVAE_ERR_INVALID_ARGUMENT;
  int rc = check_planes(a->a);
  if (rc != MVAE_OK) return rc;
  rc = check_planes(a->b);
  if (rc != MVAE_OK) return rc;
  // logical shapes
  const int a_rows = a->a_major == MVAE_K_MAJOR ? a->M : a->K, a_cols = a->a_major == MVAE_K_MAJOR ? a->K : a->M;
  const int b_rows = a->b_major == MVAE_K_MAJOR ? a->N : a->K, b_cols = a->b_major == MVAE_K_MAJOR ? a->K : a->N;
  if (a->a.rows != a_rows || a->a.ld < a_cols || a->b.rows != b_rows || a->b.ld < b_cols)
    return MVAE_ERR_INVALID_ARGUMENT;
  if (a->epilogue == MVAE_EPI_STORE && !a->out_f32 && !a->out_col) return MVAE_ERR_INVALID_ARGUMENT;
  if ((a->epilogue == MVAE_EPI_BCE_ROWSUM || a->epilogue == MVAE_EPI_NLL_ROWSUM) && (!a->aux || a->ld_aux < a->N))
    return MVAE_ERR_INVALID_ARGUMENT;
  if (a->epilogue == MVAE_EPI_RELU_MASK && (!a->mask || a->ld_mask < a->N)) return MVAE_ERR_INVALID_ARGUMENT;
  if (a->out_f32 && a->ld_out < (a->col_split == a->N - 1 ? a->N - 1 : a->N)) return MVAE_ERR_INVALID_ARGUMENT;
  if (a->out_planes.base) {
    rc = check_planes(a->out_planes);
    if (rc != MVAE_OK) return rc;
    if (a->out_planes.rows < a->M || a->out_planes.ld < a->N) return MVAE_ERR_INVALID_ARGUMENT;
  }
  DeviceInfo di;
  rc = get_device_info(&di);
  if (rc != MVAE_OK) return rc;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = a->M;
  p.N = a->N;
  p.K = a->K;
  p.a_major = a->a_major;
  p.b_major = a->b_major;
  p.a_planes = a->a.planes;
  p.b_planes = a->b.planes;
  // Tile policy.  The shapes of this workload are small (M = batch of a few thousand rows), so one CTA per SM leaves
  // the tensor pipe idle during every prologue / epilogue.  Prefer TWO co-resident CTAs per SM (each <= ~112 KiB of
  // shared memory, >= 2 stages): one CTA's epilogue overlaps the other's main loop, and twice as many tiles fit in a
  // wave.  BLOCK_N is capped accordingly; a single-CTA-per-SM deep pipeline is used only when the grid is small.
  const int kExtra = 1024 /*align*/ + 256 /*barriers*/ + 1024 /*bias slice*/;
  const int smem_one = di.max_smem_optin - kExtra;
  const int smem_two = (di.max_smem_optin + 1024) / 2 - 1024 - kExtra;  // two CTAs + 1 KiB system reservation each
  auto stage_bytes_of = [&](int bn) {
    const int bt = p.b_major == MVAE_K_MAJOR ? bn * 128 : round_up(bn, 64) * 128;
    return p.a_planes * kATileBytes + p.b_planes * bt;
  };
  int cap = 256;
  while (cap > 32 && 2 * stage_bytes_of(cap) > smem_two) cap -= 16;
  int smem_budget = smem_two;
  p.block_n = pick_block_n(a->N, cap);
  if (p.block_n <= 0) return MVAE_ERR_UNSUPPORTED;
  {
    const int64_t tiles = (int64_t)((a->M + kBlockM - 1) / kBlockM) * ((a->N + p.block_n - 1) / p.block_n);
    if (tiles * (a->split_k > 1 ? a->split_k : 1) <= di.sm_count && a->split_k != 0) smem_budget = smem_one;
  }
  // caller's tile policy (mvae_gemm_args.tile_n / ctas_per_sm), honoured when it fits
  if (a->ctas_per_sm == 1) smem_budget = smem_one;
  if (a->ctas_per_sm == 2) smem_budget = smem_two;
  if (a->tile_n >= 16 && a->tile_n <= 256 && (a->tile_n & 15) == 0 && stage_bytes_of(a->tile_n) <= smem_budget)
    p.block_n = a->tile_n;
  else if (a->ctas_per_sm == 1 && stage_bytes_of(p.block_n) > smem_budget)
    smem_budget = smem_one;
  p.b_tile_bytes = p.b_major == MVAE_K_MAJOR ? p.block_n * 128 : round_up(p.block_n, 64) * 128;
  const int stage_bytes = p.a_planes * kATileBytes + p.b_planes * p.b_tile_bytes;
  p.kb_total = (a->K + kBlockK - 1) / kBlockK;
  int split = a->split_k;
  if (split == 0) {
    // auto: enough K slices to put one CTA on every SM, each slice keeping >= 4 k-blocks
    const int tiles = ((a->M + kBlockM - 1) / kBlockM) * ((a->N + p.block_n - 1) / p.block_n);
    split = di.sm_count / tiles;
    if (split > p.kb_total / 4) split = p.kb_total / 4;
    if (split < 1) split = 1;
  }
  if (split > p.kb_total) split = p.kb_total;
  p.kb_per_split = (p.kb_total + split - 1) / split;
  split = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.atomic_out = split > 1;
  int stages = smem_budget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages > p.kb_per_split) stages = p.kb_per_split;
  if (stages < 1) return MVAE_ERR_UNSUPPORTED;
  p.stages = stages;
  p.epilogue = a->epilogue;
  p.tmem_cols = p.block_n <= 32 ? 32 : p.block_n <= 64 ? 64 : p.block_n <= 128 ? 128 : 256;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_major << 15) | ((uint32_t)p.b_major << 16) |
            ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
  p.bias = a->bias;
  p.out_f32 = a->out_f32;
  p.ld_out = a->ld_out;
  p.out_col = a->out_col;
  p.col_split = a->col_split;
  p.op_base = a->out_planes.base;
  p.op_stride = a->out_planes.planes > 1 ? a->out_planes.plane_stride : 0;
  p.op_ld = a->out_planes.ld;
  p.op_planes = a->out_planes.base ? a->out_planes.planes : 0;
  p.aux = a->aux;
  p.ld_aux = a->ld_aux;
  p.aux_rows = a->aux_rows > 0 ? a->aux_rows : 0;
  p.mask = a->mask;
  p.ld_mask = a->ld_mask;
  p.rowsum = a->rowsum;
  p.stamps = g_gemm_stamps;
  p.debug_flags = g_gemm_debug_flags;

  CUtensorMap map_a, map_b, map_out;
  memset(&map_out, 0, sizeof(map_out));
  {
    // plane outputs through a staging tile + bulk tensor stores when the tile fits the (then idle) operand stages;
    // MVAE_GEMM_TMA_STORE=0 keeps the per-thread stores
    static const bool tma_store_on = [] {
      const char* e = getenv("MVAE_GEMM_TMA_STORE");
      return !(e && e[0] == '0');
    }();
    const size_t tile_bytes = (size_t)p.op_planes * kBlockM * p.block_n * 2;
    if (tma_store_on && p.op_planes > 0 && tile_bytes <= (size_t)stages * stage_bytes) {
      rc = encode_out_planes(&map_out, a->out_planes, a->M, a->N, p.block_n);
      if (rc != MVAE_OK) return rc;
      p.tma_store = 1;
    }
  }
  // K-major: inner dim = K (logical), box rows = tile rows.  MN-major: inner dim = M or N (logical), box = 64 x 64.
  rc = encode_planes(&map_a, a->a, a_cols, a->a_major == MVAE_K_MAJOR ? kBlockM : kBlockK);
  if (rc != MVAE_OK) return rc;
  rc = encode_planes(&map_b, a->b, b_cols, a->b_major == MVAE_K_MAJOR ? p.block_n : kBlockK);
  if (rc != MVAE_OK) return rc;

  const size_t smem = (size_t)stages * stage_bytes + kExtra;
  static std::once_flag attr_once[64];
  int dev = 0;
  MVAE_CUDA_TRY(cudaGetDevice(&dev));
  cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once[dev & 63], [&] {
    attr_err = cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    di.max_smem_optin);
  });
  MVAE_CUDA_TRY(attr_err);
  dim3 grid((a->M + kBlockM - 1) / kBlockM, (a->N + p.block_n - 1) / p.block_n, split);
  if (unsigned long long* region = debug_timeline_region(0, (long long)grid.x * grid.y * grid.z, a->M, a->N, a->K))
    p.stamps = region;
  MVAE_CUDA_TRY(launch_pdl(gemm_tcgen05_kernel, grid, dim3(kGemmThreads), smem, as_stream(stream), map_a, map_b, map_out,
                           p));
  MVAE_LAUNCH_CHECK();
  return MVAE_OK;
}
