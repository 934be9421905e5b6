/*
 * mvae_b200.h — C ABI of libmvae_b200.so: the B200-native (sm_100a) hot path of the mixed-curvature VAE.
 *
 * The reference (oskopek/mvae) is pure Python/PyTorch and has no FFI of its own; its extension points are
 * Python classes.  Each entry point below replaces a group of reference Python functions and
 * cites them (paths relative to the reference root).  The Python host layer (mvae_b200/*.py) mirrors the
 * reference classes and calls these through ctypes; INTEGRATION.md shows the stub a reference maintainer adds.
 *
 * Conventions
 *   - all matrices are row-major, contiguous, fp32 unless the name says `bf16` (uint16 storage);
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - all work is enqueued on the caller's `stream` (a cudaStream_t passed as void*); no entry point
 *     synchronises the host, allocates, or keeps global state (the only cached state is per-device
 *     attributes such as the SM count and func attributes set once);
 *   - return value: 0 on success, negative mvae_status on error (never throws / aborts);
 *     mvae_strerror() maps it to text;
 *   - numerical faults (non-finite outputs) are reported through an optional device flag word, never by a sync.
 */
#ifndef MVAE_B200_H_
#define MVAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVAE_ABI_VERSION 3
#define MVAE_MAX_COMPONENTS 96

/* ------------------------------------------------------------------------------------------------ status */
typedef enum mvae_status {
  MVAE_OK = 0,
  MVAE_ERR_INVALID_ARGUMENT = -1,  /* null pointer, bad shape, unknown component type            */
  MVAE_ERR_UNSUPPORTED = -2,       /* valid request this build cannot serve (e.g. dim too large) */
  MVAE_ERR_CUDA = -3,              /* a CUDA runtime / driver call failed (see mvae_last_cuda_error) */
  MVAE_ERR_ALIGNMENT = -4,         /* pointer / leading dimension violates the documented alignment */
  MVAE_ERR_NOT_SM100 = -5          /* device is not compute capability 10.x                       */
} mvae_status;

const char* mvae_strerror(int status);
int mvae_abi_version(void);
/* Last cudaError_t seen by this library on the calling thread (0 if none); does not clear sticky errors. */
int mvae_last_cuda_error(void);

/* ------------------------------------------------------------------------------------ product manifold */
/* Component kinds, reference grammar letters (mt/mvae/utils.py:30-38). */
typedef enum mvae_manifold {
  MVAE_EUCLIDEAN = 0,   /* 'e'  mt/mvae/ops/euclidean.py  + EuclideanNormalProcedure (sampling_procedures.py:145-155) */
  MVAE_HYPERBOLOID = 1, /* 'h'  mt/mvae/ops/hyperbolics.py + WrappedNormalProcedure (:91-116)   */
  MVAE_SPHERE = 2,      /* 's'  mt/mvae/ops/spherical.py   + WrappedNormalProcedure             */
  MVAE_POINCARE = 3,    /* 'p'  mt/mvae/ops/poincare.py (+geoopt 0.1.0) + WrappedNormalProcedure */
  MVAE_PROJ_SPHERE = 4, /* 'd'  mt/mvae/ops/spherical_projected.py + WrappedNormalProcedure     */
  MVAE_UNIVERSAL = 5    /* 'u'  mt/mvae/ops/universal.py + UniversalSamplingProcedure (sampling_procedures.py:184-206):
                                the component's entry of `radius` holds its raw CURVATURE parameter kappa; the kernels
                                branch per launch on its sign — kappa < -1e-6: Poincare ball, kappa > 1e-6: projected
                                sphere, both with R = 1/sqrt|kappa| (universal.py:30-31,64-74), else Euclidean — and
                                return d/dkappa in the component's slot of the radius gradient */
} mvae_manifold;

/* One latent component.  n = true (tangent) dimension; d = ambient dimension of loc/z
 * (n+1 for h,s — component.py:122,159 — and n for e,p,d). */
typedef struct mvae_component {
  int32_t type;    /* mvae_manifold                                                      */
  int32_t n;       /* true dimension                                                      */
  int32_t d;       /* ambient dimension                                                   */
  int32_t m_off;   /* column of fc_mean's n outputs inside a row of `ml`                  */
  int32_t l_off;   /* column of fc_logvar's l_n outputs inside a row of `ml`              */
  int32_t l_n;     /* n (elliptic) or 1 (scalar_parametrization, component.py:47-50)       */
  int32_t eps_off; /* column of this component's n noise values inside a row of `eps` / `sigma` */
  int32_t z_off;   /* column of this component's d coordinates inside a row of `z` / `mu` */
} mvae_component;

typedef struct mvae_pm_desc {
  int32_t C;       /* number of components (1..MVAE_MAX_COMPONENTS)                      */
  int32_t ld_ml;   /* row stride (floats) of ml   — sum of n + sum of l_n when packed     */
  int32_t ld_eps;  /* row stride (floats) of eps, sigma — sum of n when packed            */
  int32_t ld_z;    /* row stride (floats) of z, mu, gz  — sum of d when packed            */
  mvae_component comp[MVAE_MAX_COMPONENTS];
} mvae_pm_desc;

/* Fill offsets of a packed descriptor: ml = [m_0|l_0|m_1|l_1|...], eps/sigma and z/mu concatenated in
 * component order (concat order of mt/mvae/models/vae.py:78).  types[i] in mvae_manifold, dims[i] = true dim. */
int mvae_pm_desc_init(mvae_pm_desc* desc, int32_t C, const int32_t* types, const int32_t* dims,
                      int32_t scalar_parametrization);

/*
 * Fused per-sample manifold + Wrapped-Normal forward for one product manifold, B samples.
 * Replaces, per component: Component.encode (component.py:63-75: exp_map_mu0 + softplus+1e-5),
 * SamplingProcedure.reparametrize (sampling_procedures.py:93-99,147-151), q_z.rsample_with_parts
 * (wrapped_normal.py:70-78 -> sample_projection_mu0: hyperbolics.py:138-142, spherical.py:119-123,
 * poincare.py:152-157), kl_loss (sampling_procedures.py:101-116,153-155: log_prob_from_parts
 * wrapped_normal.py:84-97, logdet hyperbolics.py:58-65 / spherical.py:58-67 / poincare.py:55-89,
 * prior log_prob wrapped_normal.py:99-103) and the concat of vae.py:78.
 *
 *   ml     [B, ld_ml]   head pre-activations (fc_mean | fc_logvar outputs)
 *   eps    [B, ld_eps]  standard-normal noise (one draw per tangent coordinate)
 *   radius [C]          raw radius parameters (_nradius/_pradius); R = clamp(relu(.),1e-8,1e8) (manifold.py:73-75);
 *                       ignored for Euclidean components
 *   z      [B, ld_z]    out: concat_z
 *   kl     [B, C]       out: per-sample, per-component KL term
 *   mu     [B, ld_z]    out, optional (NULL): q_z.loc
 *   sigma  [B, ld_eps]  out, optional (NULL): q_z.scale (expanded to n columns)
 *   nonfinite_flag      optional (NULL) device word, OR-ed with 1 if any output is not finite
 */
int mvae_pm_forward(const mvae_pm_desc* desc, int64_t B, const float* ml, const float* eps, const float* radius,
                    float* z, float* kl, float* mu, float* sigma, uint32_t* nonfinite_flag, void* stream);

/*
 * Backward of mvae_pm_forward by recomputation from its inputs (autograd replay of the functions above,
 * including the custom backward rules of ops/common.py:28-39 LeakyClamp, :46-63 Atanh, :76-94 Acosh).
 *   gz      [B, ld_z]  upstream gradient of z
 *   gkl     [B, C]     upstream gradient of kl, or NULL meaning the constant `gkl_scalar` (beta) for every entry
 *   gml     [B, ld_ml] out: gradient of ml
 *   gradius [C]        out: ACCUMULATED (+=) gradient of the raw radius parameters (zero it first)
 */
int mvae_pm_backward(const mvae_pm_desc* desc, int64_t B, const float* ml, const float* eps, const float* radius,
                     const float* gz, const float* gkl, float gkl_scalar, float* gml, float* gradius, void* stream);

/* --------------------------------------------------------------------------- standalone manifold ops */
/* Per-sample vector ops of the Manifold interface (ops/manifold.py:22-60) for a single component.
 * x/y are [B, d] or [B, n] as the op requires; out likewise.  `radius` is a device pointer to the raw radius. */
typedef enum mvae_op {
  MVAE_OP_EXP_MAP_MU0 = 0,         /* x [B,n] -> out [B,d]       (hyperbolics.py:114, spherical.py:94, poincare.py:132, euclidean.py:78) */
  MVAE_OP_INV_EXP_MAP_MU0 = 1,     /* x [B,d] -> out [B,d]       (hyperbolics.py:131, spherical.py:112, poincare.py:148, euclidean.py:86) */
  MVAE_OP_EXP_MAP = 2,             /* x tangent [B,d], y at_point [B,d] -> out [B,d] (hyperbolics.py:106, spherical.py:86, poincare.py:124, euclidean.py:74) */
  MVAE_OP_INV_EXP_MAP = 3,         /* x point, y at_point -> tangent (hyperbolics.py:124, spherical.py:104, poincare.py:140, euclidean.py:82) */
  MVAE_OP_PT_MU0 = 4,              /* x tangent at mu0, y dst -> tangent at dst (hyperbolics.py:87, spherical.py:74, poincare.py:116) */
  MVAE_OP_INV_PT_MU0 = 5,          /* x tangent at src, y src -> tangent at mu0 (hyperbolics.py:96, spherical.py:80, poincare.py:120) */
  MVAE_OP_DISTANCE = 6,            /* x, y points -> out [B,1] geodesic distance (poincare.py:96-105; tests/mvae/ops/test_hyperbolics.py:46, test_spherical.py:45, test_euclidean.py:41) */
  MVAE_OP_MOBIUS_ADD = 7,          /* Poincare / projected sphere only: x (+) y (poincare.py:100, spherical_projected.py:107-113) */
  MVAE_OP_MOBIUS_SCALAR_MUL = 8,   /* Poincare only: x[B,d], y[B,1] scalar r -> r (x) x. No reference call site: parity unpinned. */
  MVAE_OP_LOGDET = 9,              /* x = u (h,s: [B,d]) -> out [B,1] (hyperbolics.py:58, spherical.py:58); d: x = z, y = mu (spherical_projected.py:58-92) */
  MVAE_OP_TO_POINCARE = 10,        /* h: lorentz_to_poincare (hyperbolics.py:151); s: spherical_to_projected (spherical.py:132) */
  MVAE_OP_FROM_POINCARE = 11       /* p: poincare_to_lorentz (poincare.py:167); d: projected_to_spherical (spherical_projected.py:191) */
} mvae_op;

int mvae_manifold_op(int32_t op, int32_t manifold, int32_t n, int64_t B, const float* x, const float* y,
                     const float* radius, float* out, void* stream);

/* Wrapped-Normal pieces for one component (distributions/wrapped_normal.py:70-103).
 * rsample_with_parts: loc [B,d], scale [B,n], eps [B,n] -> z [B,d], u [B,d] (data[0]), v [B,n] (data[1]).
 * log_prob_from_parts: scale [B,n], u [B,d], v [B,n] (+loc,z for Poincare) -> logp [B].
 * log_prob: loc [B,d], scale [B,n], z [B,d] -> logp [B] (inverse_sample_projection_mu0 then from_parts). */
int mvae_wn_rsample(int32_t manifold, int32_t n, int64_t B, const float* loc, const float* scale, const float* eps,
                    const float* radius, float* z, float* u, float* v, void* stream);
int mvae_wn_log_prob_from_parts(int32_t manifold, int32_t n, int64_t B, const float* loc, const float* scale,
                                const float* z, const float* u, const float* v, const float* radius, float* logp,
                                void* stream);
int mvae_wn_log_prob(int32_t manifold, int32_t n, int64_t B, const float* loc, const float* scale, const float* z,
                     const float* radius, float* logp, void* stream);

/* ----------------------------------------------------------------------------------- dense layers (MLP) */
/* The dense layers of the reference (mt/mvae/models/ffnn_vae.py:42-60 fc_e0 / fc_d0 / fc_logits and the
 * fc_mean / fc_logvar heads of mt/mvae/components/component.py:64,69), forward, dgrad and wgrad, all run through
 * ONE tcgen05 GEMM kernel.  fp32 accuracy on bf16 tensor cores comes from split-bf16 operand planes:
 * an fp32 matrix X [R, C] is carried as `planes` bf16 matrices X_0 + X_1 (+ X_2) ~= X
 * (X_0 = bf16(X), X_1 = bf16(X - X_0), ...), each [R, ld] row-major with ld a multiple of 8 elements (16 B)
 * so that TMA can address it; plane p starts at base + p*plane_stride.  The kernel accumulates the products
 * X_i * Y_j with i + j < max(planes) in fp32 (2 planes: 3 MMAs per k-step, ~2^-16 relative error). */
typedef struct mvae_planes {
  uint16_t* base;        /* device pointer to plane 0                               */
  int64_t plane_stride;  /* elements between consecutive planes (multiple of 8)     */
  int32_t rows;          /* R                                                       */
  int32_t cols;          /* C (logical)                                             */
  int32_t ld;            /* row stride in elements, multiple of 8                   */
  int32_t planes;        /* 1, 2 or 3                                               */
} mvae_planes;

/* fp32 [R,K] (row stride ld_src) -> planes, and optionally the transposed planes [K,R]. */
int mvae_split_planes(const float* src, int64_t ld_src, int32_t R, int32_t K, const mvae_planes* dst,
                      const mvae_planes* dst_transposed /* may be NULL */, void* stream);

/* Epilogues of the tcgen05 GEMM  D[M,N] = sum_k A[m,k] * B[n,k]. */
typedef enum mvae_epilogue {
  MVAE_EPI_STORE = 0,       /* v = acc (+bias[n]); out_f32[m*ld_out+n] = v, or += v (atomic) when split_k > 1;
                               column `col_split` (if >= 0) is diverted to out_col[m] (bias gradients)          */
  MVAE_EPI_BIAS_RELU = 1,   /* y = relu(acc + bias[n]) -> out_planes (and out_f32 if given) — ffnn_vae.py:48,56   */
  MVAE_EPI_RELU_MASK = 2,   /* y = acc * (mask[m,n] > 0) -> out_planes (and out_f32) — backward of the relu;
                               mask = plane 0 of the forward activation                                         */
  MVAE_EPI_BCE_ROWSUM = 3,  /* logit = acc + bias[n]; rowsum[m] += sum_n BCEWithLogits(logit, x[m,n]);
                               g = sigmoid(logit) - x -> out_planes; logits -> out_f32 if given
                               (image_reconstruction.py:81-82 + vae.py:131)                                     */
  MVAE_EPI_NLL_ROWSUM = 4   /* logit = acc + bias[n]; rowsum[m] += sum_n .5(x-logit)^2 + .5 ln 2pi; g = logit - x
                               (synthetic.py:161-162)                                                           */
} mvae_epilogue;

/* Operand layouts.  K_MAJOR: `planes` describes the matrix [rows = M or N, cols = K] (nn.Linear weight for the
 * forward pass, activations as A).  MN_MAJOR: `planes` describes [rows = K, cols = M or N] — the same row-major
 * buffer read "transposed" by TMA + an MN-major UMMA descriptor, which is how dgrad reads W and wgrad reads the
 * activations without any transposed copy in HBM. */
typedef enum mvae_operand_major { MVAE_K_MAJOR = 0, MVAE_MN_MAJOR = 1 } mvae_operand_major;

typedef struct mvae_gemm_args {
  mvae_planes a;           /* K_MAJOR: [M, K]   MN_MAJOR: [K, M]                       */
  mvae_planes b;           /* K_MAJOR: [N, K]   MN_MAJOR: [K, N]                       */
  int32_t a_major;         /* mvae_operand_major                                       */
  int32_t b_major;
  int32_t M, N, K;
  int32_t epilogue;        /* mvae_epilogue                                            */
  int32_t split_k;         /* 1 = none; k > 1 = k slices of K, 0 = as many as fill the SMs — both only with
                              MVAE_EPI_STORE and atomically ACCUMULATING into out_f32/out_col (zero them first) */
  const float* bias;       /* [N] or NULL                                              */
  float* out_f32;          /* [M, ld_out] or NULL                                      */
  int64_t ld_out;
  float* out_col;          /* STORE: [M] receives column col_split, or NULL            */
  int32_t col_split;       /* -1 if unused                                             */
  mvae_planes out_planes;  /* base == NULL if unused; [M, N] planes                    */
  const float* aux;        /* BCE/NLL: targets x [M, ld_aux]                           */
  int64_t ld_aux;
  const uint16_t* mask;    /* RELU_MASK: bf16 [M, ld_mask] (plane 0 of the activation) */
  int64_t ld_mask;
  float* rowsum;           /* BCE/NLL: [M] accumulated (+=) row sums                   */
  /* Tile policy overrides (0 = automatic), for callers that time a few candidates once per shape and keep the best
   * (the shapes of this workload are launch- and L2-bound, the best tile depends on M, N, K and the operand layout):
   * tile_n = BLOCK_N (multiple of 16, 16..256); ctas_per_sm = 1 (deep pipeline) or 2 (co-resident CTAs overlap
   * epilogue and main loop).  Requests that do not fit shared memory fall back to the automatic choice. */
  int32_t tile_n;
  int32_t ctas_per_sm;
  /* BCE/NLL: when > 0 the targets have only aux_rows rows and row m reads x[m % aux_rows] — the importance-sampling
   * path decodes n samples per input row ([n*B, .] activations against [B, D] targets; vae.py:108-109 materialises
   * x.repeat((n, 1, 1)) instead). */
  int64_t aux_rows;
} mvae_gemm_args;

int mvae_gemm(const mvae_gemm_args* args, void* stream);

/* ---------------------------------------------------------------------------------- skinny dense layers */
/* Dense layers whose N or K is tiny — the latent heads (component.py:64,69: N = sum(n)+sum(l_n)) and the first decoder
 * layer (ffnn_vae.py:56: K = sum(d)) and their backward passes — run on the CUDA cores in exact fp32 FMA arithmetic
 * (they are far below a tensor-core tile).  W(n, k) = W[n*w_stride_n + k*w_stride_k], so a weight can be read as
 * stored ([out, in]) or transposed without a copy.  Wide operands are given either as fp32 or as split planes. */

/* out[b, n] = sum_k A[b, k] W(n, k) (+ bias[n]);  N <= 64, any K.   A: a_f32 [B, ld_a] or a_planes. */
int mvae_skinny_rowdot(int64_t B, int32_t K, int32_t N, const float* a_f32, int64_t ld_a, const mvae_planes* a_planes,
                       const float* W, int64_t w_stride_n, int64_t w_stride_k, const float* bias, float* out,
                       int64_t ld_out, void* stream);

/* out[b, n] = act(sum_k A[b, k] W(n, k) + bias[n]);  K <= 64, any N.  act: 0 none, 1 relu, 2 relu-mask (mask = bf16
 * [B, ld_mask], plane 0 of the forward activation).  Result as planes and / or fp32. */
int mvae_skinny_expand(int64_t B, int32_t K, int32_t N, const float* a, int64_t ld_a, const float* W,
                       int64_t w_stride_n, int64_t w_stride_k, const float* bias, int32_t act, const uint16_t* mask,
                       int64_t ld_mask, const mvae_planes* out_planes, float* out_f32, int64_t ld_out, void* stream);

/* out[s*out_stride_s + w*out_stride_w] += sum_b small[b, s] wide[b, w]   (ACCUMULATES: zero the outputs first);
 * S <= 64.  small_ones != 0: out_row[w] += sum_b wide[b, w] (bias gradient of the wide side);
 * col_split >= 0: wide column col_split (a ones column) goes to out_col[s] += sum_b small[b, s]. */
int mvae_skinny_wgrad(int64_t B, int32_t S, int32_t Wd, const float* small, int64_t ld_small, int32_t small_ones,
                      const float* wide_f32, int64_t ld_wide, const mvae_planes* wide_planes, float* out,
                      int64_t out_stride_s, int64_t out_stride_w, float* out_row, float* out_col, int32_t col_split,
                      void* stream);

/* -------------------------------------------------------------------------------- fused latent block */
/* The whole latent block of FeedForwardVAE in one launch per direction.  The per-sample intermediates (head
 * pre-activations on the way back, gz, gml) never leave shared memory and h / gdd are read exactly once.
 *
 * forward (ModelVAE.forward vae.py:69-80 between fc_e0 and fc_logits):
 *   ml = h Wh^T + bh            all fc_mean / fc_logvar heads (component.py:64,69), Wh [P, H], P = desc->ld_ml
 *   z, kl = mvae_pm_forward(ml, eps, radius)
 *   dd = relu(z Wd0^T + bd0)    fc_d0 (ffnn_vae.py:56), Wd0 [H, Sd], Sd = desc->ld_z, written as split planes
 * h: relu(fc_e0(x)) as fp32 [B, ld_h] (the producing GEMM's epilogue writes it; 16-byte aligned, ld_h % 4 == 0).
 * Requirements: H % 8 == 0, P <= 64, Sd <= 64, Wh 16-byte aligned.  Outputs ml [B, P], z [B, Sd], kl [B, C] are kept for the backward pass / the statistics. */
int mvae_latent_forward(const mvae_pm_desc* desc, int64_t B, int32_t H, const float* h, int64_t ld_h, const float* Wh,
                        const float* bh, const float* eps, const float* radius, const float* Wd0, const float* bd0,
                        float* ml, float* z, float* kl, const mvae_planes* dd_out, uint32_t* nonfinite_flag,
                        void* stream);

/* The same launch taking the head of the train step with it (what mvae_step_prologue does in a launch of its own):
 *   draw_eps != 0: eps [B, desc->ld_eps] is DRAWN here — Normal.rsample's noise (wrapped_normal.py:72), Philox4x32-10
 *                  with seed / offset *counter_dev, element for element the stream of mvae_step_prologue — and written
 *                  to `eps` for the backward pass, instead of being read;
 *   zero_ptr / zero_n (n_zero <= 4 spans of floats): zero-filled by this launch — optimizer.zero_grad() (vae.py:151)
 *                  and the reconstruction row sums; nothing before the latent block of a step accumulates into them.
 * As a separate launch the prologue was a parallel branch of the step's graph, and this kernel the node that joined it:
 * a full launch latency behind fc_e0 instead of a programmatic edge (scripts/step_timeline.py: 3.8 us). */
typedef struct mvae_latent_prologue {
  int32_t draw_eps;
  int32_t n_zero;
  uint64_t seed;
  const uint64_t* counter_dev;
  float* zero_ptr[4];
  int64_t zero_n[4];
} mvae_latent_prologue;
int mvae_latent_forward_ex(const mvae_pm_desc* desc, int64_t B, int32_t H, const float* h, int64_t ld_h, const float* Wh,
                           const float* bh, float* eps, const float* radius, const float* Wd0, const float* bd0,
                           float* ml, float* z, float* kl, const mvae_planes* dd_out, uint32_t* nonfinite_flag,
                           const mvae_latent_prologue* prologue, void* stream);

/* backward of the block for d(loss)/d(dd) = gdd (fp32 [B, ld_gdd], already masked by relu'(dd)) and
 * d(loss)/d(kl) = gkl_scalar:
 *   gz = gdd Wd0;  gml = mvae_pm_backward(ml, eps, radius, gz, gkl_scalar);  gh = (gml Wh) * 1[h > 0]  -> planes
 *   gWd0 [H, Sd] += gdd^T z;  gbd0 [H] += colsum(gdd);  gWh [P, H] += gml^T h;  gbh [P] += colsum(gml);
 *   gradius [C] += dR          (all ACCUMULATED: zero them first)
 * Requirements as above plus Wd0 / gWd0 16-byte and Wh / gWh / gbd0 8-byte aligned. */
int mvae_latent_backward(const mvae_pm_desc* desc, int64_t B, int32_t H, const float* gdd, int64_t ld_gdd,
                         const float* h, int64_t ld_h, const float* Wh, const float* Wd0, const float* ml, const float* eps, const float* radius,
                         const float* z, float gkl_scalar, const mvae_planes* gh_out, float* gWd0, float* gbd0,
                         float* gWh, float* gbh, float* gradius, void* stream);

/* ------------------------------------------------------------------------------- reconstruction + ELBO */
/* Standalone reconstruction losses (VaeDataset.reconstruction_loss + .sum(-1), vae.py:131):
 * kind 0 = BCE-with-logits (data/image_reconstruction.py:81-82), 1 = unit-variance Gaussian NLL (data/synthetic.py:161-162).
 * rowsum [B] = sum over D; glogits (optional) = d(sum)/dlogits. */
int mvae_recon_loss(int32_t kind, int64_t B, int32_t D, const float* logits, const float* x, float* rowsum,
                    float* glogits, void* stream);

/* ELBO reduction (stats.py:144-202): out = [bce_sum, kl_sum, elbo, kl_c sums (C)] with
 * elbo = sum_b(-bce_b - beta * sum_c kl_bc).  Warp-shuffle + block reduce; `out` [3+C] is overwritten. */
int mvae_elbo_reduce(int64_t B, int32_t C, const float* bce, const float* kl, float beta, float* out, void* stream);

/* ------------------------------------------------------------------- convolutional encoder / decoder */
/* ConvolutionalVAE (mt/mvae/models/conv_vae.py:28-79): nn.Conv2d / nn.ConvTranspose2d with kernel 4, stride 2,
 * padding 1 (:47-55) run through mvae_gemm with the filter taps unrolled into the contraction dimension.
 * Activations are channels-last: [B, H, W, C] = the row-major matrix [B*H*W, C] as split-bf16 planes.
 *   Conv2d           y[B*OH*OW, Co] = im2col(x)[B*OH*OW, 16 Ci] . W[Co, (ky,kx,ci)]^T      (+ bias, relu in the GEMM)
 *   ConvTranspose2d  y = col2im( x[B*H*W, Ci] . W[Ci, (ky,kx,co)] ) + bias, relu
 * and each one's input gradient is the other's data movement around the same GEMM.
 *
 * mvae_conv_im2col: dst[(b,oy,ox), (ky,kx,c)] = src[(b, 2oy-1+ky, 2ox-1+kx), c] (0 outside the image), plane by plane
 *   (dst->planes <= src->planes); ones_col != 0 additionally writes 1.0 into column 16 C (dst->ld must leave room):
 *   read as a weight-gradient operand it yields the bias gradient.  src [B*H*W, C], dst [B*(H/2)*(W/2), 16 C].
 * mvae_conv_col2im: out[(b,oy,ox), c] = bias[c] + sum over the <= 4 (ky,kx) with oy = 2iy-1+ky, ox = 2ix-1+kx of
 *   cols[(b,iy,ix), (ky,kx,c)]; act 0 none / 1 relu / 2 zero where mask (plane 0, [B*2H*2W, C]) <= 0; written as
 *   planes (out_planes) and / or fp32 (out_f32, row stride ld_out).  cols fp32 [B*H*W, ld_cols >= 16 C].
 * mvae_permute_sc: [B, C*S] rows in (c, s) order (the reference's x.view(bs, -1) of a [B, C, H, W] tensor, S = H*W,
 *   conv_vae.py:61,65,73,77) <-> [B*S, C] channels-last rows; elements of 2 (planes: `planes` of them, strides in
 *   elements) or 4 bytes; to_nhwc selects the direction; src_ld / dst_ld are row strides in elements.
 * mvae_colsum: out[c] += sum_m src[m, c] (src fp32 with row stride ld, or the sum of src_planes' planes): bias gradients
 *   of the transposed convolutions.  C <= 1024; out is accumulated into (the caller zeroes it). */
int mvae_conv_im2col(const mvae_planes* src, int32_t B, int32_t H, int32_t W, int32_t C, const mvae_planes* dst,
                     int32_t ones_col, void* stream);
int mvae_conv_col2im(const float* cols, int64_t ld_cols, int32_t B, int32_t H, int32_t W, int32_t C, const float* bias,
                     int32_t act, const mvae_planes* mask, const mvae_planes* out_planes, float* out_f32,
                     int64_t ld_out, void* stream);
int mvae_permute_sc(int32_t elem_bytes, const void* src, int64_t src_ld, int64_t src_plane_stride, void* dst,
                    int64_t dst_ld, int64_t dst_plane_stride, int32_t planes, int64_t B, int32_t S, int32_t C,
                    int32_t to_nhwc, void* stream);
int mvae_colsum(const float* src_f32, const mvae_planes* src_planes, int64_t M, int32_t C, int64_t ld, float* out,
                void* stream);

/* ------------------------------------------------------------------------------------ input pipeline */
/* Dynamic binarisation of image batches on the device.  The reference does it per sample on the CPU inside its
 * DataLoader (ImageDynamicBinarization, mt/data/image_reconstruction.py:37-53) and ships float batches; here the batch
 * travels as stored (uint8 grayscale src [B, ld_src]) and one kernel produces
 *   x [B, ld_x] fp32 (reconstruction targets) and / or x_planes (plane 0 of the fc_e0 operand; 0 / 1 are exact in bf16)
 *   x[b,k] = (v > u[b,k]) ? 1 : 0,  v = src/255 as float32 (ToTensor), v <- 1 - v if invert
 * mode 0: dynamic — u [B, D] supplied (bit-exact against the reference's comparison for the same draws), or NULL =
 *         drawn in the kernel (Philox4x32-10: key `seed`, offset 8 * *offset_dev; the device word is read, not advanced,
 *         so the call can be captured in a CUDA graph — the caller increments it once per step);
 * mode 1: fixed threshold 0.5 (the reference's evaluation transform). */
int mvae_binarize(const uint8_t* src, int64_t ld_src, int64_t B, int32_t D, int32_t mode, int32_t invert, const float* u,
                  uint64_t seed, const uint64_t* offset_dev, float* x, int64_t ld_x, const mvae_planes* x_planes,
                  void* stream);

/* Head of a train step, ONE launch: (a) eps [n_eps] ~ N(0, 1) — the draw of Normal.rsample behind every Wrapped-Normal
 * / Euclidean-Normal sample (wrapped_normal.py:72, wrapped_distributions.py:25-27; torch's generator there, Philox4x32-10
 * + Box-Muller here: key `seed`, offset 4 * *counter_dev, which is read, not advanced — graph replayable); eps NULL or
 * n_eps 0 = the caller supplies the noise; (b) zero fill of up to 6 float buffers (the step's accumulating outputs:
 * gradient bucket, reconstruction row sums, split-K targets), replacing optimizer.zero_grad() (vae.py:151). */
int mvae_step_prologue(float* eps, int64_t n_eps, uint64_t seed, const uint64_t* counter_dev, int32_t n_zero,
                       float* const* zero_ptr, const int64_t* zero_n, void* stream);
/* *counter_dev += inc (the per-model step counter behind the Philox offsets above; one launch per step). */
int mvae_counter_add(uint64_t* counter_dev, uint64_t inc, void* stream);

/* ring[(*counter_dev) % capacity][0..n) = src[0..n), then ++*counter_dev: the per-step ELBO statistics
 * (BatchStatsFloat, stats.py:115-127 — the reference reads them with 3 + C .item() calls per step) are parked in a
 * device-side ring by the step itself and fetched by the epoch loop with ONE device->host copy per `capacity` steps
 * instead of one per step. */
int mvae_ring_push(const float* src, int32_t n, float* ring, int32_t capacity, uint64_t* counter_dev, void* stream);

/* ------------------------------------------------------ importance-weighted log-likelihood (evaluation) */
/* ModelVAE.log_likelihood (vae.py:82-123) draws n samples per input row.  The encoder runs once; per chunk of `ns`
 * samples mvae_iwae_latent does Component.encode's manifold part + rsample_log_probs of EVERY component
 * (sampling_procedures.py:47-50,106-110; wrapped_normal.py:70-103; EuclideanNormal.log_prob
 * wrapped_distributions.py:39-42) from the ONE set of head pre-activations ml [B, ld_ml]:
 *   eps  [ns, B, ld_eps]  standard normal draws
 *   z    [ns, B, ld_z]    samples, concat order of vae.py:105
 *   diff [ns, B]          sum_c (log q_c(z|x) - log p_c(z))   (single-sample Monte-Carlo terms for every component,
 *                         Euclidean ones included — unlike the analytic KL of the training path)
 *   zsum [B, ld_z]        += sum_s z   (atomic accumulate; feeds cov_norm, vae.py:119-121) — may be NULL
 * ld_ml <= 64 and ld_z <= 64. */
int mvae_iwae_latent(const mvae_pm_desc* desc, int64_t B, int32_t ns, const float* ml, const float* eps,
                     const float* radius, float* z, float* diff, float* zsum, void* stream);

/* log_p_x[b] = logsumexp_s(-recon[s,b] - diff[s,b]) - log n;  mi[b] = logsumexp_s(diff[s,b]) - log n
 * (vae.py:111-117); recon, diff [n, B]. */
int mvae_iwae_reduce(int32_t n, int64_t B, const float* recon, const float* diff, float* log_p_x, float* mi,
                     void* stream);

/* cov_norm (vae.py:119-121): with zbar[b] = zsum[b] / n,
 *   || sum_b (x[b] - mean_b x)^T (zbar[b] - mean_b zbar) ||_F  ->  out[0].
 * work: [B * Sd + D * Sd] floats of scratch.  Three small launches. */
int mvae_iwae_cov_norm(int64_t B, int32_t D, int32_t Sd, int32_t n, const float* x, const float* zsum, float* work,
                       float* out, void* stream);

/* Fused Adam (torch.optim.Adam defaults: betas .9/.999, eps 1e-8, no weight decay — train.py:343) over a flat
 * fp32 parameter bucket.  step is 1-based.  grad_scale multiplies the gradient first. */
int mvae_adam_step(int64_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr,
                   float beta1, float beta2, float eps, int32_t step, float grad_scale, void* stream);

/* Same update with the 1-based step counter on the device: *step_dev is incremented first (in stream order), then
 * used for the bias corrections — nothing step-dependent is baked into launch parameters, so the call can be
 * captured once in a CUDA graph and replayed. */
int mvae_adam_step_dev(int64_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr,
                       float beta1, float beta2, float eps, int32_t* step_dev, float grad_scale, void* stream);

/* torch.nn.utils.clip_grad_norm_(params, max_norm, norm_type=2) over the n <= 4096 scalar gradients selected by mask
 * (mask[i] != 0): the reference clips the "curvature"-named parameters of universal components in every train step
 * (vae.py:161-163).  grad[i] *= min(1, max_norm / (||grad . mask||_2 + 1e-6)) for the selected entries. */
int mvae_clip_grad_norm(int32_t n, float* grad, const float* mask, float max_norm, void* stream);

/* Plain SGD step p -= lr * grad_scale * g (torch.optim.SGD defaults — the curvature optimizers of train.py:346-355). */
int mvae_sgd_step(int64_t n, float* param, const float* grad, float lr, float grad_scale, void* stream);

/* The whole optimizer step of Trainer.build_optimizer / CurvatureOptimizer (train.py:327-360, utils.py:148-180) in one
 * launch: Adam over the flat bucket of n (multiple of 4) network parameters with the step counter on the device
 * (incremented here), SGD with radius_lr on the C raw radii (radius may be NULL; radius_lr 0 = no curvature step;
 * gradient times radius_mask, NULL = ones), and the split-bf16 planes of up to eight weight matrices refreshed from the
 * updated values: matrix t occupies [target_begin[t], target_begin[t] + target_rows[t] * targets[t].cols) of the bucket
 * (begin and cols multiples of 4) and is written to targets[t].  done_counter: one zero-initialised device word; NULL =
 * the step counter is read but NOT advanced: a step may be several launches over disjoint parts of the parameters
 * (one that is ready early can run under the rest of the backward pass), the last of which passes the counter. */
int mvae_opt_step_fused(int64_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr,
                        float beta1, float beta2, float eps, int32_t* step_dev, uint32_t* done_counter, float* radius,
                        const float* gradius, const float* radius_mask, float radius_lr, int32_t C, int32_t n_targets,
                        const int64_t* target_begin, const int32_t* target_rows, const mvae_planes* targets,
                        void* stream);

/* ------------------------------------------------------------------- data-parallel step over peer memory */
/* The reference has no distributed mode.  Data parallelism here shards the batch by rank; the
 * only exchange of a step is the SUM of the flat bucket [network grads (n_net) | radius grads (C) | ELBO statistics
 * (3 + C)] (the reference's ELBO is a sum over the batch, stats.py:200-202).  mvae_dp_step does that exchange and
 * the optimizer update (Adam on the network parameters, SGD on the radii: train.py:327-360, utils.py:148-180, gradient
 * clip of the "curvature" parameters: vae.py:161-163) for one RANGE of the parameter buffer in ONE kernel over NVLink
 * peer memory: reduce-scatter of the gradients with peer loads, Adam on the rank's slice (its moments are the only
 * optimizer state the rank keeps current), all-gather of the new parameters with peer loads (no remote stores: a
 * load completes when its data arrives, a pushed store needs a system-scope fence).  A step is one or more
 * launches over disjoint ranges (a range whose gradient is complete early can be exchanged on a side stream while
 * the backward pass continues; concurrent launches use different channels); exactly one of them owns the tail.
 *
 * Every rank allocates one region with mvae_dp_alloc, exports it (mvae_dp_ipc_export), and opens its peers' regions
 * (mvae_dp_ipc_open) after exchanging the 64-byte handles out of band (e.g. torch.distributed.all_gather_object).
 * bucket[r] / flat[r] / flags[r] are rank r's gradient bucket, parameter buffer and flag area (MVAE_DP_FLAG_BYTES,
 * zero-initialised: [channel][phase][source rank] slots of 128 bytes) as mapped in THIS process. */
#define MVAE_DP_MAX_RANKS 8
#define MVAE_DP_CHANNELS 2
#define MVAE_DP_FLAG_BYTES ((MVAE_DP_CHANNELS + 1) * 2 * MVAE_DP_MAX_RANKS * 128)  /* + the rendezvous slots */
#define MVAE_DP_SYNC_WORDS 528
#define MVAE_DP_HANDLE_BYTES 64
typedef struct mvae_dp_comm {
  int32_t rank, world;
  float* bucket[MVAE_DP_MAX_RANKS];
  float* flat[MVAE_DP_MAX_RANKS];
  uint32_t* flags[MVAE_DP_MAX_RANKS];
} mvae_dp_comm;

int mvae_dp_alloc(size_t bytes, void** dev_ptr);           /* cudaMalloc + zero fill (synchronises once, at set-up) */
int mvae_dp_free(void* dev_ptr);
int mvae_dp_ipc_export(void* dev_ptr, uint8_t* handle_out /* [MVAE_DP_HANDLE_BYTES] */);
int mvae_dp_ipc_open(const uint8_t* handle, void** peer_ptr);
int mvae_dp_ipc_close(void* peer_ptr);

/* One launch of the data-parallel optimizer step over the float range [begin, end) (multiples of 4) of the n_net
 * network parameters.  exp_avg / exp_avg_sq: full-size moment buffers of which this rank updates its slice only.
 * step_dev: device step counter, read as "step - 1" by every launch and incremented by the launch with do_tail != 0
 * (the LAST launch of a step).  That launch also sums the n_tail = C + 3 + C tail entries into tail_out [n_tail + 1]
 * (radius grads, then bce, kl, elbo, kl_c sums; the extra entry is the error word below as a float), scales the
 * radius gradients selected by clip_mask [C] (NULL = none) so that their 2-norm is at most clip_max_norm, and steps
 * radius [C] (may be NULL) with radius_lr (0 = the curvature optimizers do not step) on the summed gradient times
 * radius_mask (NULL = ones).
 * sync_words: MVAE_DP_SYNC_WORDS zero-initialised device words private to this rank (epochs and tickets, the error
 * word, and %globaltimer stamps of the last launch's phases for diagnosis).  Word [8] is a STICKY error:
 * non-zero = a peer did not arrive within MVAE_DP_TIMEOUT_S seconds (environment, default 30); the launch that saw it
 * and every later launch write no parameter.  The host must check it (tail_out[n_tail] carries it with the statistics).
 * targets: weight matrices inside [begin, end) whose split-bf16 planes are refreshed from the gathered parameters.
 * max_ctas: 0 = size the grid for the range; a launch that overlaps other kernels should ask for a few CTAs.
 * Captured in a CUDA graph like any other launch; all ranks must issue the same sequence of launches. */
typedef struct mvae_dp_step_args {
  int64_t n_net, begin, end;
  int32_t channel, do_tail, n_tail, C;
  float* exp_avg;
  float* exp_avg_sq;
  float lr, beta1, beta2, eps;
  int32_t* step_dev;
  float* radius;
  const float* radius_mask;
  const float* clip_mask;
  float radius_lr, clip_max_norm;
  float* tail_out;
  uint32_t* sync_words;
  int32_t max_ctas, n_targets;
  const int64_t* target_begin;
  const int32_t* target_rows;
  const mvae_planes* targets;
} mvae_dp_step_args;
int mvae_dp_step(const mvae_dp_comm* comm, const mvae_dp_step_args* args, void* stream);
/* Returns (on the stream) once every rank has launched it: re-aligns the ranks, e.g. after per-rank work of uneven
 * length that precedes a step (bench.py uses it after its L2 flush, outside the timed region).  Same flags / error
 * word / time limit as mvae_dp_step. */
int mvae_dp_rendezvous(const mvae_dp_comm* comm, uint32_t* sync_words, void* stream);

/* Thin wrappers over the CUDA runtime (linked statically into the library) for the host side of the pipelined epoch
 * loop — the batch loop of Trainer._train_epoch (train.py:197-210) with the host->device copy of batch i+1 and the
 * device->host copy of the statistics overlapped with the kernels: asynchronous copy on a stream (cudaMemcpyDefault),
 * event record, stream-wait-event.  `stream` / `event` are the caller framework's cudaStream_t / cudaEvent_t handles.
 * They exist because the same operations through a Python framework cost 10-20 us of host time EACH (stream context
 * switches), which made the end-to-end step host-bound at 0.2 ms per step. */
int mvae_rt_memcpy_async(void* dst, const void* src, size_t bytes, void* stream);
int mvae_rt_event_record(void* event, void* stream);
int mvae_rt_stream_wait_event(void* stream, void* event);

/* Diagnostics of the fused latent block (scripts/latent_phases.py; not part of the reference-facing surface).  With a
 * non-null `stamps` (device memory, 8 words per CTA of the next launches' grids) every CTA of mvae_latent_forward /
 * mvae_latent_backward records %globaltimer at its phase boundaries; bit 0 of `flags` drops the weight-gradient
 * reductions of the backward kernel (timing experiments only — the gradients are then wrong).  (NULL, 0) restores the
 * production behaviour.  Process-wide, not thread-safe. */
int mvae_debug_latent(unsigned long long* stamps, int32_t flags);
/* The same for mvae_gemm: 16 words per CTA (linear CTA index x + gridDim.x (y + gridDim.y z)) — 0 kernel entry, 1
 * barriers and TMEM set up, 2 predecessor grid complete (griddepcontrol.wait), 3 first operand stage landed, 4 last MMA
 * issued, 5 accumulator complete, 6 first epilogue warp done, 7 CTA done, 8-10 first epilogue chunk (loaded / loss
 * math / complete), 11 all epilogue warps done, 12 bulk stores issued (scripts/gemm_phases.py).  `flags` (timing
 * experiments only, results are then wrong): 1 = no plane staging / stores, 2 = no loss arithmetic, 4 = no prefetch of
 * the epilogue's global operands. */
int mvae_debug_gemm(unsigned long long* stamps, int32_t flags);
/* Timeline mode (scripts/step_timeline.py): with a non-null buffer every following mvae_gemm / mvae_latent_* launch
 * stamps into its OWN region of the buffer (16 words per CTA, regions in launch order) — launches captured into a CUDA
 * graph keep their regions, so one replay of a captured train step yields the start / end times of all its stamped
 * kernels on one clock.  mvae_debug_timeline_log returns the (kind 0 gemm | 1 latent forward | 2 latent backward,
 * CTAs, M|B, N|H, K) records of the launches since the buffer was set; mvae_debug_timeline(NULL) ends the mode. */
int mvae_debug_timeline(unsigned long long* buffer);
int mvae_debug_timeline_log(int32_t* out, int32_t max_entries);

/* Device attributes the host layer needs for grid sizing / reporting. */
int mvae_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

#ifdef __cplusplus
}
#endif
#endif /* MVAE_B200_H_ */
