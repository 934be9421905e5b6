"""Tensor-level wrappers over the C ABI (include/mvae_b200.h).  torch is used for device memory and streams only:
every function enqueues hand-written sm_100a kernels from libmvae_b200.so on torch's current CUDA stream.
All tensors must be contiguous CUDA tensors (float32 unless stated).  No fallbacks."""
import ctypes
from typing import Optional, Sequence

import torch

from . import _lib as L


_LAUNCHES = [0]


def launch_count() -> int:
    """Number of kernel launches issued through this module (each C entry point = its kernels; graph replays are
    added by the caller through add_launches)."""
    return _LAUNCHES[0]


def add_launches(n: int) -> None:
    _LAUNCHES[0] += n


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise L.MvaeError(f"{name}: expected a CUDA tensor (the mvae_b200 path has no CPU fallback)")
    if t.dtype != torch.float32:
        raise L.MvaeError(f"{name}: expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def parse_signature(sig: str):
    """'2h3,s2,e2' -> ([types], [dims]); grammar of mt/mvae/utils.py:78-140 (letters e,h,s,p,d,u; 'c' unsupported)."""
    types, dims = [], []
    for tok in sig.lower().strip().split(","):
        tok = tok.strip().split("-")[0]
        i = 0
        while i < len(tok) and tok[i].isdigit():
            i += 1
        mult = int(tok[:i]) if i else 1
        j = i
        while j < len(tok) and tok[j].isalpha():
            j += 1
        letter, dim = tok[i:j], int(tok[j:])
        if letter not in L.TYPE_OF_LETTER:
            raise ValueError(f"unsupported component letter {letter!r} in {sig!r}")
        for _ in range(mult):
            types.append(L.TYPE_OF_LETTER[letter])
            dims.append(dim)
    return types, dims


def make_desc(sig_or_types, dims: Optional[Sequence[int]] = None, scalar_parametrization: bool = False) -> L.PmDesc:
    if isinstance(sig_or_types, str):
        types, dims = parse_signature(sig_or_types)
    else:
        types = list(sig_or_types)
    return L.make_desc(types, list(dims), scalar_parametrization)


# ------------------------------------------------------------------------------------------ product manifold
def pm_forward(desc: L.PmDesc, ml, eps, radius, want_mu_sigma: bool = False, flag: Optional[torch.Tensor] = None,
               out: Optional[dict] = None):
    """Fused manifold + Wrapped-Normal forward (mvae_pm_forward).  Returns dict(z, kl[, mu, sigma])."""
    ml, eps, radius = _f32(ml, "ml"), _f32(eps, "eps"), _f32(radius, "radius")
    B = ml.shape[0]
    assert ml.shape[1] == desc.ld_ml and eps.shape == (B, desc.ld_eps) and radius.numel() == desc.C
    out = out or {}
    dev = ml.device
    z = out.get("z") if out.get("z") is not None else torch.empty(B, desc.ld_z, device=dev)
    kl = out.get("kl") if out.get("kl") is not None else torch.empty(B, desc.C, device=dev)
    mu = sigma = None
    if want_mu_sigma:
        mu = out.get("mu") if out.get("mu") is not None else torch.empty(B, desc.ld_z, device=dev)
        sigma = out.get("sigma") if out.get("sigma") is not None else torch.empty(B, desc.ld_eps, device=dev)
    rc = L.lib().mvae_pm_forward(ctypes.byref(desc), B, _ptr(ml), _ptr(eps), _ptr(radius), _ptr(z), _ptr(kl), _ptr(mu),
                                 _ptr(sigma), _ptr(flag), _stream())
    L.check(rc, "mvae_pm_forward")
    _LAUNCHES[0] += 1
    res = {"z": z, "kl": kl}
    if want_mu_sigma:
        res.update(mu=mu, sigma=sigma)
    return res


def pm_backward(desc: L.PmDesc, ml, eps, radius, gz, gkl: Optional[torch.Tensor] = None, gkl_scalar: float = 1.0,
                gml: Optional[torch.Tensor] = None, gradius: Optional[torch.Tensor] = None):
    """Backward of pm_forward by recomputation (mvae_pm_backward).  gradius is ACCUMULATED (zeroed here if created)."""
    ml, eps, radius, gz = _f32(ml, "ml"), _f32(eps, "eps"), _f32(radius, "radius"), _f32(gz, "gz")
    gkl = None if gkl is None else _f32(gkl, "gkl")
    B = ml.shape[0]
    if gml is None:
        gml = torch.empty_like(ml)
    if gradius is None:
        gradius = torch.zeros(desc.C, device=ml.device)
    rc = L.lib().mvae_pm_backward(ctypes.byref(desc), B, _ptr(ml), _ptr(eps), _ptr(radius), _ptr(gz), _ptr(gkl),
                                  float(gkl_scalar), _ptr(gml), _ptr(gradius), _stream())
    L.check(rc, "mvae_pm_backward")
    _LAUNCHES[0] += 1
    return gml, gradius


# ------------------------------------------------------------------------------------------ standalone ops
def manifold_op(op: int, manifold: int, n: int, x, y=None, radius=None):
    x = _f32(x, "x")
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    B = x2.shape[0]
    d = n + 1 if manifold in (L.HYPERBOLOID, L.SPHERE) else n
    out_w = {L.OP_EXP_MAP_MU0: d, L.OP_DISTANCE: 1, L.OP_LOGDET: 1, L.OP_TO_POINCARE: n,
             L.OP_FROM_POINCARE: n + 1}.get(op, d)
    y2 = None
    if y is not None:
        y2 = _f32(y, "y").reshape(B, -1)
    out = torch.empty(B, out_w, device=x.device)
    rc = L.lib().mvae_manifold_op(op, manifold, n, B, _ptr(x2), _ptr(y2), _ptr(radius), _ptr(out), _stream())
    L.check(rc, "mvae_manifold_op")
    _LAUNCHES[0] += 1
    return out.reshape(*lead, out_w)


def wn_rsample(manifold: int, n: int, loc, scale, eps, radius):
    loc, scale, eps = _f32(loc, "loc"), _f32(scale, "scale"), _f32(eps, "eps")
    B = loc.shape[0]
    z, u, v = torch.empty_like(loc), torch.empty_like(loc), torch.empty_like(scale)
    rc = L.lib().mvae_wn_rsample(manifold, n, B, _ptr(loc), _ptr(scale), _ptr(eps), _ptr(radius), _ptr(z), _ptr(u),
                                 _ptr(v), _stream())
    L.check(rc, "mvae_wn_rsample")
    _LAUNCHES[0] += 1
    return z, u, v


def wn_log_prob_from_parts(manifold: int, n: int, loc, scale, z, u, v, radius):
    args = [_f32(t, "arg") for t in (loc, scale, z, u, v)]
    B = args[0].shape[0]
    logp = torch.empty(B, device=args[0].device)
    rc = L.lib().mvae_wn_log_prob_from_parts(manifold, n, B, *[_ptr(t) for t in args], _ptr(radius), _ptr(logp),
                                             _stream())
    L.check(rc, "mvae_wn_log_prob_from_parts")
    _LAUNCHES[0] += 1
    return logp


def wn_log_prob(manifold: int, n: int, loc, scale, z, radius):
    args = [_f32(t, "arg") for t in (loc, scale, z)]
    B = args[0].shape[0]
    logp = torch.empty(B, device=args[0].device)
    rc = L.lib().mvae_wn_log_prob(manifold, n, B, *[_ptr(t) for t in args], _ptr(radius), _ptr(logp), _stream())
    L.check(rc, "mvae_wn_log_prob")
    _LAUNCHES[0] += 1
    return logp


# ------------------------------------------------------------------------------------------ losses / optimizer
def recon_loss(kind: str, logits, x, want_grad: bool = False, out: Optional[tuple] = None):
    """kind 'bce' | 'nll' -> (rowsum [B], glogits | None)  (mvae_recon_loss).  out = (rowsum, glogits | None) writes
    into the caller's buffers."""
    logits, x = _f32(logits, "logits"), _f32(x, "x")
    B, D = logits.shape
    if out is not None:
        rs, g = out
    else:
        rs = torch.empty(B, device=logits.device)
        g = torch.empty_like(logits) if want_grad else None
    rc = L.lib().mvae_recon_loss(0 if kind == "bce" else 1, B, D, _ptr(logits), _ptr(x), _ptr(rs), _ptr(g), _stream())
    L.check(rc, "mvae_recon_loss")
    _LAUNCHES[0] += 1
    return rs, g


def elbo_reduce(bce, kl, beta: float, out: Optional[torch.Tensor] = None):
    """-> float32 [3 + C] = [bce_sum, kl_sum, elbo, kl_c sums]  (mvae_elbo_reduce)."""
    bce, kl = _f32(bce, "bce"), _f32(kl, "kl")
    B, C = kl.shape
    if out is None:
        out = torch.empty(3 + C, device=kl.device)
    rc = L.lib().mvae_elbo_reduce(B, C, _ptr(bce), _ptr(kl), float(beta), _ptr(out), _stream())
    L.check(rc, "mvae_elbo_reduce")
    _LAUNCHES[0] += 1
    return out


# ------------------------------------------------------------------------------------------ input pipeline
def binarize(src_u8: torch.Tensor, x: Optional[torch.Tensor] = None, planes: Optional["PlaneBuf"] = None,
             u: Optional[torch.Tensor] = None, seed: int = 0, offset_dev: Optional[torch.Tensor] = None,
             dynamic: bool = True, invert: bool = False):
    """uint8 grayscale batch [B, D] -> binarised fp32 x and / or its bf16 operand plane (mvae_binarize).
    dynamic: x/255 > u with u supplied or drawn in the kernel (Philox; offset_dev = int64 device step counter);
    else the fixed evaluation threshold 0.5."""
    if not src_u8.is_cuda or src_u8.dtype != torch.uint8:
        raise L.MvaeError("binarize: expected a CUDA uint8 tensor")
    B, D = src_u8.shape
    if x is None and planes is None:
        x = torch.empty(B, D, device=src_u8.device)
    ps = planes.struct() if planes is not None else None
    rc = L.lib().mvae_binarize(_ptr(src_u8), src_u8.stride(0), B, D, 0 if dynamic else 1, int(invert),
                               _ptr(None if u is None else _f32(u, "u")), int(seed) & (2**64 - 1), _ptr(offset_dev),
                               _ptr(x), x.stride(0) if x is not None else 0,
                               ctypes.byref(ps) if ps is not None else None, _stream())
    L.check(rc, "mvae_binarize")
    _LAUNCHES[0] += 1
    return x


# ------------------------------------------------------------------------------------------ convolutions
def conv_im2col(src: "PlaneBuf", B: int, H: int, W: int, C: int, dst: "PlaneBuf", ones_col: bool = False):
    """4x4 / stride 2 / pad 1 patches of a channels-last plane buffer [B*H*W, C] -> [B*(H/2)*(W/2), 16 C] (mvae_conv_im2col)."""
    ss, ds = src.struct(rows=B * H * W), dst.struct(rows=B * (H // 2) * (W // 2))
    rc = L.lib().mvae_conv_im2col(ctypes.byref(ss), B, H, W, C, ctypes.byref(ds), int(ones_col), _stream())
    L.check(rc, "mvae_conv_im2col")
    _LAUNCHES[0] += 1


def conv_col2im(cols: torch.Tensor, B: int, H: int, W: int, C: int, bias=None, act: int = 0,
                mask: Optional["PlaneBuf"] = None, out_planes: Optional["PlaneBuf"] = None, out_f32=None):
    """Adjoint gather of conv_im2col: fp32 tap columns [B*H*W, 16 C] -> [B*2H*2W, C] (+ bias, act 1 relu / 2 mask) as
    planes and / or fp32 (mvae_conv_col2im)."""
    cols = _f32(cols, "cols")
    rows = B * H * W * 4
    ms = mask.struct(rows=rows) if mask is not None else None
    os_ = out_planes.struct(rows=rows) if out_planes is not None else None
    rc = L.lib().mvae_conv_col2im(_ptr(cols), cols.stride(0), B, H, W, C, _ptr(bias), act,
                                  ctypes.byref(ms) if ms is not None else None,
                                  ctypes.byref(os_) if os_ is not None else None, _ptr(out_f32),
                                  out_f32.stride(0) if out_f32 is not None else 0, _stream())
    L.check(rc, "mvae_conv_col2im")
    _LAUNCHES[0] += 1


def permute_sc(src, dst, B: int, S: int, C: int, to_nhwc: bool):
    """[B, C*S] rows in (c, s) order <-> channels-last [B*S, C] rows (mvae_permute_sc); src / dst: float32 tensors or
    PlaneBufs (all of the destination's planes are moved)."""
    if isinstance(src, PlaneBuf):
        assert isinstance(dst, PlaneBuf) and dst.planes <= src.planes
        rc = L.lib().mvae_permute_sc(2, src.t.data_ptr(), src.ld, src.rows * src.ld, dst.t.data_ptr(), dst.ld,
                                     dst.rows * dst.ld, dst.planes, B, S, C, int(to_nhwc), _stream())
    else:
        src, dst = _f32(src, "src"), _f32(dst, "dst")
        rc = L.lib().mvae_permute_sc(4, src.data_ptr(), src.stride(0), 0, dst.data_ptr(), dst.stride(0), 0, 1, B, S, C,
                                     int(to_nhwc), _stream())
    L.check(rc, "mvae_permute_sc")
    _LAUNCHES[0] += 1


def colsum(src, M: int, C: int, out: torch.Tensor):
    """out[c] += sum_m src[m, c] for a float32 matrix or a PlaneBuf (mvae_colsum)."""
    if isinstance(src, PlaneBuf):
        ps = src.struct(rows=M)
        rc = L.lib().mvae_colsum(None, ctypes.byref(ps), M, C, 0, _ptr(out), _stream())
    else:
        src = _f32(src, "src")
        rc = L.lib().mvae_colsum(_ptr(src), None, M, C, src.stride(0), _ptr(out), _stream())
    L.check(rc, "mvae_colsum")
    _LAUNCHES[0] += 1


def step_prologue(eps: Optional[torch.Tensor], seed: int, counter_dev: Optional[torch.Tensor], zero=()):
    """Head of a train step in one launch (mvae_step_prologue): standard-normal noise into `eps` (None: the caller
    supplies it) drawn with Philox at offset *counter_dev, and zero fill of the float tensors in `zero`."""
    zero = [t for t in zero if t is not None and t.numel() > 0]
    for t in zero:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise L.MvaeError("step_prologue: zero targets must be contiguous float32 CUDA tensors")
    n = len(zero)
    ptrs = (ctypes.c_void_p * max(n, 1))(*[t.data_ptr() for t in zero])
    counts = (ctypes.c_int64 * max(n, 1))(*[t.numel() for t in zero])
    if eps is not None:
        eps = _f32(eps, "eps")
    rc = L.lib().mvae_step_prologue(_ptr(eps), eps.numel() if eps is not None else 0, int(seed) & (2**64 - 1),
                                    _ptr(counter_dev), n, ptrs, counts, _stream())
    L.check(rc, "mvae_step_prologue")
    _LAUNCHES[0] += 1


def ring_push(src: torch.Tensor, ring: torch.Tensor, counter_dev: torch.Tensor):
    """ring[counter % capacity] = src; counter += 1 on the device (mvae_ring_push)."""
    rc = L.lib().mvae_ring_push(_ptr(src), src.numel(), _ptr(ring), ring.shape[0], _ptr(counter_dev), _stream())
    L.check(rc, "mvae_ring_push")
    _LAUNCHES[0] += 1


def counter_add(counter_dev: torch.Tensor, inc: int = 1):
    rc = L.lib().mvae_counter_add(_ptr(counter_dev), int(inc), _stream())
    L.check(rc, "mvae_counter_add")
    _LAUNCHES[0] += 1


# ------------------------------------------------------------------------------------------ IWAE log-likelihood
def iwae_latent(desc: L.PmDesc, ml, eps, radius, z, diff, zsum: Optional[torch.Tensor] = None):
    """z [ns, B, ld_z] and diff [ns, B] = sum_c (log q_c - log p_c) for ns samples per row of ml [B, ld_ml]
    (mvae_iwae_latent); zsum [B, ld_z] accumulates sum_s z."""
    ns, B = eps.shape[0], ml.shape[0]
    assert eps.shape == (ns, B, desc.ld_eps) and z.shape == (ns, B, desc.ld_z) and diff.shape == (ns, B)
    rc = L.lib().mvae_iwae_latent(ctypes.byref(desc), B, ns, _ptr(_f32(ml, "ml")), _ptr(_f32(eps, "eps")),
                                  _ptr(radius), _ptr(z), _ptr(diff), _ptr(zsum), _stream())
    L.check(rc, "mvae_iwae_latent")
    _LAUNCHES[0] += 1


def iwae_reduce(recon, diff, log_p_x: Optional[torch.Tensor] = None, mi: Optional[torch.Tensor] = None):
    """(logsumexp_s(-recon - diff) - log n, logsumexp_s(diff) - log n) over [n, B] inputs (mvae_iwae_reduce)."""
    n, B = diff.shape
    if log_p_x is None:
        log_p_x = torch.empty(B, device=diff.device)
    if mi is None:
        mi = torch.empty(B, device=diff.device)
    rc = L.lib().mvae_iwae_reduce(n, B, _ptr(_f32(recon, "recon")), _ptr(_f32(diff, "diff")), _ptr(log_p_x), _ptr(mi),
                                  _stream())
    L.check(rc, "mvae_iwae_reduce")
    _LAUNCHES[0] += 1
    return log_p_x, mi


def iwae_cov_norm(x, zsum, n: int, out: Optional[torch.Tensor] = None):
    """cov_norm of vae.py:119-121 from x [B, D] and zsum [B, Sd] = sum over the n samples of z (mvae_iwae_cov_norm)."""
    B, D = x.shape
    Sd = zsum.shape[1]
    work = torch.empty(B * Sd + D * Sd, device=x.device)
    if out is None:
        out = torch.empty(1, device=x.device)
    rc = L.lib().mvae_iwae_cov_norm(B, D, Sd, n, _ptr(_f32(x, "x")), _ptr(_f32(zsum, "zsum")), _ptr(work), _ptr(out),
                                    _stream())
    L.check(rc, "mvae_iwae_cov_norm")
    _LAUNCHES[0] += 3
    return out


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    rc = L.lib().mvae_adam_step(param.numel(), _ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), float(lr),
                                float(beta1), float(beta2), float(eps), int(step), float(grad_scale), _stream())
    L.check(rc, "mvae_adam_step")
    _LAUNCHES[0] += 1


def adam_step_dev(param, grad, exp_avg, exp_avg_sq, lr, step_dev, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    """Adam with the step counter on the device (int32 tensor, incremented by the call): CUDA-graph replayable."""
    rc = L.lib().mvae_adam_step_dev(param.numel(), _ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), float(lr),
                                    float(beta1), float(beta2), float(eps), _ptr(step_dev), float(grad_scale),
                                    _stream())
    L.check(rc, "mvae_adam_step_dev")
    _LAUNCHES[0] += 2


def clip_grad_norm(grad, mask, max_norm: float = 1.0):
    """In-place clip_grad_norm_ (2-norm) over the entries of `grad` selected by `mask` (mvae_clip_grad_norm)."""
    rc = L.lib().mvae_clip_grad_norm(grad.numel(), _ptr(grad), _ptr(mask), float(max_norm), _stream())
    L.check(rc, "mvae_clip_grad_norm")
    _LAUNCHES[0] += 1


def sgd_step(param, grad, lr, grad_scale=1.0):
    rc = L.lib().mvae_sgd_step(param.numel(), _ptr(param), _ptr(grad), float(lr), float(grad_scale), _stream())
    L.check(rc, "mvae_sgd_step")
    _LAUNCHES[0] += 1


# ------------------------------------------------------------------------------------------ planes + GEMM
def _round_up(x, m):
    return (x + m - 1) // m * m


class PlaneBuf:
    """Split-bf16 operand planes [planes, rows, ld] of an fp32 matrix [rows, cols] (mvae_planes).
    `ones_col=True` reserves column `cols` and fills it with 1.0: read as part of an MN-major wgrad operand it
    yields the bias gradient (column sums of the other operand) for free."""

    def __init__(self, rows: int, cols: int, planes: int = 2, device="cuda", ones_col: bool = False):
        self.rows, self.cols, self.planes, self.ones_col = rows, cols, planes, ones_col
        self.ld = _round_up(cols + (1 if ones_col else 0), 8)
        self.t = torch.zeros(planes, rows, self.ld, dtype=torch.bfloat16, device=device)
        if ones_col:
            self.t[0, :, cols] = 1.0

    def struct(self, rows: Optional[int] = None, cols: Optional[int] = None, planes: Optional[int] = None) -> L.Planes:
        """`planes` < self.planes reads only the leading planes (a lower-precision view of the same buffer)."""
        return L.Planes(self.t.data_ptr(), self.rows * self.ld, rows if rows is not None else self.rows,
                        cols if cols is not None else self.cols, self.ld,
                        min(planes, self.planes) if planes is not None else self.planes)

    def to_float(self) -> torch.Tensor:
        return self.t.float().sum(0)[:, :self.cols]

    def plane0(self) -> torch.Tensor:
        return self.t[0]


def split_planes(src: torch.Tensor, dst: Optional[PlaneBuf] = None, dst_t: Optional[PlaneBuf] = None):
    src = _f32(src, "src")
    R, K = src.shape
    ds = dst.struct(rows=R) if dst is not None else None  # the leading R rows of a (possibly larger) plane buffer
    dt = dst_t.struct() if dst_t is not None else None
    rc = L.lib().mvae_split_planes(_ptr(src), src.stride(0), R, K, ctypes.byref(ds) if ds is not None else None,
                                   ctypes.byref(dt) if dt is not None else None, _stream())
    L.check(rc, "mvae_split_planes")
    _LAUNCHES[0] += 1


def gemm(a: PlaneBuf, b: PlaneBuf, M: int, N: int, K: int, a_major: int = L.K_MAJOR, b_major: int = L.K_MAJOR,
         epilogue: int = L.EPI_STORE, split_k: int = 1, bias=None, out_f32=None, ld_out: Optional[int] = None,
         out_col=None, col_split: int = -1, out_planes: Optional[PlaneBuf] = None, aux=None, mask: Optional[PlaneBuf] = None,
         rowsum=None, a_planes: Optional[int] = None, b_planes: Optional[int] = None, tile: Optional[tuple] = None,
         aux_rows: int = 0):
    """D[M,N] = sum_k A[m,k] B[n,k] on tcgen05 tensor cores with a fused epilogue (mvae_gemm).
    tile = (BLOCK_N, CTAs per SM[, split_k]) overrides the automatic tile policy (0 = automatic)."""
    g = L.GemmArgs()
    if tile is not None:
        g.tile_n, g.ctas_per_sm = int(tile[0]), int(tile[1])
        if len(tile) > 2 and split_k != 1 and epilogue == L.EPI_STORE:
            split_k = int(tile[2])
    g.a = a.struct(rows=M if a_major == L.K_MAJOR else K, cols=K if a_major == L.K_MAJOR else M, planes=a_planes)
    g.b = b.struct(rows=N if b_major == L.K_MAJOR else K, cols=K if b_major == L.K_MAJOR else N, planes=b_planes)
    g.a_major, g.b_major = a_major, b_major
    g.M, g.N, g.K = M, N, K
    g.epilogue, g.split_k = epilogue, split_k
    g.bias = bias.data_ptr() if bias is not None else None
    if out_f32 is not None:
        g.out_f32 = out_f32.data_ptr()
        g.ld_out = ld_out if ld_out is not None else out_f32.stride(0)
    g.out_col = out_col.data_ptr() if out_col is not None else None
    g.col_split = col_split
    if out_planes is not None:
        g.out_planes = out_planes.struct()
    if aux is not None:
        g.aux = aux.data_ptr()
        g.ld_aux = aux.stride(0)
        g.aux_rows = int(aux_rows)
    if mask is not None:
        g.mask = mask.t.data_ptr()
        g.ld_mask = mask.ld
    g.rowsum = rowsum.data_ptr() if rowsum is not None else None
    rc = L.lib().mvae_gemm(ctypes.byref(g), _stream())
    L.check(rc, "mvae_gemm")
    _LAUNCHES[0] += 1


# ------------------------------------------------------------------------------------------ skinny dense layers
ACT_NONE, ACT_RELU, ACT_MASK = 0, 1, 2


def _wide(x):
    """(f32 ptr, ld, planes struct ptr) of a wide operand given as fp32 tensor or PlaneBuf / (PlaneBuf, planes)."""
    if isinstance(x, torch.Tensor):
        return _ptr(x), x.stride(0), None
    buf, planes = x if isinstance(x, tuple) else (x, None)
    st = buf.struct(planes=planes)
    return None, 0, st


def latent_forward(desc: L.PmDesc, h: torch.Tensor, Wh, bh, eps, radius, Wd0, bd0, ml, z, kl, dd: "PlaneBuf",
                   flag: Optional[torch.Tensor] = None, draw: Optional[tuple] = None, zero=()):
    """heads + product-manifold forward + fc_d0/relu in one launch (mvae_latent_forward); h: fp32 [B, H].
    draw = (seed, counter_dev): eps is DRAWN by this launch (the Philox stream of step_prologue) and written out;
    zero: float tensors this launch zero-fills (mvae_latent_forward_ex: the head of a train step taken along)."""
    B, H = ml.shape[0], Wh.shape[1]
    ds = dd.struct()
    zero = [t for t in zero if t is not None and t.numel() > 0]
    if draw is None and not zero:
        rc = L.lib().mvae_latent_forward(ctypes.byref(desc), B, H, _ptr(h), h.stride(0), _ptr(Wh), _ptr(bh), _ptr(eps),
                                         _ptr(radius), _ptr(Wd0), _ptr(bd0), _ptr(ml), _ptr(z), _ptr(kl),
                                         ctypes.byref(ds), _ptr(flag), _stream())
        L.check(rc, "mvae_latent_forward")
    else:
        if len(zero) > 4:
            raise L.MvaeError("latent_forward: at most four zero-fill spans")
        for t in zero:
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise L.MvaeError("latent_forward: zero targets must be contiguous float32 CUDA tensors")
        pro = L.LatentPrologue()
        pro.draw_eps = 1 if draw is not None else 0
        if draw is not None:
            pro.seed = int(draw[0]) & (2**64 - 1)
            pro.counter_dev = draw[1].data_ptr() if draw[1] is not None else None
        pro.n_zero = len(zero)
        for i, t in enumerate(zero):
            pro.zero_ptr[i] = t.data_ptr()
            pro.zero_n[i] = t.numel()
        rc = L.lib().mvae_latent_forward_ex(ctypes.byref(desc), B, H, _ptr(h), h.stride(0), _ptr(Wh), _ptr(bh),
                                            _ptr(eps), _ptr(radius), _ptr(Wd0), _ptr(bd0), _ptr(ml), _ptr(z), _ptr(kl),
                                            ctypes.byref(ds), _ptr(flag), ctypes.byref(pro), _stream())
        L.check(rc, "mvae_latent_forward_ex")
    _LAUNCHES[0] += 1


def latent_backward(desc: L.PmDesc, gdd: torch.Tensor, h: torch.Tensor, Wh, Wd0, ml, eps, radius, z,
                    gkl_scalar: float, gh: "PlaneBuf", gWd0, gbd0, gWh, gbh, gradius):
    """fc_d0 dgrad/wgrad + product-manifold backward + heads dgrad/wgrad in one launch (mvae_latent_backward).
    The gradient outputs are ACCUMULATED."""
    B, H = ml.shape[0], Wh.shape[1]
    os_ = gh.struct()
    rc = L.lib().mvae_latent_backward(ctypes.byref(desc), B, H, _ptr(gdd), gdd.stride(0), _ptr(h), h.stride(0), _ptr(Wh),
                                      _ptr(Wd0),
                                      _ptr(ml), _ptr(eps), _ptr(radius), _ptr(z), float(gkl_scalar), ctypes.byref(os_),
                                      _ptr(gWd0), _ptr(gbd0), _ptr(gWh), _ptr(gbh), _ptr(gradius), _stream())
    L.check(rc, "mvae_latent_backward")
    _LAUNCHES[0] += 1


def skinny_rowdot(a, W, w_stride_n: int, w_stride_k: int, K: int, N: int, bias, out):
    """out[b, n] = sum_k a[b, k] W(n, k) + bias[n]  (mvae_skinny_rowdot); a: fp32 [B, >=K] or planes."""
    af, lda, ap = _wide(a)
    B = out.shape[0]
    rc = L.lib().mvae_skinny_rowdot(B, K, N, af, lda, ctypes.byref(ap) if ap is not None else None, _ptr(W), w_stride_n,
                                    w_stride_k, _ptr(bias), _ptr(out), out.stride(0), _stream())
    L.check(rc, "mvae_skinny_rowdot")
    _LAUNCHES[0] += (N + 15) // 16


def skinny_expand(a, W, w_stride_n: int, w_stride_k: int, K: int, N: int, bias=None, act: int = ACT_NONE,
                  mask: Optional[PlaneBuf] = None, out_planes: Optional[PlaneBuf] = None, out_f32=None):
    """out[b, n] = act(sum_k a[b, k] W(n, k) + bias[n])  (mvae_skinny_expand); a: fp32 [B, K]."""
    B = a.shape[0]
    op = out_planes.struct() if out_planes is not None else None
    rc = L.lib().mvae_skinny_expand(B, K, N, _ptr(a), a.stride(0), _ptr(W), w_stride_n, w_stride_k, _ptr(bias), act,
                                    ctypes.c_void_p(mask.t.data_ptr()) if mask is not None else None,
                                    mask.ld if mask is not None else 0, ctypes.byref(op) if op is not None else None,
                                    _ptr(out_f32), out_f32.stride(0) if out_f32 is not None else 0, _stream())
    L.check(rc, "mvae_skinny_expand")
    _LAUNCHES[0] += 1


def skinny_wgrad(small, S: int, wide, Wd: int, out, out_stride_s: int, out_stride_w: int, small_ones: bool = False,
                 out_row=None, out_col=None, col_split: int = -1):
    """out(s, w) += sum_b small[b, s] wide[b, w]  (mvae_skinny_wgrad, accumulating)."""
    wf, ldw, wp = _wide(wide)
    B = small.shape[0]
    rc = L.lib().mvae_skinny_wgrad(B, S, Wd, _ptr(small), small.stride(0), int(small_ones), wf, ldw,
                                   ctypes.byref(wp) if wp is not None else None, _ptr(out), out_stride_s, out_stride_w,
                                   _ptr(out_row), _ptr(out_col), col_split, _stream())
    L.check(rc, "mvae_skinny_wgrad")
    _LAUNCHES[0] += 1


def dp_step(comm: "L.DpComm", n_net: int, begin: int, end: int, channel: int, do_tail: bool, n_tail: int, C: int,
            exp_avg, exp_avg_sq, lr: float, beta1: float, beta2: float, eps: float, step_dev, radius, radius_lr: float,
            radius_mask, clip_mask, clip_max_norm: float, tail_out, sync_words, targets=(), max_ctas: int = 0):
    """Gradient reduce-scatter + Adam + parameter all-gather over peer memory + weight-plane refresh for the float
    range [begin, end) of the parameter buffer, one kernel (mvae_dp_step).  targets: list of (flat offset, rows,
    PlaneBuf) inside the range."""
    nt = len(targets)
    begins = (ctypes.c_int64 * max(nt, 1))(*[t[0] for t in targets])
    rows = (ctypes.c_int32 * max(nt, 1))(*[t[1] for t in targets])
    planes = (L.Planes * max(nt, 1))(*[t[2].struct() for t in targets])
    vp = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    a = L.DpStepArgs(n_net=n_net, begin=begin, end=end, channel=channel, do_tail=int(bool(do_tail)), n_tail=n_tail, C=C,
                     exp_avg=vp(exp_avg), exp_avg_sq=vp(exp_avg_sq), lr=lr, beta1=beta1, beta2=beta2, eps=eps,
                     step_dev=vp(step_dev), radius=vp(radius), radius_mask=vp(radius_mask), clip_mask=vp(clip_mask),
                     radius_lr=radius_lr, clip_max_norm=clip_max_norm, tail_out=vp(tail_out),
                     sync_words=vp(sync_words), max_ctas=max_ctas, n_targets=nt, target_begin=begins, target_rows=rows,
                     targets=planes)
    rc = L.lib().mvae_dp_step(ctypes.byref(comm), ctypes.byref(a), _stream())
    L.check(rc, "mvae_dp_step")
    _LAUNCHES[0] += 1


def opt_step_fused(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step_dev, done_counter, radius, gradius,
                   radius_mask, radius_lr: float, targets):
    """Adam + radii SGD + weight-plane refresh + step counter in one launch (mvae_opt_step_fused).
    targets: list of (flat offset, rows, PlaneBuf)."""
    nt = len(targets)
    begins = (ctypes.c_int64 * max(nt, 1))(*[t[0] for t in targets])
    rows = (ctypes.c_int32 * max(nt, 1))(*[t[1] for t in targets])
    planes = (L.Planes * max(nt, 1))(*[t[2].struct() for t in targets])
    C = radius.numel() if radius is not None else 0
    rc = L.lib().mvae_opt_step_fused(param.numel(), _ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), lr, beta1,
                                     beta2, eps, _ptr(step_dev), _ptr(done_counter), _ptr(radius), _ptr(gradius),
                                     _ptr(radius_mask), radius_lr, C, nt, begins, rows, planes, _stream())
    L.check(rc, "mvae_opt_step_fused")
    _LAUNCHES[0] += 1
