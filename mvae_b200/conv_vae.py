"""FusedConvolutionalVAE — drop-in for the reference's `ConvolutionalVAE` (mt/mvae/models/conv_vae.py:28-79; BASELINE
cfg5: CIFAR, --architecture=conv, --h_dim=8192) on the same kernels as FusedFeedForwardVAE.

  encoder  e0, e1, e2: nn.Conv2d(k4, s2, p1) + relu (:47-49,62-64)     im2col gather (mvae_conv_im2col) -> tcgen05 GEMM
                                                                      with the bias + relu epilogue -> planes
  flatten  x.view(bs, -1) of [B, 512, 4, 4] (:65)                      one layout kernel (mvae_permute_sc): activations
                                                                      are channels-last here, the heads' weights keep the
                                                                      reference's (c, y, x) column order
  heads / manifold chain / KL                                         tcgen05 GEMM (K = 8192, one CTA per 64-wide K
                                                                      slice) + the fused product-manifold kernels
  decoder  d0: nn.Linear(z, 2048) + relu (:52,72-73)                   tcgen05 GEMM, layout kernel
           d1, d2, d3: nn.ConvTranspose2d(k4, s2, p1) (+ relu) (:53-55,74-76)
                                                                      tcgen05 GEMM -> fp32 tap columns -> col2im gather
                                                                      (mvae_conv_col2im: + bias, relu, planes)
  loss     BCE-with-logits on real-valued targets, row sums (image_reconstruction.py:142-143, vae.py:131)
                                                                      mvae_recon_loss on channels-last logits / targets
  backward every GEMM's dgrad / wgrad through the same kernel (MN-major reads, no transposed copies), the two gathers
           swapped (the adjoint of im2col is col2im), bias gradients from the ones column / mvae_colsum
  update   the same fused optimizer kernels over the flat buffer

Parameters keep the reference's names and SHAPES (`e0.weight` [64, 3, 4, 4], `d1.weight` [128, 256, 4, 4], ...); the
flat buffer stores conv filters with the channel innermost ([Co, ky, kx, Ci] / [Ci, ky, kx, Co]) — the GEMM's layout —
and the parameters are permuted views of it (FusedFeedForwardVAE._flatten)."""
from typing import Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib as L
from . import ops
from .vae import FusedFeedForwardVAE, _Workspace

# (name, Cin, Cout, input side) of the three stride-2 stages on each side
_ENC = (("e0", 3, 64, 32), ("e1", 64, 128, 16), ("e2", 128, 512, 8))
_DEC = (("d1", 128, 256, 4), ("d2", 256, 64, 8), ("d3", 64, 3, 16))


class _ConvWorkspace(_Workspace):
    """Per-batch-size buffers of the convolutional step (allocated once; fixed addresses for CUDA graphs)."""

    def __init__(self, m: "FusedConvolutionalVAE", B: int) -> None:
        dev, P, Sn, Sd, C = m.device, m.desc.ld_ml, m.desc.ld_eps, m.desc.ld_z, m.desc.C
        f = dict(device=dev, dtype=torch.float32)
        PB = ops.PlaneBuf
        self.B = B
        self.xbuf = [torch.zeros(B, 3072, **f), torch.zeros(B, 3072, **f)]
        self.slot = 0
        self.x8buf = None
        self._u8 = [False, False]
        self.bin_ctr = m._bin_ctr
        self.eps = torch.zeros(B, Sn, **f)
        self.flag = torch.zeros(1, device=dev, dtype=torch.int32)
        # ---- encoder (channels-last): image, patch matrices A_i (+ ones column), activations a_i ----
        self.x_cl = torch.zeros(B, 3072, **f)            # targets in the order of the channels-last logits
        self.xp = PB(B * 1024, 3, 3, dev)                # real-valued pixels: 3 planes
        self.A = [PB(B * (s // 2) ** 2, 16 * ci, 3, dev, ones_col=True) for _, ci, _, s in _ENC]
        self.a = [PB(B * (s // 2) ** 2, co, 3, dev) for _, _, co, s in _ENC]
        self.hp = PB(B, 8192, 3, dev, ones_col=True)     # flattened (c, y, x): what fc_mean / fc_logvar read
        self.ml = torch.zeros(B, P, **f)
        self.z = torch.zeros(B, Sd, **f)
        self.kl = torch.zeros(B, C, **f)
        self.mu = torch.zeros(B, Sd, **f)
        self.sigma = torch.zeros(B, Sn, **f)
        self.zp = PB(B, Sd, 3, dev, ones_col=True)
        # ---- decoder ----
        self.ddf = PB(B, 2048, 3, dev)                   # relu(d0 z) in the reference's (c, y, x) order
        self.dd = PB(B * 16, 128, 3, dev)                # ... channels-last: input of d1
        self.b = [PB(B * (2 * s) ** 2, co, 3, dev) for _, _, co, s in _DEC[:2]]   # relu(d1), relu(d2)
        self.cols = torch.zeros(B * 65536, **f)          # fp32 tap columns (largest: d1 [B*16, 4096] = d2 [B*64, 1024])
        self.logits_cl = torch.zeros(B, 3072, **f)       # channels-last logits (y, x, c)
        self.logits = None                               # reference order, on demand (forward())
        self.bce = torch.zeros(B, **f)
        # ---- backward ----
        self.gl = torch.zeros(B, 3072, **f)
        self.glp = PB(B * 1024, 3, 2, dev)
        self.G = [PB(B * s * s, 16 * co, 2, dev) for _, _, co, s in _DEC]          # im2col of the output gradients
        self.gb = [PB(B * (2 * s) ** 2, co, 2, dev) for _, _, co, s in _DEC[:2]]   # d loss / d pre-activation of d1, d2
        self.gdd = PB(B * 16, 128, 3, dev)               # feeds the reverse sweep of the manifold chain: 3 planes
        self.gddf = PB(B, 2048, 3, dev)
        self.gz = torch.zeros(B, Sd, **f)
        self.gml = torch.zeros(B, P, **f)
        self.gmlp = PB(B, P, 3, dev)
        self.ghf = PB(B, 8192, 2, dev)
        self.ga = [PB(B * (s // 2) ** 2, co, 2, dev) for _, _, co, s in _ENC]      # d loss / d pre-activation of e0..e2
        self.fused = False
        self.has_mu_sigma = False
        self.drew_eps = False


class FusedConvolutionalVAE(FusedFeedForwardVAE):

    def __init__(self, h_dim: int, components, dataset, scalar_parametrization: bool,
                 img_dims: Tuple[int, int, int] = (3, 32, 32), device="cuda") -> None:
        """Same signature as ConvolutionalVAE (conv_vae.py:30-35) plus the device."""
        if tuple(img_dims) != (3, 32, 32) or h_dim != 8192 or dataset.in_dim != 3072:
            raise NotImplementedError("ConvolutionalVAE's layers fix img_dims = (3, 32, 32), in_dim = 3072 and "
                                      "h_dim = 512 * 4 * 4 = 8192 (conv_vae.py:47-55,65)")
        self.img_dims = tuple(img_dims)
        self.img_dims_flat = 3072
        super().__init__(h_dim, components, dataset, scalar_parametrization, device=device, input_planes=3)

    # ------------------------------------------------------------------------------------------ parameters
    def _build_layers(self) -> None:
        """conv_vae.py:47-55, in that order (same default initialisation per seed)."""
        self.e0 = nn.Conv2d(3, 64, 4, 2, 1)
        self.e1 = nn.Conv2d(64, 128, 4, 2, 1)
        self.e2 = nn.Conv2d(128, 512, 4, 2, 1)
        self.d0 = nn.Linear(self.total_z_dim, 2048)
        self.d1 = nn.ConvTranspose2d(128, 256, 4, 2, 1)
        self.d2 = nn.ConvTranspose2d(256, 64, 4, 2, 1)
        self.d3 = nn.ConvTranspose2d(64, 3, 4, 2, 1)

    def _layers(self):
        return [("e0", self.e0), ("e1", self.e1), ("e2", self.e2), ("d0", self.d0), ("d1", self.d1), ("d2", self.d2),
                ("d3", self.d3)]

    def _net_params(self):
        rest = []
        for nm, layer in self._layers():
            rest += [(nm + ".weight", layer.weight), (nm + ".bias", layer.bias)]
        return self._head_params() + rest

    def _master_perm(self, name: str):
        # conv filters [Co, Ci, ky, kx] / transposed-conv filters [Ci, Co, ky, kx] are stored with the second axis
        # innermost: rows of the GEMM operand W[Co, (ky, kx, ci)] / W[Ci, (ky, kx, co)]
        if name.endswith(".weight") and name[:2] in ("e0", "e1", "e2", "d1", "d2", "d3"):
            return (0, 2, 3, 1)
        return None

    def _bind_views(self, flat: Tensor, bucket: Tensor) -> None:
        dev, P, Sd = self.device, self.desc.ld_ml, self.desc.ld_z
        self.fused_latent, self.latent_gemm = False, True

        def mat(buf, name, rows):
            o, n = self._slices[name]
            return buf[o:o + n].view(rows, n // rows)

        o0, _ = self._slices["components.0.fc_mean.weight"]
        self.Wh, self.gWh = flat[o0:o0 + P * 8192].view(P, 8192), bucket[o0:o0 + P * 8192].view(P, 8192)
        b0, _ = self._slices["components.0.fc_mean.bias"]
        self.bh, self.gbh = flat[b0:b0 + P], bucket[b0:b0 + P]
        # GEMM views of the weights / gradients (rows of the master layout) and their operand planes.  Operands ahead
        # of a relu / the manifold maps carry 3 planes (fp32 accuracy, vae.py's precision policy); d3 feeds the BCE.
        self._W, self._gW, self._bias, self._gbias, self._Wp = {}, {}, {}, {}, {}
        rows = {"e0": 64, "e1": 128, "e2": 512, "d0": 2048, "d1": 128, "d2": 256, "d3": 64}
        for nm, r in rows.items():
            self._W[nm], self._gW[nm] = mat(flat, nm + ".weight", r), mat(bucket, nm + ".weight", r)
            o, n = self._slices[nm + ".bias"]
            self._bias[nm], self._gbias[nm] = flat[o:o + n], bucket[o:o + n]
            self._Wp[nm] = ops.PlaneBuf(r, self._W[nm].shape[1], 2 if nm == "d3" else 3, dev)
        self.Whp = ops.PlaneBuf(P, 8192, 3, dev)

    def plane_targets(self):
        t = [(self._slices["components.0.fc_mean.weight"][0], self.desc.ld_ml, self.Whp, self.Wh)]
        for nm in ("e0", "e1", "e2", "d0", "d1", "d2", "d3"):
            t.append((self._slices[nm + ".weight"][0], self._W[nm].shape[0], self._Wp[nm], self._W[nm]))
        return t

    def dp_early_begin(self) -> int:
        return self._slices["d0.weight"][0]   # the decoder's gradients are complete once d0's are

    def refresh_weight_planes(self) -> None:
        for _, _, buf, w in self.plane_targets():
            ops.split_planes(w, buf)
        self._planes_stale = False

    def _workspace(self, B: int) -> _ConvWorkspace:
        ws = self._ws.get(B)
        if ws is None:
            ws = self._ws[B] = _ConvWorkspace(self, B)
        return ws

    def _gemm(self, site, a, b, M, N, K, **kw):
        """No per-site autotuning here.  The automatic tile policy of mvae_gemm is tuned for the MLP's small GEMMs (two
        co-resident CTAs per SM); the convolutional stacks' GEMMs are large and carry up to 3 + 3 operand planes:
        128-wide tiles, one CTA per SM (a deeper pipeline fits)."""
        if N >= 128 and "tile" not in kw:
            kw["tile"] = (128, 1)
        ops.gemm(a, b, M, N, K, **kw)

    # ------------------------------------------------------------------------------------------ kernel sequences
    def _input_planes(self, ws: _ConvWorkspace, train: bool) -> None:
        if ws.u8:
            raise NotImplementedError("uint8 batches are the binarised-image path (MNIST / Omniglot); CIFAR batches are "
                                      "real-valued floats (image_reconstruction.py:116-143)")
        B = ws.B
        ops.permute_sc(ws.x, ws.x_cl.view(B * 1024, 3), B, 1024, 3, to_nhwc=True)   # (c, y, x) -> (y, x, c)
        ops.split_planes(ws.x_cl.view(B * 1024, 3), ws.xp)

    def _encoder_kernels(self, ws: _ConvWorkspace) -> None:
        B = ws.B
        src = ws.xp
        for i, (nm, ci, co, s) in enumerate(_ENC):
            ops.conv_im2col(src, B, s, s, ci, ws.A[i], ones_col=True)
            M = B * (s // 2) ** 2
            self._gemm(nm + "_fwd", ws.A[i], self._Wp[nm], M, co, 16 * ci, epilogue=L.EPI_BIAS_RELU, bias=self._bias[nm],
                       out_planes=ws.a[i])
            src = ws.a[i]
        ops.permute_sc(ws.a[2], ws.hp, B, 16, 512, to_nhwc=False)        # x.view(bs, -1): (y, x, c) -> (c, y, x)

    def _decoder_kernels(self, ws: _ConvWorkspace) -> None:
        B, Sd = ws.B, self.desc.ld_z
        ops.split_planes(ws.z, ws.zp)
        self._gemm("d0_fwd", ws.zp, self._Wp["d0"], B, 2048, Sd, epilogue=L.EPI_BIAS_RELU, bias=self._bias["d0"],
                   out_planes=ws.ddf)
        ops.permute_sc(ws.ddf, ws.dd, B, 16, 128, to_nhwc=True)          # x.view(-1, 128, 4, 4) -> channels-last
        src = ws.dd
        for i, (nm, ci, co, s) in enumerate(_DEC):
            M = B * s * s
            cols = ws.cols[:M * 16 * co].view(M, 16 * co)
            # tap columns: x[M, Ci] . W[Ci, (ky, kx, co)]  (the weight read MN-major: no transposed copy)
            self._gemm(nm + "_fwd", src, self._Wp[nm], M, 16 * co, ci, b_major=L.MN_MAJOR, out_f32=cols,
                       a_planes=2 if i == 2 else None)   # d3 feeds the (smooth) reconstruction loss: 2 x 2 planes
            if i < 2:
                ops.conv_col2im(cols, B, s, s, co, bias=self._bias[nm], act=1, out_planes=ws.b[i])
                src = ws.b[i]
            else:
                ops.conv_col2im(cols, B, s, s, co, bias=self._bias[nm], act=0, out_f32=ws.logits_cl.view(B * 1024, 3))

    def _forward_kernels(self, ws: _ConvWorkspace, beta: float, train: bool, want_mu_sigma: bool,
                         logits: Optional[Tensor], draw_eps: bool = False):
        if self._planes_stale:
            self.refresh_weight_planes()
        main, side = torch.cuda.current_stream(self.device), self._side_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ops.step_prologue(ws.eps if draw_eps else None, self.noise_seed, self._bin_ctr,
                              [ws.ml, ws.gz if train else None,
                               self._bucket[:self._n_net + self.desc.C] if train else None])
        ws.drew_eps = draw_eps
        ws.has_mu_sigma = want_mu_sigma
        ws.fused = False
        self._input_planes(ws, train)
        self._encoder_kernels(ws)
        main.wait_stream(side)
        self._heads_gemm(ws)
        ops.pm_forward(self.desc, ws.ml, ws.eps, self._rflat, want_mu_sigma=want_mu_sigma,
                       flag=ws.flag if self.check_finite else None,
                       out={"z": ws.z, "kl": ws.kl, "mu": ws.mu, "sigma": ws.sigma})
        self._decoder_kernels(ws)
        # BCE-with-logits on real-valued targets (image_reconstruction.py:142-143) + row sums (vae.py:131); logits and
        # targets are both channels-last, the sum over a row does not care
        ops.recon_loss(self.recon_kind, ws.logits_cl, ws.x_cl, out=(ws.bce, ws.gl if train else None))
        if logits is not None:
            ops.permute_sc(ws.logits_cl.view(ws.B * 1024, 3), logits, ws.B, 1024, 3, to_nhwc=False)
        if not train:
            ops.elbo_reduce(ws.bce, ws.kl, beta, out=self._stats)

    def _heads_gemm(self, ws) -> None:
        ops.gemm(ws.hp, self.Whp, ws.B, self.desc.ld_ml, 8192, bias=self.bh, out_f32=ws.ml, split_k=8192 // 64)

    def _backward_kernels(self, ws: _ConvWorkspace, beta: float, early: bool = True, advance: Optional[bool] = None):
        B, P, Sd = ws.B, self.desc.ld_ml, self.desc.ld_z
        MN = L.MN_MAJOR
        advance = early if advance is None else advance
        main, side = torch.cuda.current_stream(self.device), self._side_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):   # weight / bias gradients run beside the chain of input gradients
            if advance and ws.drew_eps:
                ops.counter_add(self._bin_ctr)
            ops.elbo_reduce(ws.bce, ws.kl, beta, out=self._stats)
            ops.colsum(ws.gl.view(B * 1024, 3), B * 1024, 3, self._gbias["d3"])
        ops.split_planes(ws.gl.view(B * 1024, 3), ws.glp)
        # ---- decoder: d3, d2, d1 (ConvTranspose2d: the output gradient is gathered like a conv's input) ----
        g_out = ws.glp
        for i in (2, 1, 0):
            nm, ci, co, s = _DEC[i]
            M = B * s * s
            x_in = ws.b[i - 1] if i > 0 else ws.dd
            ops.conv_im2col(g_out, B, 2 * s, 2 * s, co, ws.G[i])
            side.wait_stream(main)
            with torch.cuda.stream(side):
                # gW[Ci, (ky, kx, co)] = x^T G
                self._gemm(nm + "_wgrad", x_in, ws.G[i], ci, 16 * co, M, a_major=MN, b_major=MN, split_k=0,
                           out_f32=self._gW[nm], a_planes=2, b_planes=2)
            # g x = G W^T, masked by the relu that produced x
            g_in = ws.gb[i - 1] if i > 0 else ws.gdd
            self._gemm(nm + "_dgrad", ws.G[i], self._Wp[nm], M, ci, 16 * co, epilogue=L.EPI_RELU_MASK, mask=x_in,
                       out_planes=g_in, a_planes=2, b_planes=2)
            if i > 0:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    pn = _DEC[i - 1][0]
                    ops.colsum(g_in, M, ci, self._gbias[pn])   # bias gradient of the layer below (its output = x)
            g_out = g_in
        # ---- d0 ----
        ops.permute_sc(ws.gdd, ws.gddf, B, 16, 128, to_nhwc=False)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._gemm("d0_wgrad", ws.gddf, ws.zp, 2048, Sd + 1, B, a_major=MN, b_major=MN, split_k=0,
                       out_f32=self._gW["d0"], out_col=self._gbias["d0"], col_split=Sd, a_planes=2, b_planes=2)
        # (N = total_z_dim is a single narrow tile and K = 2048 is long: K slices over the SMs, summed into the zeroed gz)
        self._gemm("d0_dgrad", ws.gddf, self._Wp["d0"], B, Sd, 2048, b_major=MN, out_f32=ws.gz, split_k=0)
        early = early and self._early_step is not None
        if early:   # the decoder's parameters are final: exchange + update them under the rest of the backward pass
            comm = self._comm_stream()
            comm.wait_stream(side)
            comm.wait_stream(main)
            with torch.cuda.stream(comm):
                self._early_step()
        # ---- latent ----
        ops.pm_backward(self.desc, ws.ml, ws.eps, self._rflat, ws.gz, None, beta, gml=ws.gml, gradius=self._gradius)
        if self._any_fixed_radius:
            self._gradius.mul_(self._radius_mask)
        ops.split_planes(ws.gml, ws.gmlp)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._gemm("heads_wgrad", ws.gmlp, ws.hp, P, 8193, B, a_major=MN, b_major=MN, split_k=0, out_f32=self.gWh,
                       out_col=self.gbh, col_split=8192)
        self._gemm("heads_dgrad", ws.gmlp, self.Whp, B, 8192, P, b_major=MN, epilogue=L.EPI_RELU_MASK, mask=ws.hp,
                   out_planes=ws.ghf, a_planes=2, b_planes=2)
        ops.permute_sc(ws.ghf, ws.ga[2], B, 16, 512, to_nhwc=True)
        # ---- encoder: e2, e1, e0 (Conv2d: the input gradient is scattered back like a transposed conv's output) ----
        for i in (2, 1, 0):
            nm, ci, co, s = _ENC[i]
            M = B * (s // 2) ** 2
            side.wait_stream(main)
            with torch.cuda.stream(side):
                # gW[Co, (ky, kx, ci)] = g^T A, bias gradient from A's ones column
                self._gemm(nm + "_wgrad", ws.ga[i], ws.A[i], co, 16 * ci + 1, M, a_major=MN, b_major=MN, split_k=0,
                           out_f32=self._gW[nm], out_col=self._gbias[nm], col_split=16 * ci, b_planes=2)
            if i > 0:
                cols = ws.cols[:M * 16 * ci].view(M, 16 * ci)
                self._gemm(nm + "_dgrad", ws.ga[i], self._Wp[nm], M, 16 * ci, co, b_major=MN, out_f32=cols, b_planes=2)
                ops.conv_col2im(cols, B, s // 2, s // 2, ci, act=2, mask=ws.a[i - 1], out_planes=ws.ga[i - 1])
        main.wait_stream(side)
        if early:
            main.wait_stream(comm)

    # ------------------------------------------------------------------------------------------ reference API
    @torch.no_grad()
    def encode(self, x: Tensor) -> Tensor:
        """conv_vae.py:57-66 -> [B, 8192] in the reference's (c, y, x) order."""
        ws = self._workspace(x.shape[0])
        self._stage_x(ws, ws.slot, x)
        if self._planes_stale:
            self.refresh_weight_planes()
        self._input_planes(ws, train=False)
        self._encoder_kernels(ws)
        return ws.hp.to_float()

    @torch.no_grad()
    def decode(self, concat_z: Tensor) -> Tensor:
        """conv_vae.py:68-79 for [B, total_z_dim] or [n, B, total_z_dim] -> logits in the reference's (c, y, x) order."""
        lead = concat_z.shape[:-1]
        z2 = concat_z.reshape(-1, self.total_z_dim).float().contiguous()
        ws = self._workspace(z2.shape[0])
        if self._planes_stale:
            self.refresh_weight_planes()
        if z2.data_ptr() != ws.z.data_ptr():
            ws.z.copy_(z2)
        self._decoder_kernels(ws)
        out = torch.empty(z2.shape[0], 3072, device=self.device)
        ops.permute_sc(ws.logits_cl.view(-1, 3), out, z2.shape[0], 1024, 3, to_nhwc=False)
        return out.reshape(*lead, 3072)

    @torch.no_grad()
    def log_likelihood(self, x: Tensor, n: int = 500, eps: Optional[Tensor] = None):
        """vae.py:82-123 with the convolutional decoder: the encoder and the heads run once, then per sample the fused
        latent kernel (z, sum_c log q - log p), the decoder on the B rows and the BCE row sums; one streaming logsumexp
        at the end (the same kernels as FusedFeedForwardVAE.log_likelihood, one sample per decoder pass)."""
        B = x.shape[0]
        self._sync_radii()
        ws = self._workspace(B)
        self._stage_x(ws, ws.slot, x)
        if self._planes_stale:
            self.refresh_weight_planes()
        self._input_planes(ws, train=False)
        self._encoder_kernels(ws)
        ws.ml.zero_()
        self._heads_gemm(ws)
        f = dict(device=self.device, dtype=torch.float32)
        recon, diff = torch.zeros(n, B, **f), torch.empty(n, B, **f)
        zsum, e1 = torch.zeros(B, self.desc.ld_z, **f), torch.empty(1, B, self.desc.ld_eps, **f)
        for s in range(n):
            if eps is None:
                ops.step_prologue(e1.view(-1), self.noise_seed ^ 0x1CEB00DA, self._bin_ctr)
                ops.counter_add(self._bin_ctr)
            else:
                e1.copy_(eps[s:s + 1], non_blocking=True)
            ops.iwae_latent(self.desc, ws.ml, e1, self._rflat, ws.z.view(1, B, -1), diff[s:s + 1], zsum)
            self._decoder_kernels(ws)
            ops.recon_loss(self.recon_kind, ws.logits_cl, ws.x_cl, out=(recon[s], None))
        log_p_x, mi = ops.iwae_reduce(recon, diff)
        cov_norm = ops.iwae_cov_norm(ws.x, zsum, n)
        return log_p_x, mi, cov_norm.reshape(())
