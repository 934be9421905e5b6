"""The CPU arm of bench.py (`--impl reference`, `cpu_baseline`) — TEST / MEASUREMENT INFRASTRUCTURE ONLY.

One full train step of the reference's algorithm (ModelVAE.train_step, vae.py:149-166, with the optimizer of
Trainer.build_optimizer, train.py:327-360) on the host cores, float32, all threads:

  dense layers, relu, BCE-with-logits   the same multi-threaded ATen CPU kernels the reference's own path dispatches
                                        (ffnn_vae.py:42-60, image_reconstruction.py:81-82): addmm / relu / sigmoid
  latent chain + its reverse sweep      the C / OpenMP oracle (oracle/mvae_oracle_impl.h), one fused pass per direction
                                        instead of the reference's ~4000 tiny ops
  Adam + radii SGD                      in place

The reference itself is pure Python and absent from the GPU box; in the build container this step is FASTER than the
unmodified reference on the same cores (scripts/time_reference_cpu.py), so ratios against it do not
flatter the GPU path.  tests/test_oracle_golden.py::test_cpu_baseline_step_matches_oracle pins it to OracleVAE.step."""
import numpy as np
import torch
import torch.nn.functional as F

import oracle as orc


class CpuTrainStep:

    def __init__(self, sig, in_dim, h_dim, recon_kind, params, lr=1e-3, curvature_lr=1e-4, betas=(0.9, 0.999),
                 eps=1e-8):
        self.ov = orc.OracleVAE(sig, in_dim, h_dim, recon_kind, False)
        self.desc, self.C, self.recon_kind = self.ov.desc, self.ov.C, recon_kind
        Wh, bh = self.ov.heads_matrix(params)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32).copy())  # noqa: E731
        self.p = {"Wh": t(Wh), "bh": t(bh)}
        for nm in ("fc_e0", "fc_d0", "fc_logits"):
            self.p[nm + ".W"] = t(params[nm + ".weight"])
            self.p[nm + ".b"] = t(params[nm + ".bias"])
        self.R = self.ov.radii(params, np.float32).copy()
        self.learn_R = np.array([any(f"components.{i}.{nm}" in params for nm in ("_nradius", "_pradius", "_curvature"))
                                 for i in range(self.C)])
        self.m = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.lr, self.curvature_lr, self.betas, self.eps, self.t = lr, curvature_lr, betas, eps, 0

    @torch.no_grad()
    def step(self, x: torch.Tensor, eps: torch.Tensor, beta: float = 1.0, update: bool = True) -> dict:
        p = self.p
        h = torch.relu_(torch.addmm(p["fc_e0.b"], x, p["fc_e0.W"].t()))
        ml = torch.addmm(p["bh"], h, p["Wh"].t())
        f = orc.pm_forward(self.desc, ml.numpy(), eps.numpy(), self.R, want=("z", "kl"))
        z = torch.from_numpy(f["z"])
        dd = torch.relu_(torch.addmm(p["fc_d0.b"], z, p["fc_d0.W"].t()))
        logits = torch.addmm(p["fc_logits.b"], dd, p["fc_logits.W"].t())
        if self.recon_kind == "bce":
            bce = F.binary_cross_entropy_with_logits(logits, x, reduction="none").sum(-1)
            gl = torch.sigmoid_(logits).sub_(x)
        else:
            gl = logits.sub_(x)
            bce = (gl * gl).sum(-1).mul_(0.5).add_(0.9189385332046727 * x.shape[1])
        stats = orc.elbo(bce.numpy(), f["kl"], beta)
        g = {"fc_logits.W": gl.t() @ dd, "fc_logits.b": gl.sum(0)}
        gdd = (gl @ p["fc_logits.W"]).mul_(dd > 0)
        g["fc_d0.W"], g["fc_d0.b"] = gdd.t() @ z, gdd.sum(0)
        gz = gdd @ p["fc_d0.W"]
        gml, gR = orc.pm_backward(self.desc, ml.numpy(), eps.numpy(), self.R, gz.numpy(), None, beta)
        gml = torch.from_numpy(gml)
        g["Wh"], g["bh"] = gml.t() @ h, gml.sum(0)
        gh = (gml @ p["Wh"]).mul_(h > 0)
        g["fc_e0.W"], g["fc_e0.b"] = gh.t() @ x, gh.sum(0)
        if update:
            self.t += 1
            b1, b2 = self.betas
            bc1, bc2 = 1.0 - b1 ** self.t, 1.0 - b2 ** self.t
            for k, gk in g.items():
                self.m[k].mul_(b1).add_(gk, alpha=1.0 - b1)
                self.v[k].mul_(b2).addcmul_(gk, gk, value=1.0 - b2)
                p[k].addcdiv_(self.m[k], self.v[k].sqrt().div_(bc2 ** 0.5).add_(self.eps), value=-self.lr / bc1)
            if self.curvature_lr:
                self.R = (self.R - self.curvature_lr * np.where(self.learn_R, gR, 0.0)).astype(np.float32)
        return {"elbo": stats[2], "bce_sum": stats[0], "kl_sum": stats[1], "grads": g, "gR": gR}


class CpuConvTrainStep:
    """The same for ConvolutionalVAE (conv_vae.py:28-79; BASELINE cfg5): the convolutions, transposed convolutions
    and their gradients run on the multi-threaded ATen / oneDNN CPU kernels the reference's own path dispatches
    (F.conv2d / F.conv_transpose2d under autograd), the latent chain on the C / OpenMP oracle, Adam + radii SGD in place."""

    def __init__(self, sig, params, lr=1e-3, curvature_lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
        self.ov = orc.OracleConvVAE(sig)
        self.desc, self.C = self.ov.desc, self.ov.C
        Wh, bh = self.ov.heads_matrix(params)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32).copy()).requires_grad_(True)  # noqa: E731
        self.p = {"Wh": t(Wh), "bh": t(bh)}
        for nm in ("e0", "e1", "e2", "d0", "d1", "d2", "d3"):
            self.p[nm + ".W"] = t(params[nm + ".weight"])
            self.p[nm + ".b"] = t(params[nm + ".bias"])
        self.R = self.ov.radii(params, np.float32).copy()
        self.learn_R = np.array([any(f"components.{i}.{nm}" in params for nm in ("_nradius", "_pradius", "_curvature"))
                                 for i in range(self.C)])
        self.m = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.lr, self.curvature_lr, self.betas, self.eps, self.t = lr, curvature_lr, betas, eps, 0

    def step(self, x: torch.Tensor, eps: torch.Tensor, beta: float = 1.0, update: bool = True) -> dict:
        p, B = self.p, x.shape[0]
        for v in p.values():
            v.grad = None
        a = x.view(B, 3, 32, 32)
        for nm in ("e0", "e1", "e2"):                                        # conv_vae.py:62-64
            a = torch.relu(F.conv2d(a, p[nm + ".W"], p[nm + ".b"], stride=2, padding=1))
        h = a.reshape(B, -1)
        ml = torch.addmm(p["bh"], h, p["Wh"].t())
        f = orc.pm_forward(self.desc, ml.detach().numpy(), eps.numpy(), self.R, want=("z", "kl"))
        z = torch.from_numpy(f["z"]).requires_grad_(True)
        d = torch.relu(torch.addmm(p["d0.b"], z, p["d0.W"].t())).view(B, 128, 4, 4)     # :72-73
        d = torch.relu(F.conv_transpose2d(d, p["d1.W"], p["d1.b"], stride=2, padding=1))
        d = torch.relu(F.conv_transpose2d(d, p["d2.W"], p["d2.b"], stride=2, padding=1))
        logits = F.conv_transpose2d(d, p["d3.W"], p["d3.b"], stride=2, padding=1).reshape(B, -1)
        bce = F.binary_cross_entropy_with_logits(logits, x, reduction="none").sum(-1)  # image_reconstruction.py:142-143
        stats = orc.elbo(bce.detach().numpy(), f["kl"], beta)
        bce.sum().backward()
        gml, gR = orc.pm_backward(self.desc, ml.detach().numpy(), eps.numpy(), self.R, z.grad.numpy(), None, beta)
        ml.backward(torch.from_numpy(gml))
        if update:
            with torch.no_grad():
                self.t += 1
                b1, b2 = self.betas
                bc1, bc2 = 1.0 - b1 ** self.t, 1.0 - b2 ** self.t
                for k, w in p.items():
                    gk = w.grad
                    self.m[k].mul_(b1).add_(gk, alpha=1.0 - b1)
                    self.v[k].mul_(b2).addcmul_(gk, gk, value=1.0 - b2)
                    w.addcdiv_(self.m[k], self.v[k].sqrt().div_(bc2 ** 0.5).add_(self.eps), value=-self.lr / bc1)
                if self.curvature_lr:
                    self.R = (self.R - self.curvature_lr * np.where(self.learn_R, gR, 0.0)).astype(np.float32)
        return {"elbo": stats[2], "bce_sum": stats[0], "kl_sum": stats[1], "grads": {k: w.grad for k, w in p.items()},
                "gR": gR}
