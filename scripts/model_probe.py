"""Diagnostic: per-tensor errors of one fused train step against the float64 oracle (intermediates included)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as orc  # noqa: E402
from helpers import normwise  # noqa: E402
from mvae_b200 import components, data, vae  # noqa: E402

dev = torch.device("cuda:0")


class NoOpt:
    def zero_grad(self):
        pass

    def step(self):
        pass


for sig, B, D, H, recon in [("p2", 2048, 50, 400, "nll"), ("h2,s2,e2", 4096, 784, 400, "bce")]:
    torch.manual_seed(0)
    model = vae.FusedFeedForwardVAE(H, components.parse_components(sig, False), data.GenericDataset(B, D, recon), False,
                                    device=dev)
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(B, D, generator=g) < 0.1307).float() if recon == "bce" else torch.randn(B, D, generator=g)
    eps = torch.randn(B, model.desc.ld_eps, generator=g)
    params = {k: v.detach().cpu().double().numpy() for k, v in model.state_dict().items()}
    ref = orc.OracleVAE(sig, D, H, recon, False).step(params, x.double().numpy(), eps.double().numpy(), beta=0.8)
    bs, _ = model.train_step(NoOpt(), x, 0.8, eps=eps.to(dev))
    ws = model._last_ws
    print(sig, "elbo", bs.elbo, ref["elbo"])
    print("  h   ", normwise(ws.hp.to_float().cpu().numpy(), ref["h"]))
    print("  ml  ", normwise(ws.ml.cpu().numpy(), ref["ml"]))
    print("  z   ", normwise(ws.z.cpu().numpy(), ref["z"]))
    print("  kl  ", normwise(ws.kl.cpu().numpy(), ref["kl"]))
    print("  bce ", normwise(ws.bce.cpu().numpy(), ref["bce"]))
    print("  gz  ", normwise(ws.gz.cpu().numpy(), ref["gz"]))
    gml = ws.gml.cpu().numpy()
    print("  gml ", normwise(gml, ref["gml"]), "cols", ["%.1e" % normwise(gml[:, j], ref["gml"][:, j]) for j in range(gml.shape[1])])
    print("  gmlp", normwise(ws.gmlp.to_float().cpu().numpy(), ref["gml"]))
    e = np.abs(gml - ref["gml"])
    b = np.unravel_index(np.argmax(e), e.shape)
    print("  worst gml row", b, gml[b[0]], ref["gml"][b[0]], "ml", ref["ml"][b[0]], "eps", eps[b[0]].numpy())
    for k, p in model.named_parameters():
        if k in ref["grads"]:
            got = p.grad.detach().cpu().numpy()
            err = normwise(got, ref["grads"][k]) if got.ndim else abs(got - ref["grads"][k]) / max(1, abs(ref["grads"][k]))
            print("  grad %-36s %.2e" % (k, err))
    # exact-arithmetic cross-check of the heads wgrad from the kernel's own gml and h
    gWh = ws.gml.double().t() @ ws.hp.to_float().double()
    print("  heads wgrad vs fp64 product of the kernel's own operands:",
          normwise(model.gWh.cpu().numpy(), gWh.cpu().numpy()))
