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
    # ---- fp32-inherent floor: the oracle itself in float32 vs float64, and relu decision flips ----
    p32 = {k: v.astype(np.float32) for k, v in params.items()}
    r32 = orc.OracleVAE(sig, D, H, recon, False).step(p32, x.float().numpy(), eps.float().numpy(), beta=np.float32(0.8))
    print("  oracle32 vs oracle64: gz %.2e gml %.2e" % (normwise(r32["gz"], ref["gz"]), normwise(r32["gml"], ref["gml"])))
    for k in ref["grads"]:
        a, b = np.asarray(r32["grads"][k], dtype=np.float64), np.asarray(ref["grads"][k])
        if a.ndim:
            print("  oracle32 grad %-36s max %.2e fro %.2e" % (k, normwise(a, b), np.linalg.norm(a - b) / np.linalg.norm(b)))
    dd64 = np.maximum(ref["z"] @ params["fc_d0.weight"].T + params["fc_d0.bias"], 0)
    dd_k = ws.ddp.to_float().cpu().numpy()
    flips_k = np.argwhere((dd_k > 0) != (dd64 > 0))
    dd32 = np.maximum(r32["z"] @ p32["fc_d0.weight"].T + p32["fc_d0.bias"], 0)
    flips_o = np.argwhere((dd32 > 0) != (dd64 > 0))
    h_k = ws.hp.to_float().cpu().numpy()
    print("  relu flips vs f64: fc_d0 kernel %d oracle32 %d | fc_e0 kernel %d oracle32 %d" % (
        len(flips_k), len(flips_o), int(((h_k > 0) != (ref["h"] > 0)).sum()), int(((r32["h"] > 0) != (ref["h"] > 0)).sum())))
    egz = np.abs(ws.gz.cpu().numpy() - ref["gz"]).max(1)
    top = np.argsort(-egz)[:5]
    print("  worst gz rows", top, egz[top], "flip rows kernel", sorted(set(flips_k[:, 0].tolist()))[:10])
    egml = np.abs(gml - ref["gml"])
    for j in range(min(gml.shape[1], 4)):
        t5 = np.argsort(-egml[:, j])[:4]
        print("  gml col", j, "worst rows", t5, egml[t5, j], "ref", ref["gml"][t5, j])
    # wgrad of heads from kernel's gml & ORACLE h, and oracle gml & kernel h: which operand carries the error
    hk = h_k.astype(np.float64)
    for nm, A, Bm in (("kernel gml x oracle h", gml.astype(np.float64), ref["h"]), ("oracle gml x kernel h", ref["gml"], hk)):
        W = A.T @ Bm
        Wr = ref["gml"].T @ ref["h"]
        print("   ", nm, "rows fro:", ["%.1e" % (np.linalg.norm(W[i] - Wr[i]) / np.linalg.norm(Wr[i])) for i in range(min(4, W.shape[0]))])
