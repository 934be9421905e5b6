"""Diagnostic: per-tensor deviation of the fused train step from the float64 oracle after each of k steps."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc  # noqa: E402
from mvae_b200 import components, data, vae  # noqa: E402

dev = torch.device("cuda:0")
sig, B, D, H = (sys.argv[1] if len(sys.argv) > 1 else "h6,h6,s6,s6,e6"), int(sys.argv[2]) if len(sys.argv) > 2 else 512, 784, 400
dense = len(sys.argv) > 3 and sys.argv[3] == "dense"
for graph in (False, True):
    torch.manual_seed(0)
    model = vae.FusedFeedForwardVAE(H, components.parse_components(sig, False),
                                    data.GenericDataset(B, D, "bce", binary_inputs=True), False, device=dev)
    model.use_cuda_graph = graph
    opt = vae.FusedCurvatureOptimizer(model, 1e-3, fixed_curvature=False, should_do_curvature_step=lambda: True)
    ov = orc.OracleVAE(sig, D, H, "bce", False)
    p0 = {k: v.detach().cpu().double().numpy() for k, v in model.state_dict().items()}
    tr = orc.OracleTrainer(ov, p0)
    g = torch.Generator().manual_seed(1)
    for i in range(3):
        x = (torch.rand(B, D, generator=g) < (0.5 if dense else 0.1307)).float()
        eps = torch.randn(B, model.desc.ld_eps, generator=g)
        prev = {k: v.copy() for k, v in tr.params.items()}
        bs, _ = model.train_step(opt, x.to(dev), 0.8, eps=eps.to(dev))
        ws = model._last_ws
        fwd = ov.step(tr.params, x.double().numpy(), eps.double().numpy(), beta=0.8, backward=False)
        dec = {}
        for name, act, pre in (("h", ws.h32 if ws.fused else ws.hp.to_float(), fwd["h_pre"]), ("dd", ws.ddp.to_float(), fwd["dd_pre"])):
            on = act.cpu().numpy() > 0
            diff = on != (pre > 0)
            print(f"  graph={graph} step {i} {name}: flips {int(diff.sum())} max|pre| at flips {np.abs(pre[diff]).max() if diff.any() else 0:.2e} max|pre| {np.abs(pre).max():.2f}"
                  f"  act err {np.abs(act.cpu().numpy() - np.maximum(pre, 0)).max():.2e}")
            dec[name] = on
        ref = tr.step(x.double().numpy(), eps.double().numpy(), 0.8, relu_decisions=dec)
        print(f" graph={graph} step {i}: elbo rel {abs(bs.elbo - ref['elbo']) / abs(ref['elbo']):.2e} bce rel {abs(bs.bce - ref['bce_sum']) / abs(ref['bce_sum']):.2e} "
              f"kl rel {abs(bs.kl - ref['kl_sum']) / abs(ref['kl_sum']):.2e}  kl_c rel {np.abs((np.array(bs.component_kl) - ref['kl_comp']) / ref['kl_comp']).max():.2e}")
        errs = {}
        for k, v in model.state_dict().items():
            den = np.linalg.norm(tr.params[k] - prev[k])
            if den > 0:
                errs[k] = np.linalg.norm(v.detach().cpu().double().numpy() - tr.params[k]) / den
        bad = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
        print("   worst step-movement errors:", ", ".join(f"{k} {e:.1e}" for k, e in bad))
        g_dev = model._gradius.detach().cpu().numpy()
        g_ref = np.array([float(ref["grads"].get(f"components.{j}.{nm}", 0.0)) for j, nm in
                          enumerate(["_nradius" if hasattr(c, "_nradius") else "_pradius" if hasattr(c, "_pradius") else "x" for c in model.components])])
        print("   gradius dev", np.round(g_dev, 2), "ref", np.round(g_ref, 2))
