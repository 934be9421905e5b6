"""Time FusedFeedForwardVAE.log_likelihood (IWAE, vae.py:82-123) at a BASELINE workload with the reference's n = 500
samples per row, and (optionally) the CPU oracle port on a bounded sample beside it.
usage: python scripts/loglik_bench.py [workload] [n] [--cpu]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mvae_b200 import components, data, ops, vae  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
wl = args[0] if args else "cfg2"
n = int(args[1]) if len(args) > 1 else 500
sig, B, D, H, recon, fixed, desc = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = vae.FusedFeedForwardVAE(H, components.parse_components(sig, fixed),
                                data.GenericDataset(B, D, recon, binary_inputs=(recon == "bce")), False, device=dev)
x = bench.synthetic_x(recon, B, D, 0).to(dev)
for _ in range(2):
    ll, mi, cov = model.log_likelihood(x, n=n)
torch.cuda.synchronize()
n0 = ops.launch_count()
reps = 5
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(reps):
    ll, mi, cov = model.log_likelihood(x, n=n)
e.record()
e.synchronize()
ms = s.elapsed_time(e) / reps
rows = n * B
flops = 2.0 * rows * (model.desc.ld_z * H + H * D)
out = {"workload": desc, "n_samples": n, "batch": B, "ms_per_batch": ms, "sample_rows_per_s": rows / ms * 1e3,
       "decoder_tflops_fp32_equiv": flops / ms / 1e9, "launches_per_batch": (ops.launch_count() - n0) // reps,
       "log_likelihood_per_row": float(ll.mean()), "mi_per_row": float(mi.mean()), "cov_norm": float(cov)}
if "--cpu" in sys.argv:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import numpy as np
    import oracle as orc
    nb, nn = min(B, 512), min(n, 8)
    params = {k: v.detach().cpu().numpy().astype(np.float32) for k, v in model.state_dict().items()}
    o = orc.OracleVAE(sig, D, H, recon, False)
    xe = x[:nb].cpu().numpy()
    eps = np.random.default_rng(0).standard_normal((nn, nb, model.desc.ld_eps)).astype(np.float32)
    o.log_likelihood(params, xe, eps[:1])
    t0 = time.perf_counter()
    o.log_likelihood(params, xe, eps)
    dt = time.perf_counter() - t0
    out["cpu_port"] = {"sample_rows_per_s": nn * nb / dt, "cores": os.cpu_count(), "sample": f"{nn} samples x {nb} rows, float32"}
print(json.dumps(out))
