"""Runs a few eager train steps of cfg2 and marks the LAST one with cudaProfilerStart/Stop, so that
`ncu --profile-from-start off` captures exactly the kernels of one step with the tuned GEMM tiles.
usage: prof_step.py [workload]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mvae_b200 import components, data, vae  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
sig, B, D, H, recon, fixed, desc = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = vae.FusedFeedForwardVAE(H, components.parse_components(sig, fixed),
                                data.GenericDataset(B, D, recon, binary_inputs=(recon == "bce")), False, device=dev)
model.autotune_gemm = os.environ.get("MVAE_GEMM_AUTOTUNE", "1") != "0"
opt = vae.FusedCurvatureOptimizer(model, 1e-3, fixed_curvature=fixed, should_do_curvature_step=lambda: True)
x = bench.synthetic_x(recon, B, D, 0).to(dev)
for _ in range(4):
    model.train_step(opt, x, 1.0, sync_stats=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model.train_step(opt, x, 1.0, sync_stats=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
