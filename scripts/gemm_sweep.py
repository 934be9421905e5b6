"""GEMM-only variant of step_breakdown.py (tile-policy sweeps through MVAE_GEMM_TUNE): per-kernel time of one train step with a warm L2 (as inside the step's CUDA graph): every op of the step is recorded
once, then captured alone in a CUDA graph (REP back-to-back launches) and timed with CUDA events.
usage: python scripts/step_breakdown.py [workload]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mvae_b200 import components, data, ops, vae  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
sig, B, D, H, recon, fixed, desc = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = vae.FusedFeedForwardVAE(H, components.parse_components(sig, fixed),
                                data.GenericDataset(B, D, recon, binary_inputs=(recon == "bce")), False, device=dev)
opt = vae.FusedCurvatureOptimizer(model, 1e-3, fixed_curvature=fixed, should_do_curvature_step=lambda: True)
x = bench.synthetic_x(recon, B, D, 0).to(dev)
for _ in range(3):
    model.train_step(opt, x, 1.0, sync_stats=False)
torch.cuda.synchronize()

calls = []
NAMES = ["gemm"]
orig = {n: getattr(ops, n) for n in NAMES if hasattr(ops, n)}


def wrap(name, fn):
    def inner(*a, **k):
        calls.append((name, fn, a, k))
        return fn(*a, **k)
    return inner


for n, f in orig.items():
    setattr(ops, n, wrap(n, f))
model.use_cuda_graph = False
model.train_step(opt, x, 1.0, sync_stats=False)
model.refresh_weight_planes()
torch.cuda.synchronize()
for n, f in orig.items():
    setattr(ops, n, f)

REP = 20
total = 0.0
rows = []
for i, (name, fn, a, k) in enumerate(calls):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        fn(*a, **k)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            fn(*a, **k)
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / (5 * REP)
    shape = ""
    if name == "gemm":
        shape = f"M={a[2]} N={a[3]} K={a[4]} epi={k.get('epilogue', 0)} a_pl={a[0].planes} b_pl={k.get('b_planes') or a[1].planes}"
    elif name.startswith("skinny"):
        shape = f"K={k.get('K')} N={k.get('N')}" if "K" in k else ""
    rows.append((name, shape, us))
    total += us
for name, shape, us in rows:
    print(f"{os.environ.get('MVAE_GEMM_TUNE', 'auto'):10s} {us:8.2f} us  {shape}")
print(f"{os.environ.get('MVAE_GEMM_TUNE', 'auto'):10s} sum {total:.1f} us")
