"""Where the time of the tcgen05 GEMM launches of one train step goes: per-CTA %globaltimer stamps (mvae_debug_gemm)
of one warm launch of every GEMM call of the step, next to its back-to-back time.
usage: python scripts/gemm_phases.py [workload]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mvae_b200 import _lib as L  # noqa: E402
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
orig = ops.gemm


def rec(*a, **k):
    calls.append((a, k))
    return orig(*a, **k)


ops.gemm = rec
model.use_cuda_graph = False
model.train_step(opt, x, 1.0, sync_stats=False)
model.refresh_weight_planes()
torch.cuda.synchronize()
ops.gemm = orig

PH = ["set-up (barriers, TMEM)", "wait for predecessor", "first stage lands", "main loop (MMA issue)",
      "MMAs retire", "epilogue (first warp)", "stores, last warp, exit"]
lib = L.lib()
st = torch.zeros(8192, 16, dtype=torch.int64, device=dev)
FINE = [(8, "first accumulator chunk in registers"), (9, "first chunk: loss arithmetic done"),
        (10, "first chunk complete (planes staged)"), (11, "all epilogue warps at the barrier"),
        (12, "bulk stores issued and read")]


def timed(a, k):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            orig(*a, **k)
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / 100


for a, k in calls:
    us = timed(a, k)
    variants = []
    for flags, what in ((1, "no plane staging / stores"), (2, "no loss arithmetic"), (4, "no operand prefetch"),
                        (7, "none of the three")):
        lib.mvae_debug_gemm(None, flags)
        variants.append(f"{what}: {timed(a, k):.2f}")
    lib.mvae_debug_gemm(None, 0)
    st.zero_()
    torch.cuda.synchronize()
    lib.mvae_debug_gemm(ctypes.c_void_p(st.data_ptr()), 0)
    orig(*a, **k)
    torch.cuda.synchronize()
    lib.mvae_debug_gemm(None, 0)
    t = st.cpu().numpy().astype(np.float64)
    t = t[t[:, 7] > 0]
    t0 = t[:, 0].min()
    print(f"== gemm M={a[2]} N={a[3]} K={a[4]} epi={k.get('epilogue', 0)} a_pl={a[0].planes} "
          f"b_pl={k.get('b_planes') or a[1].planes} tile={k.get('tile')} "
          f"split_k={k.get('split_k')}: {len(t)} CTAs, back-to-back {us:.2f} us; "
          f"CTA starts spread {(t[:, 0].max() - t0) / 1e3:.2f} us, last CTA ends at {(t[:, 7].max() - t0) / 1e3:.2f} us")
    for i, ph in enumerate(PH):
        d = (t[:, i + 1] - t[:, i]) / 1e3
        print(f"     {ph:26s} median {np.median(d):6.2f}  p90 {np.percentile(d, 90):6.2f}  max {d.max():6.2f} us"
              f"   (ends at median {np.median(t[:, i + 1] - t0) / 1e3:6.2f} us)")
    for i, what in FINE:   # SM cycle counter of the first epilogue thread, relative to "accumulator complete" (slot 15)
        if (t[:, i] > 0).all():
            d = (t[:, i] - t[:, 15]) / 1.965e3
            print(f"     {what:40s} {np.median(d):6.2f} us after the accumulator was complete (p90 {np.percentile(d, 90):6.2f})")
    print("     back-to-back us with parts of the epilogue removed — " + "; ".join(variants))
