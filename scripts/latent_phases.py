"""Where the time of the fused latent kernels goes: per-CTA %globaltimer stamps at the phase boundaries
(mvae_debug_latent) of one warm launch of latent_forward / latent_backward with the arguments of a real train step,
plus the kernel time with and without the weight-gradient reductions.
usage: python scripts/latent_phases.py [workload]"""
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

calls = {}
orig = {n: getattr(ops, n) for n in ("latent_forward", "latent_backward")}
for n, f in orig.items():
    def inner(*a, _n=n, _f=f, **k):
        calls[_n] = (_f, a, k)
        return _f(*a, **k)
    setattr(ops, n, inner)
model.use_cuda_graph = False
model.train_step(opt, x, 1.0, sync_stats=False)
torch.cuda.synchronize()
for n, f in orig.items():
    setattr(ops, n, f)

PH = {"latent_forward": ["pdl_wait", "load h", "heads", "manifold chain", "fc_d0 + stores"],
      "latent_backward": ["pdl_wait", "load gdd,h", "gz", "fc_d0 wgrad (+red)", "reverse sweep", "heads gWh (+red)",
                          "gh planes store"]}


FINE = {"latent_forward": [(8, "heads: partial dot products done (warp 0)"), (9, "heads: reduced + stored (warp 0)"),
                           (10, "heads: all warps at the combine barrier"), (11, "fc_d0: first column pair accumulated")],
        "latent_backward": [(8, "gz: partial dot products done (warp 0)"), (9, "heads: h columns in registers"),
                            (10, "heads: first pass of four head rows done")]}


def timed(fn, a, k, rep=20):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(rep):
            fn(*a, **k)
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / (5 * rep)


def single(fn, a, k):
    """one launch alone, warm L2, CUDA events (median of 20)"""
    ts = []
    for _ in range(20):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        fn(*a, **k)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    return float(np.median(ts))


lib = L.lib()
for name, (fn, a, k) in calls.items():
    rows = 16 if name == "latent_backward" else 8
    grid = (B + rows - 1) // rows
    print(f"== {name} ({wl}: B={B}, grid {grid}) back-to-back {timed(fn, a, k):.2f} us, alone {single(fn, a, k):.2f} us")
    if name == "latent_backward":
        lib.mvae_debug_latent(None, 1)
        print(f"   without reductions: back-to-back {timed(fn, a, k):.2f} us, alone {single(fn, a, k):.2f} us")
        lib.mvae_debug_latent(None, 0)
    st = torch.zeros(grid, 16, dtype=torch.int64, device=dev)
    for flags in ((0, 1) if name == "latent_backward" else (0,)):
        st.zero_()
        torch.cuda.synchronize()
        lib.mvae_debug_latent(ctypes.c_void_p(st.data_ptr()), flags)
        fn(*a, **k)
        torch.cuda.synchronize()
        lib.mvae_debug_latent(None, 0)
        t = st.cpu().numpy().astype(np.float64)
        nph = len(PH[name])
        t0 = t[:, 0].min()
        print(f"   stamps (flags={flags}): CTA starts spread {(t[:, 0].max() - t0) / 1e3:.2f} us, "
              f"last CTA ends at {(t[:, nph].max() - t0) / 1e3:.2f} us")
        for i, ph in enumerate(PH[name]):
            d = (t[:, i + 1] - t[:, i]) / 1e3
            print(f"     {ph:22s} median {np.median(d):6.2f}  p90 {np.percentile(d, 90):6.2f}  max {d.max():6.2f} us"
                  f"   (phase ends at median {np.median(t[:, i + 1] - t0) / 1e3:6.2f} us)")
        for i, what in FINE[name]:   # SM cycle counter, relative to the end of the load phase (slot 15)
            if (t[:, i] > 0).all():
                d = (t[:, i] - t[:, 15]) / 1.965e3
                print(f"     {what:46s} {np.median(d):6.2f} us after the loads landed (p90 {np.percentile(d, 90):6.2f})")
