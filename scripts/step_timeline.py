"""Timeline of ONE replay of the captured train step: start / inputs-ready / end of every stamped kernel (the five GEMMs and
the two latent kernels) on the device's %globaltimer (mvae_debug_timeline), L2 flushed before the step like bench.py.
The kernels without stamps (plane split, ELBO, optimizer, prologue) show up as the gaps between them.
usage: python scripts/step_timeline.py [workload]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mvae_b200 import _lib as L  # noqa: E402
from mvae_b200 import components, data, vae  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
sig, B, D, H, recon, fixed, desc = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = vae.FusedFeedForwardVAE(H, components.parse_components(sig, fixed),
                                data.GenericDataset(B, D, recon, binary_inputs=(recon == "bce")), False, device=dev)
model.use_cuda_graph = True
model.adopt_device_inputs = True
with torch.no_grad():
    for rp in model._radius_params:
        if rp is not None and rp.requires_grad:
            rp.fill_(10.0)
opt = vae.FusedCurvatureOptimizer(model, 1e-3, fixed_curvature=fixed, should_do_curvature_step=lambda: True)
x = bench.synthetic_x(recon, B, D, 0).to(dev)
for _ in range(5):
    model.train_step(opt, x, 1.0, sync_stats=False)
torch.cuda.synchronize()

lib = L.lib()
buf = torch.zeros(1 << 20, dtype=torch.int64, device=dev)
lib.mvae_debug_timeline(ctypes.c_void_p(buf.data_ptr()))
model._graphs.clear()                      # the next step warms up eagerly, then captures: both are logged
model.train_step(opt, x, 1.0, sync_stats=False)
torch.cuda.synchronize()
log = (ctypes.c_int32 * (5 * 256))()
n = lib.mvae_debug_timeline_log(log, 256)
lib.mvae_debug_timeline(None)
entries = [tuple(log[5 * i + j] for j in range(5)) for i in range(n)]
offs = np.concatenate([[0], np.cumsum([e[1] * 16 for e in entries])])
half = n // 2                              # eager warm-up launches first, the captured ones second

flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for rep in range(3):
    flush_buf.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    model.train_step(opt, x, 1.0, sync_stats=False)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
t = buf.cpu().numpy().astype(np.float64)
rows = []
for i in range(half, n):
    kind, ncta, a, b, c = entries[i]
    r = t[offs[i]:offs[i + 1]].reshape(ncta, 16)
    if kind == 0:
        name, ready, end = f"gemm M={a} N={b} K={c}", r[:, 2], r[:, 7]
        extra = (f"first stage +{np.median(r[:, 3] - r[:, 2]) / 1e3:.2f}, main loop {np.median(r[:, 4] - r[:, 3]) / 1e3:.2f}, "
                 f"epilogue+exit {np.median(r[:, 7] - r[:, 4]) / 1e3:.2f}")
    else:
        last = 7 if kind == 2 else 5
        name, ready, end = ("latent_forward" if kind == 1 else "latent_backward"), r[:, 1], r[:, last]
        ph = ["loads", "heads", "chain", "fc_d0+stores"] if kind == 1 else ["loads", "gz", "wgrad d0", "sweep", "heads", "gh store"]
        extra = ", ".join(f"{nm} {np.median(r[:, j + 2] - r[:, j + 1]) / 1e3:.2f}" for j, nm in enumerate(ph))
    rows.append((r[:, 0].min(), name, ncta, np.median(r[:, 0]), np.median(ready), ready.max(), np.median(end), end.max(), extra))
rows.sort()
t0 = rows[0][0]
print(f"{wl}: step {ms * 1e3:.1f} us (CUDA events, L2 flushed before); times in us from the first stamped kernel's start")
print(f"{'kernel':32s} {'CTAs':>5s} {'start':>7s} {'inputs ready (median/max)':>26s} {'end (median/max)':>18s}")
prev_end = None
for st, name, ncta, st_med, rd_med, rd_max, en_med, en_max, extra in rows:
    gap = "" if prev_end is None else f"   [{(rd_med - prev_end) / 1e3:+.2f} us after the previous end]"
    print(f"{name:32s} {ncta:5d} {(st - t0) / 1e3:7.2f} {(rd_med - t0) / 1e3:12.2f} /{(rd_max - t0) / 1e3:7.2f}      "
          f"{(en_med - t0) / 1e3:8.2f} /{(en_max - t0) / 1e3:7.2f}   {extra}{gap}")
    prev_end = en_max
