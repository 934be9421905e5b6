"""Where the HOST time of a step goes (the GPU step is ~0.19 ms; the host must enqueue faster than that):
times, over 300 iterations each and without synchronising inside, the pieces of bench.py's timed loop."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mvae_b200 import components, data, ops, vae  # noqa: E402

dev = torch.device("cuda:0")
sig, B, D, H, recon, fixed, desc = bench.WORKLOADS["cfg2"]
torch.manual_seed(0)
model = vae.FusedFeedForwardVAE(H, components.parse_components(sig, fixed), data.GenericDataset(B, D, recon, binary_inputs=True),
                                False, device=dev)
model.use_cuda_graph = True
model.adopt_device_inputs = len(sys.argv) < 2 or sys.argv[1] != "copy"
opt = vae.FusedCurvatureOptimizer(model, 1e-3, fixed_curvature=fixed, should_do_curvature_step=lambda: True)
xs = [bench.synthetic_x(recon, B, D, i).to(dev) for i in range(4)]
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(8):
    model.train_step(opt, xs[i % 4], 1.0, sync_stats=False)
torch.cuda.synchronize()
N = 300


def timeit(name, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(N):
        fn(i)
    dt = (time.perf_counter() - t0) / N * 1e6
    torch.cuda.synchronize()
    print(f"{name:44s} {dt:8.1f} us per call (host, enqueue only)")


ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * N)]
timeit("flush_buf.zero_()", lambda i: flush_buf.zero_())
timeit("2 x event.record()", lambda i: (ev[2 * i].record(), ev[2 * i + 1].record()))
timeit("model.train_step(sync_stats=False)", lambda i: model.train_step(opt, xs[i % 4], 1.0, sync_stats=False))
ws = model._workspace(B)
entry = next(iter(model._graphs.values()))
timeit("CUDAGraph.replay() alone", lambda i: entry[0].replay())
timeit("model._sync_radii()", lambda i: model._sync_radii())
timeit("model._stage(ws, x, None)", lambda i: model._stage(ws, xs[i % 4], None))
timeit("ops.counter_add (one tiny kernel via ctypes)", lambda i: ops.counter_add(model._bin_ctr))
x_host = bench.synthetic_pixels(B, D, 0).pin_memory()
timeit("train_epoch, 1 uint8 batch per call", lambda i: model.train_epoch(opt, [x_host], 1.0))
batches = [x_host] * 50
torch.cuda.synchronize()
t0 = time.perf_counter()
model.train_epoch(opt, batches, 1.0)
print(f"{'train_epoch, per step of a 50-batch epoch':44s} {(time.perf_counter() - t0) / 50 * 1e6:8.1f} us (incl. the final drain)")
