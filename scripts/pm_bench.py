"""Times the fused product-manifold kernels over signatures and batch sizes (CUDA events, L2 flushed).
usage: python scripts/pm_bench.py [sig ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvae_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
sigs = sys.argv[1:] or ["h2,s2,e2", "h6,h6,s6,s6,e6", "h2", "p2", "e2"]
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    s = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    e = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for i in range(iters):
        flush_buf.zero_()
        s[i].record()
        fn()
        e[i].record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in zip(s, e))
    return t[len(t) // 2] * 1e3  # us (median)


for sig in sigs:
    desc = ops.make_desc(sig)
    C, Sn, Sd, P = desc.C, desc.ld_eps, desc.ld_z, desc.ld_ml
    bf, bb = 4 * (3 * Sn + Sd + C), 4 * (5 * Sn + Sd)
    for B in (4096, 8192, 16384, 1 << 18, 1 << 22):
        g = torch.Generator(device=dev).manual_seed(0)
        ml = torch.randn(B, P, device=dev, generator=g) * 0.5
        eps = torch.randn(B, Sn, device=dev, generator=g)
        R = torch.ones(C, device=dev)
        out = {"z": torch.empty(B, Sd, device=dev), "kl": torch.empty(B, C, device=dev)}
        gz = torch.randn(B, Sd, device=dev, generator=g)
        gml = torch.empty_like(ml)
        gR = torch.zeros(C, device=dev)
        tf = timeit(lambda: ops.pm_forward(desc, ml, eps, R, out=out))
        tb = timeit(lambda: ops.pm_backward(desc, ml, eps, R, gz, None, 1.0, gml=gml, gradius=gR))
        print(f"{sig:18s} B={B:8d} fwd {tf:9.1f} us {B*bf/tf/1e3:8.1f} GB/s | bwd {tb:9.1f} us {B*bb/tb/1e3:8.1f} GB/s"
              f"  tile={os.environ.get('MVAE_PM_TILE', 'auto')}", flush=True)
