"""One (nb, nst) point of the product-manifold kernels' tile sweep: MVAE_PM_TUNE=nb,nst python scripts/pm_tune_sweep.py [sig]
(the library reads MVAE_PM_TUNE once per process).  Prints the achieved fraction of the measured HBM bandwidth at 2^22 samples."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvae_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
sig = sys.argv[1] if len(sys.argv) > 1 else "h2,s2,e2"
desc = ops.make_desc(sig)
C, Sn, Sd, P = desc.C, desc.ld_eps, desc.ld_z, desc.ld_ml
bf, bb = 4 * (3 * Sn + Sd + C), 4 * (5 * Sn + Sd)
B = 1 << 22
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
ml = torch.randn(B, P, device=dev, generator=g) * 0.5
eps = torch.randn(B, Sn, device=dev, generator=g)
R = torch.ones(C, device=dev)
out = {"z": torch.empty(B, Sd, device=dev), "kl": torch.empty(B, C, device=dev)}
gz = torch.randn(B, Sd, device=dev, generator=g)
gml = torch.empty_like(ml)
gR = torch.zeros(C, device=dev)


def timeit(fn, iters=8):
    for _ in range(2):
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
    return t[len(t) // 2] * 1e3


tf = timeit(lambda: ops.pm_forward(desc, ml, eps, R, out=out))
tb = timeit(lambda: ops.pm_backward(desc, ml, eps, R, gz, None, 1.0, gml=gml, gradius=gR))
print(f"tune={os.environ.get('MVAE_PM_TUNE', 'auto'):6s} {sig} fwd {tf:7.1f} us {B * bf / tf / 1e3 / 6553.9:.3f} | bwd {tb:7.1f} us {B * bb / tb / 1e3 / 6553.9:.3f}", flush=True)
