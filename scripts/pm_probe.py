"""Diagnostic (not a test): error table of the fused product-manifold kernels against the float64 oracle."""
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as orc  # noqa: E402
from helpers import normwise  # noqa: E402
from mvae_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
cases = [("h2,s2,e2", 4096, [1, 1, 0], 1.0), ("h2,s2,e2", 4096, [10, 10, 0], 1.0), ("p2", 16384, [1.0], 0.8),
         ("h3,s5,p4,e7,h8,s1", 1000, [1.0, 2.0, 1.5, 0.0, 0.5, 1.0], 0.5),
         ("h6,h6,s6,s6,e6", 8229, [1.3, 0.9, 1.1, 2.0, 0.0], 0.7), ("h2,s2,p2", 4096, [1, 1, 1], 2.0),
         ("h40,s40,p40", 512, [1.0, 1.0, 3.0], 0.25), ("h2,s2,p2", 4096, [0.3, 0.3, 0.3], 0.3)]
for sig, B, Rs, sm in cases:
    d, od = ops.make_desc(sig), orc.make_desc(sig)
    rng = np.random.default_rng(zlib.crc32(sig.encode()) + B)
    ml = rng.standard_normal((B, d.ld_ml)) * sm
    eps = rng.standard_normal((B, d.ld_eps))
    R = np.asarray([r if r else 1.0 for r in Rs], dtype=np.float64)
    gz = rng.standard_normal((B, d.ld_z))
    gkl = rng.standard_normal((B, d.C))
    ref = orc.pm_forward(od, ml, eps, R, want=("z", "kl", "mu", "sigma"))
    f = np.float32
    r32 = orc.pm_forward(od, ml.astype(f), eps.astype(f), R.astype(f), want=("z", "kl", "mu", "sigma"))
    out = ops.pm_forward(d, t(ml), t(eps), t(R), want_mu_sigma=True)
    line = {k: "%.1e/%.1e" % (normwise(out[k].cpu().numpy(), ref[k]), normwise(r32[k], ref[k])) for k in ref}
    rg, rR = orc.pm_backward(od, ml, eps, R, gz, gkl)
    o32, oR32 = orc.pm_backward(od, ml.astype(f), eps.astype(f), R.astype(f), gz.astype(f), gkl.astype(f))
    g, gR = ops.pm_backward(d, t(ml), t(eps), t(R), t(gz), t(gkl))
    line["gml"] = "%.1e/%.1e" % (normwise(g.cpu().numpy(), rg), normwise(o32, rg))
    line["gR"] = "%.1e/%.1e" % (normwise(gR.cpu().numpy(), rR), normwise(oR32, rR))
    print(sig, B, Rs, "cuda/oracle32 vs f64:", line, flush=True)
