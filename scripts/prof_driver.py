"""Launches the kernels of interest a few times (for ncu captures).  usage: prof_driver.py pm|gemm"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvae_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "pm"
g = torch.Generator(device=dev).manual_seed(0)
if what == "pm":
    sig = sys.argv[2] if len(sys.argv) > 2 else "h2,s2,e2"
    desc = ops.make_desc(sig)
    B = 1 << 22
    ml = torch.randn(B, desc.ld_ml, device=dev, generator=g) * 0.5
    eps = torch.randn(B, desc.ld_eps, device=dev, generator=g)
    R = torch.ones(desc.C, device=dev)
    out = {"z": torch.empty(B, desc.ld_z, device=dev), "kl": torch.empty(B, desc.C, device=dev)}
    gz = torch.randn(B, desc.ld_z, device=dev, generator=g)
    gml = torch.empty_like(ml)
    gR = torch.zeros(desc.C, device=dev)
    for _ in range(4):
        ops.pm_forward(desc, ml, eps, R, out=out)
        ops.pm_backward(desc, ml, eps, R, gz, None, 1.0, gml=gml, gradius=gR)
    torch.cuda.synchronize()
else:
    def planes_of(x, ones_col=False):
        buf = ops.PlaneBuf(x.shape[0], x.shape[1], 2, dev, ones_col=ones_col)
        ops.split_planes(x, buf)
        return buf
    for (M, N, K) in [(4096, 400, 784), (4096, 400, 8), (4096, 784, 400)]:
        x = torch.randn(M, K, device=dev, generator=g)
        W = torch.randn(N, K, device=dev, generator=g)
        b = torch.randn(N, device=dev, generator=g)
        xp, Wp = planes_of(x), planes_of(W)
        outp = ops.PlaneBuf(M, N, 2, dev, ones_col=True)
        for _ in range(3):
            ops.gemm(xp, Wp, M, N, K, epilogue=_lib.EPI_BIAS_RELU, bias=b, out_planes=outp)
    torch.cuda.synchronize()
