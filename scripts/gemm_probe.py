"""Diagnostic (not a test): run the tcgen05 GEMM on a few shapes / operand layouts and print error statistics.
Usage on the GPU box: python scripts/gemm_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvae_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")


def planes_of(x, ones_col=False, planes=2):
    buf = ops.PlaneBuf(x.shape[0], x.shape[1], planes, dev, ones_col=ones_col)
    ops.split_planes(x, buf)
    return buf


def report(tag, out, ref):
    out = out.double()
    bad = ~torch.isfinite(out)
    err = (out - ref).abs()
    err[bad] = float("inf")
    nw = (err.max() / ref.abs().max()).item()
    print(f"{tag}: normwise={nw:.3e} nonfinite={int(bad.sum())} mean|out|={out[~bad].abs().mean().item():.4f} "
          f"mean|ref|={ref.abs().mean().item():.4f}", flush=True)
    if not nw < 1e-3:
        r = min(4, out.shape[0])
        c = min(8, out.shape[1])
        print("  out[:4,:8]", out[:r, :c].cpu().numpy().round(3).tolist())
        print("  ref[:4,:8]", ref[:r, :c].cpu().numpy().round(3).tolist())
        rows_bad = (err.max(1).values > 1e-3 * ref.abs().max()).nonzero().flatten()[:10].tolist()
        cols_bad = (err.max(0).values > 1e-3 * ref.abs().max()).nonzero().flatten()[:10].tolist()
        print("  first bad rows", rows_bad, "first bad cols", cols_bad)


def main():
    g = torch.Generator(device=dev).manual_seed(0)
    for (M, N, K) in [(128, 16, 64), (128, 64, 16), (256, 128, 128), (300, 50, 70), (4096, 400, 784)]:
        x = torch.randn(M, K, device=dev, generator=g)
        W = torch.randn(N, K, device=dev, generator=g)
        out = torch.full((M, N), float("nan"), device=dev)
        ops.gemm(planes_of(x), planes_of(W), M, N, K, out_f32=out)
        torch.cuda.synchronize()
        report(f"KK  M{M} N{N} K{K}", out, x.double() @ W.double().t())
        # single plane (pure bf16 product) to separate plane logic from layout logic
        out1 = torch.full((M, N), float("nan"), device=dev)
        ops.gemm(planes_of(x, planes=1), planes_of(W, planes=1), M, N, K, out_f32=out1)
        torch.cuda.synchronize()
        report(f"KK1 M{M} N{N} K{K}", out1, x.bfloat16().double() @ W.bfloat16().double().t())
        Wt = W.t().contiguous()  # [K, N]
        out2 = torch.full((M, N), float("nan"), device=dev)
        ops.gemm(planes_of(x), planes_of(Wt), M, N, K, b_major=_lib.MN_MAJOR, out_f32=out2)
        torch.cuda.synchronize()
        report(f"K-MN M{M} N{N} K{K}", out2, x.double() @ W.double().t())
        xt = x.t().contiguous()  # [K, M]
        out3 = torch.full((M, N), float("nan"), device=dev)
        ops.gemm(planes_of(xt), planes_of(Wt), M, N, K, a_major=_lib.MN_MAJOR, b_major=_lib.MN_MAJOR, out_f32=out3)
        torch.cuda.synchronize()
        report(f"MN-MN M{M} N{N} K{K}", out3, x.double() @ W.double().t())
    # timing of the big shapes
    for (M, N, K) in [(4096, 400, 784), (4096, 784, 400), (8192, 784, 400)]:
        x = torch.randn(M, K, device=dev, generator=g)
        W = torch.randn(N, K, device=dev, generator=g)
        xp, Wp = planes_of(x), planes_of(W)
        out = torch.empty(M, N, device=dev)
        for _ in range(3):
            ops.gemm(xp, Wp, M, N, K, out_f32=out)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            ops.gemm(xp, Wp, M, N, K, out_f32=out)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        print(f"time M{M} N{N} K{K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s (fp32-equivalent)", flush=True)


if __name__ == "__main__":
    main()
