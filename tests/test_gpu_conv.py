"""GPU parity tests of the convolutional VAE (BASELINE cfg5): the data-movement kernels around
the tcgen05 GEMM (bit-exact against plain PyTorch indexing), the whole ConvolutionalVAE step against the fixtures the
reference's own model produced (tests/golden/conv_*.npz, conv_vae.py:28-79) and against the float64 oracle at a
cfg5-sized per-GPU batch, eager and from the CUDA graph with the fused optimizer."""
import numpy as np
import pytest

from helpers import check_conv_digest, conv_golden_names, conv_params_from_seed, load_golden, normwise

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mvae_b200 import _lib
    _lib.lib()
    return torch.device("cuda:0")


def _planes_of(x, planes, dev, ones_col=False):
    from mvae_b200 import ops
    buf = ops.PlaneBuf(x.shape[0], x.shape[1], planes, dev, ones_col=ones_col)
    ops.split_planes(x.to(dev).contiguous(), buf)
    return buf


@pytest.mark.parametrize("B,H,C,planes", [(3, 8, 16, 3), (2, 32, 3, 3), (2, 16, 64, 2), (1, 4, 256, 1)])
def test_im2col_is_an_exact_gather(dev, B, H, C, planes):
    """mvae_conv_im2col against torch indexing (k4, s2, p1; channels-last), plane by plane, bit for bit."""
    from mvae_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + C)
    x = torch.randn(B * H * H, C, generator=g)
    src = _planes_of(x, planes, dev)
    OH = H // 2
    dst = ops.PlaneBuf(B * OH * OH, 16 * C, planes, dev, ones_col=True)
    dst.t.fill_(7.0)
    ops.conv_im2col(src, B, H, H, C, dst, ones_col=True)
    img = src.t[:, :, :C].reshape(planes, B, H, H, C)
    pad = torch.zeros(planes, B, H + 2, H + 2, C, dtype=img.dtype, device=dev)
    pad[:, :, 1:-1, 1:-1] = img
    want = torch.empty(planes, B, OH, OH, 16, C, dtype=img.dtype, device=dev)
    for ky in range(4):
        for kx in range(4):
            want[:, :, :, :, ky * 4 + kx] = pad[:, :, ky:ky + 2 * OH:2, kx:kx + 2 * OH:2]
    assert torch.equal(dst.t[:, :, :16 * C], want.reshape(planes, B * OH * OH, 16 * C))
    assert bool((dst.t[0, :, 16 * C] == 1).all()) and (planes == 1 or bool((dst.t[1:, :, 16 * C] == 0).all()))


@pytest.mark.parametrize("B,H,C,act", [(2, 4, 8, 1), (3, 8, 3, 0), (2, 16, 64, 2), (1, 4, 256, 1)])
def test_col2im_is_the_adjoint_gather(dev, B, H, C, act):
    """mvae_conv_col2im = the adjoint of the patch gather (+ bias, relu / mask), against a float64 scatter-add."""
    from mvae_b200 import ops
    g = torch.Generator().manual_seed(B + 10 * C)
    cols = torch.randn(B * H * H, 16 * C, generator=g)
    bias = torch.randn(C, generator=g)
    OH = 2 * H
    mask_src = torch.randn(B * OH * OH, C, generator=g)
    want = torch.zeros(B, OH + 2, OH + 2, C, dtype=torch.float64)
    c6 = cols.double().reshape(B, H, H, 4, 4, C)
    for ky in range(4):
        for kx in range(4):
            want[:, ky:ky + 2 * H:2, kx:kx + 2 * H:2] += c6[:, :, :, ky, kx]
    want = want[:, 1:-1, 1:-1].reshape(B * OH * OH, C) + bias.double()
    if act == 1:
        want = want.clamp_min(0)
    elif act == 2:
        want = want * (mask_src.double() > 0)
    out_p = ops.PlaneBuf(B * OH * OH, C, 3, dev)
    out_f = torch.empty(B * OH * OH, C, device=dev)
    mask = _planes_of(mask_src, 2, dev) if act == 2 else None
    ops.conv_col2im(cols.to(dev), B, H, H, C, bias=bias.to(dev), act=act, mask=mask, out_planes=out_p, out_f32=out_f)
    assert normwise(out_f.cpu().numpy(), want.numpy()) < 1e-6
    assert normwise(out_p.to_float().cpu().numpy(), want.numpy()) < 1e-6


def test_layout_and_column_sum_kernels(dev):
    from mvae_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, S, C = 5, 16, 24
    x = torch.randn(B, C * S, generator=g).to(dev)
    cl = torch.empty(B * S, C, device=dev)
    ops.permute_sc(x, cl, B, S, C, to_nhwc=True)
    assert torch.equal(cl, x.view(B, C, S).permute(0, 2, 1).reshape(B * S, C))
    back = torch.empty(B, C * S, device=dev)
    ops.permute_sc(cl, back, B, S, C, to_nhwc=False)
    assert torch.equal(back, x)
    src = _planes_of(cl.cpu(), 3, dev)
    dst = ops.PlaneBuf(B, C * S, 3, dev, ones_col=True)
    ops.permute_sc(src, dst, B, S, C, to_nhwc=False)
    assert torch.equal(dst.t[:, :, :C * S].reshape(3, B, C, S), src.t[:, :, :C].reshape(3, B, S, C).permute(0, 1, 3, 2))
    assert bool((dst.t[0, :, C * S] == 1).all())
    for M, Cc in ((1000, 3), (777, 64), (300, 256), (50, 600)):
        m = torch.randn(M, Cc, generator=g)
        out = torch.full((Cc,), 2.0, device=dev)
        ops.colsum(m.to(dev), M, Cc, out)
        assert normwise(out.cpu().numpy() - 2.0, m.double().sum(0).numpy()) < 1e-5
        out.zero_()
        ops.colsum(_planes_of(m, 2, dev), M, Cc, out)
        assert normwise(out.cpu().numpy(), m.double().sum(0).numpy()) < 2e-5


def _build(sig, B, dev, seed=0, radius=1.0, graph=False, from_reference_seed=None):
    from mvae_b200 import components, conv_vae, data, vae
    if from_reference_seed is not None:
        old = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)   # the fixture's model was initialised in float64 (see helpers.py)
        try:
            torch.manual_seed(from_reference_seed)
            model = conv_vae.FusedConvolutionalVAE(8192, components.parse_components(sig, False),
                                                   data.GenericDataset(B, 3072, "bce"), False, device=dev)
        finally:
            torch.set_default_dtype(old)
    else:
        torch.manual_seed(seed)
        model = conv_vae.FusedConvolutionalVAE(8192, components.parse_components(sig, False),
                                               data.GenericDataset(B, 3072, "bce"), False, device=dev)
    with torch.no_grad():
        for rp in model._radius_params:
            if rp is not None:
                rp.fill_(radius)
    model.use_cuda_graph = graph
    opt = vae.FusedCurvatureOptimizer(model, 1e-3, fixed_curvature=False, should_do_curvature_step=lambda: True)
    return model, opt


class _NoOpt:
    def zero_grad(self):
        pass

    def step(self):
        pass


@pytest.mark.parametrize("name", conv_golden_names())
def test_conv_model_matches_reference_golden(dev, name):
    """FusedConvolutionalVAE.forward / train_step against the reference's own ConvolutionalVAE run (float64):
    parameters rebuilt from the fixture's seed (digest checked), state_dict keys / shapes, forward tensors, ELBO,
    and the fixture's strided subsample of every autograd gradient."""
    g, meta = load_golden(name)
    B = g["x"].shape[0]
    model, _ = _build(meta["sig"], B, dev, radius=meta["radius"], from_reference_seed=meta["seed"])
    want_params = conv_params_from_seed(meta["sig"], meta["seed"], meta["radius"])
    sd = model.state_dict()
    assert sorted(sd) == sorted(want_params)
    for k, v in want_params.items():
        assert tuple(sd[k].shape) == tuple(np.asarray(v).shape), k
    check_conv_digest(want_params, g)
    for k, v in want_params.items():   # float32 copies of the reference's float64 initialisation
        assert normwise(sd[k].detach().cpu().numpy(), v) < 1e-6, k
    x = torch.from_numpy(g["x"].astype(np.float32))
    eps = torch.from_numpy(g["eps"].astype(np.float32)).to(dev)
    beta = meta["beta"]
    rep, concat_z, logits = model.forward(x, eps=eps, beta=beta)
    assert normwise(concat_z.cpu().numpy(), g["z"]) < 1e-4
    assert normwise(logits.cpu().numpy(), g["logits"]) < 1e-4
    assert normwise(torch.cat([r.q_z.loc for r in rep], -1).cpu().numpy(), g["mu"]) < 1e-4
    assert normwise(torch.stack([r.kl for r in rep], -1).cpu().numpy(), g["kl"]) < 1e-4
    stats = model.compute_batch_stats(x, logits, rep, beta=beta).convert_to_float()
    assert abs(stats.elbo - g["elbo"]) < 1e-5 * abs(g["elbo"])
    assert tuple(model.encode(x.to(dev)).shape) == (B, 8192)
    assert normwise(model.decode(concat_z).cpu().numpy(), g["logits"]) < 1e-4
    bs, _ = model.train_step(_NoOpt(), x, beta, eps=eps)
    assert abs(bs.elbo - g["elbo"]) < 1e-5 * abs(g["elbo"])
    assert abs(bs.bce - g["bce_sum"]) < 1e-5 * abs(g["bce_sum"])
    stride = lambda n: max(1, n // meta["subsample"])  # noqa: E731
    for k, p in model.named_parameters():
        got = p.grad.detach().cpu().double().numpy().reshape(-1)   # logical (reference) order of the permuted view
        sub = got[::stride(got.size)]
        if got.size == 1:
            assert abs(sub[0] - g["gsub." + k][0]) < 2e-4 * max(1.0, abs(g["gsub." + k][0])), k
            continue
        scale = g["gnorm." + k][1]   # max |gradient| of the full tensor
        assert np.abs(sub - g["gsub." + k]).max() < 3e-4 * scale, (k, np.abs(sub - g["gsub." + k]).max() / scale)
        assert abs(np.linalg.norm(got) - g["gnorm." + k][0]) < 1e-3 * g["gnorm." + k][0], k


def test_conv_steps_from_the_graph_vs_oracle(dev, oracle):
    """cfg5 shape (256 rows per GPU): 3 steps of the fused optimizer from the CUDA graph against the float64 oracle's
    ConvolutionalVAE step + Adam + radii SGD (statistics per step, movement of every tensor at the end)."""
    sig, B = "h2,s2,e2", 256
    model, opt = _build(sig, B, dev, seed=3, radius=10.0, graph=True)
    p0 = {k: v.detach().cpu().double().numpy() for k, v in model.state_dict().items()}
    ov = oracle.OracleConvVAE(sig)
    tr = oracle.OracleTrainer(ov, p0)
    g = torch.Generator().manual_seed(5)
    for i in range(3):
        x = torch.rand(B, 3072, generator=g)
        eps = torch.randn(B, model.desc.ld_eps, generator=g)
        bs, _ = model.train_step(opt, x.to(dev), 0.9, eps=eps.to(dev))
        ref = tr.step(x.double().numpy(), eps.double().numpy(), 0.9)
        assert abs(bs.elbo - ref["elbo"]) < 3e-5 * abs(ref["elbo"]), (i, bs.elbo, ref["elbo"])
        # (at R = 10 the KL is ~0.3 per sample against a BCE of ~2100: the bar is on the ELBO; the KL term follows the
        # parameters, on which Adam amplifies float32-level gradient differences of rarely active units)
        assert abs(bs.kl - ref["kl_sum"]) < 1e-3 * abs(ref["kl_sum"]) + 1e-3, (i, bs.kl, ref["kl_sum"])
    errs = {}
    for k, v in model.state_dict().items():
        den = float(np.linalg.norm(tr.params[k] - p0[k]))
        if den > 0:
            errs[k] = float(np.linalg.norm(v.detach().cpu().double().numpy() - tr.params[k])) / den
    worst = max(errs, key=errs.get)
    print(f"\n[conv {sig}] graph + fused optimizer vs oracle after 3 steps: worst movement error {errs[worst]:.2e} ({worst})")
    med = sorted(errs.values())[len(errs) // 2]
    print(f"[conv {sig}] median movement error {med:.2e}")
    # Five relu layers whose kink flips are not fed back, 256 rows per step: bias gradients that nearly cancel over the
    # batch get a large RELATIVE float32 error, which Adam's normalisation turns into a visible difference of that
    # tensor's movement (measured: worst 7e-2 on a head bias, median 1e-3).  The gradients themselves are held to
    # 3e-4 by the reference fixtures above; this test is about the graph / optimizer plumbing, where errors are O(1).
    assert errs[worst] < 0.25 and med < 2e-2, (worst, errs[worst], med)
    for nm in ("e0", "d3"):   # operand planes follow the updated filters (master layout)
        assert normwise(model._Wp[nm].to_float().cpu().numpy(), model._W[nm].detach().cpu().numpy()) < 1e-4
    # evaluation API on the conv model
    ll, mi, cov = model.log_likelihood(torch.rand(8, 3072, generator=g).to(dev), n=3)
    assert ll.shape == (8,) and torch.isfinite(ll).all() and torch.isfinite(mi).all()
