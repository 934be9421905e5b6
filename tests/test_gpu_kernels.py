"""GPU parity tests (run on the B200 box with `-m gpu`): every CUDA kernel, called through the C ABI, against the
CPU oracle (oracle/) and the golden vectors produced by the reference itself (tests/golden/).

Tolerances (BASELINE.json north_star: 1e-4 relative, fp32; defined normwise — max|a-b| / max|b| — per tensor,
SURVEY.md App. E, because the reference's own fp32 path is only that accurate elementwise):
  forward tensors            normwise <= 1e-4 against the float64 golden / float64 oracle
  gradients                  normwise <= max(5e-4, 3 x the reference's own fp32-vs-fp64 error) (same bar as the oracle's f32 test)
  sums (ELBO, bce, kl)       relative <= 1e-5
  index / layout work        bit-exact (transposes, plane splits reassemble exactly to the rounded value)
"""
import zlib

import numpy as np
import pytest

from helpers import load_golden, normwise, pack_ml, pm_golden_names, radii_array, unpack_gml

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

FWD_TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mvae_b200 import _lib
    _lib.lib()  # fail loudly if the extension is missing
    return torch.device("cuda:0")


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)


# ------------------------------------------------------------------------------------------------ K3 forward
@pytest.mark.parametrize("name", pm_golden_names())
def test_pm_forward_golden(dev, oracle, name):
    from mvae_b200 import ops
    g, meta = load_golden(name)
    desc = ops.make_desc(meta["sig"], scalar_parametrization=meta["scalar_parametrization"])
    odesc = oracle.make_desc(meta["sig"], scalar_parametrization=meta["scalar_parametrization"])
    ml = pack_ml(odesc, g["m"].astype(np.float32), g["l"].astype(np.float32))
    R = radii_array(g["radii"], np.float32)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    out = ops.pm_forward(desc, _t(ml, dev), _t(g["eps"], dev), _t(R, dev), want_mu_sigma=True, flag=flag)
    torch.cuda.synchronize()
    for k in ("z", "kl", "mu", "sigma"):
        got = out[k].cpu().numpy()
        assert normwise(got, g[k]) < FWD_TOL, (k, normwise(got, g[k]))
        # against the reference's own float32 run — itself up to ~1e-3 from its float64 run on the small-radius fixtures
        assert normwise(got, g[k + "_f32"]) < max(2e-4, 3 * normwise(g[k + "_f32"], g[k])), k
    assert int(flag.item()) == 0
    # mu/sigma are optional outputs: the result must not depend on asking for them
    out2 = ops.pm_forward(desc, _t(ml, dev), _t(g["eps"], dev), _t(R, dev))
    assert torch.equal(out2["z"], out["z"]) and torch.equal(out2["kl"], out["kl"])


@pytest.mark.parametrize("name", pm_golden_names())
def test_pm_backward_golden(dev, oracle, name):
    from mvae_b200 import ops
    g, meta = load_golden(name)
    desc = ops.make_desc(meta["sig"], scalar_parametrization=meta["scalar_parametrization"])
    odesc = oracle.make_desc(meta["sig"], scalar_parametrization=meta["scalar_parametrization"])
    ml = pack_ml(odesc, g["m"].astype(np.float32), g["l"].astype(np.float32))
    R = radii_array(g["radii"], np.float32)
    gml, gR = ops.pm_backward(desc, _t(ml, dev), _t(g["eps"], dev), _t(R, dev), _t(g["gz"], dev), _t(g["gkl"], dev))
    torch.cuda.synchronize()
    gm, gl = unpack_gml(odesc, gml.cpu().numpy())
    ref_err_m = normwise(g["gm_f32"], g["gm"])
    ref_err_l = normwise(g["gl_f32"], g["gl"])
    assert normwise(gm, g["gm"]) < max(5e-4, 3 * ref_err_m)
    assert normwise(gl, g["gl"]) < max(5e-4, 3 * ref_err_l)
    ref_err_R = normwise(g["gR_f32"], g["gR"])
    assert normwise(gR.cpu().numpy(), g["gR"]) < max(1e-3, 3 * ref_err_R)


@pytest.mark.parametrize("sig,B,Rs,scale_m", [
    ("h2,s2,e2", 4096, [1.0, 1.0, 0.0], 1.0),           # BASELINE cfg2
    ("h6,h6,s6,s6,e6", 8192 + 37, [1.3, 0.9, 1.1, 2.0, 0.0], 0.7),  # cfg3, ragged
    ("h2", 16384, [1.0], 1.0),                           # cfg4a
    ("p2", 16384, [1.0], 0.8),                           # cfg4b
    ("e2", 128, [0.0], 1.0),                             # cfg1
    ("h3,s5,p4,e7,h8,s1", 1000, [1.0, 2.0, 1.5, 0.0, 0.5, 1.0], 0.5),  # mixed static / dynamic dims
    ("12e2,12h2,12s2", 517, [0.0] * 12 + [1.0] * 24, 1.0),  # many components
    ("h2,s2,e2", 1, [1.0, 1.0, 0.0], 1.0),               # single row
])
def test_pm_forward_backward_vs_oracle(dev, oracle, sig, B, Rs, scale_m):
    """Seeded inputs at the BASELINE config sizes: CUDA vs the float64 oracle (which is pinned to the reference)."""
    from mvae_b200 import ops
    desc = ops.make_desc(sig)
    odesc = oracle.make_desc(sig)
    rng = np.random.default_rng(zlib.crc32(sig.encode()) + B)
    ml = rng.standard_normal((B, desc.ld_ml)) * scale_m
    eps = rng.standard_normal((B, desc.ld_eps))
    R = np.asarray([r if r else 1.0 for r in Rs])
    gz = rng.standard_normal((B, desc.ld_z))
    gkl = rng.standard_normal((B, desc.C))
    ref = oracle.pm_forward(odesc, ml, eps, R, want=("z", "kl", "mu", "sigma"))
    out = ops.pm_forward(desc, _t(ml, dev), _t(eps, dev), _t(R, dev), want_mu_sigma=True)
    for k in ("z", "kl", "mu", "sigma"):
        err = normwise(out[k].cpu().numpy(), ref[k])
        assert err < FWD_TOL, (k, err)
    # sum over the batch of the KL terms: the quantity the ELBO consumes
    got_kl = out["kl"].double().sum(0).cpu().numpy()
    np.testing.assert_allclose(got_kl, ref["kl"].sum(0), rtol=2e-5, atol=1e-3)
    rgml, rgR = oracle.pm_backward(odesc, ml, eps, R, gz, gkl)
    f32 = np.float32
    o32, _ = oracle.pm_backward(odesc, ml.astype(f32), eps.astype(f32), R.astype(f32), gz.astype(f32), gkl.astype(f32))
    gml, gR = ops.pm_backward(desc, _t(ml, dev), _t(eps, dev), _t(R, dev), _t(gz, dev), _t(gkl, dev))
    bar = max(5e-4, 3 * normwise(o32, rgml))
    assert normwise(gml.cpu().numpy(), rgml) < bar
    assert normwise(gR.cpu().numpy(), rgR) < 2e-3
    # scalar gkl (the beta path used by train_step)
    rgml2, rgR2 = oracle.pm_backward(odesc, ml, eps, R, gz, None, 0.7)
    gml2, gR2 = ops.pm_backward(desc, _t(ml, dev), _t(eps, dev), _t(R, dev), _t(gz, dev), None, 0.7)
    assert normwise(gml2.cpu().numpy(), rgml2) < bar
    assert normwise(gR2.cpu().numpy(), rgR2) < 2e-3


def test_pm_edge_cases(dev, oracle):
    from mvae_b200 import _lib, ops
    desc = ops.make_desc("h2,s2,e2")
    # empty batch is a no-op
    out = ops.pm_forward(desc, torch.zeros(0, 12, device=dev), torch.zeros(0, 6, device=dev), torch.ones(3, device=dev))
    assert out["z"].shape == (0, 8)
    # unsupported manifold kind and bad descriptor are rejected with a status, not a crash
    with pytest.raises(_lib.MvaeError):
        ops.make_desc([_lib.PROJ_SPHERE + 3], [2])
    bad = ops.make_desc("h2")
    bad.comp[0].d = 7
    with pytest.raises(_lib.MvaeError):
        ops.pm_forward(bad, torch.zeros(4, 4, device=dev), torch.zeros(4, 2, device=dev), torch.ones(1, device=dev))
    # non-finite inputs raise the device flag instead of synchronising
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    ml = torch.zeros(64, 12, device=dev)
    ml[5, 0] = float("nan")
    ops.pm_forward(desc, ml, torch.zeros(64, 6, device=dev), torch.ones(3, device=dev), flag=flag)
    assert int(flag.item()) == 1
    # zero head pre-activations (m = 0): finite outputs, mu = mu_0
    flag.zero_()
    out = ops.pm_forward(desc, torch.zeros(64, 12, device=dev), torch.randn(64, 6, device=dev),
                         torch.ones(3, device=dev), want_mu_sigma=True, flag=flag)
    assert int(flag.item()) == 0 and torch.isfinite(out["z"]).all() and torch.isfinite(out["kl"]).all()
    np.testing.assert_allclose(out["mu"][0].cpu().numpy(), [1, 0, 0, 1, 0, 0, 0, 0], atol=1e-7)
    # unaligned views take the scalar path and give identical results
    base_ml = torch.randn(100 * 12 + 1, device=dev)
    base_eps = torch.randn(100 * 6 + 1, device=dev)
    a = ops.pm_forward(desc, base_ml[1:].view(100, 12), base_eps[1:].view(100, 6), torch.ones(3, device=dev))
    b = ops.pm_forward(desc, base_ml[1:].view(100, 12).clone(), base_eps[1:].view(100, 6).clone(),
                       torch.ones(3, device=dev))
    assert torch.equal(a["z"], b["z"]) and torch.equal(a["kl"], b["kl"])


def test_pm_properties_large(dev):
    """Size-independent properties at a batch far beyond what the oracle is run on: points lie on their manifolds
    (<z,z>_L = -R^2, |z|^2 = R^2), the Euclidean block is mu + eps*sigma, and the result is independent of tiling."""
    from mvae_b200 import ops
    desc = ops.make_desc("h2,s2,e2")
    B = 1 << 20
    g = torch.Generator(device=dev).manual_seed(0)
    ml = torch.randn(B, 12, device=dev, generator=g)
    eps = torch.randn(B, 6, device=dev, generator=g)
    R = torch.tensor([1.5, 0.8, 1.0], device=dev)
    out = ops.pm_forward(desc, ml, eps, R, want_mu_sigma=True)
    z = out["z"].double()
    lor = (z[:, 1:3]**2).sum(-1) - z[:, 0]**2
    rel = ((lor + 1.5**2).abs() / z[:, 0]**2)
    # relative to the cancelling terms; the worst of 2^20 samples has cosh(t)cosh(a) - sinh(t)sinh(a) cancel ~1e3-fold
    assert rel.max().item() < 2e-4 and rel.median().item() < 1e-6
    sph = (z[:, 3:6]**2).sum(-1)
    assert (sph - 0.8**2).abs().max().item() < 1e-5
    e = out["mu"][:, 6:8] + eps[:, 4:6] * out["sigma"][:, 4:6]
    assert (e - out["z"][:, 6:8]).abs().max().item() < 1e-6
    # a sub-batch that starts mid-tile gives bit-identical rows
    sub = ops.pm_forward(desc, ml[777:5000].contiguous(), eps[777:5000].contiguous(), R)
    assert torch.equal(sub["z"], out["z"][777:5000]) and torch.equal(sub["kl"], out["kl"][777:5000])


# ------------------------------------------------------------------------------------------ K5: recon + ELBO
@pytest.mark.parametrize("kind,B,D", [("bce", 4096, 784), ("nll", 16384, 50), ("bce", 37, 13), ("nll", 5, 3)])
def test_recon_loss(dev, oracle, kind, B, D):
    from mvae_b200 import ops
    rng = np.random.default_rng(B + D)
    lg = rng.standard_normal((B, D)) * 3
    x = (rng.random((B, D)) < 0.3).astype(np.float64) if kind == "bce" else rng.standard_normal((B, D))
    rs_ref, g_ref = oracle.recon(kind, lg, x, want_grad=True)
    rs, g = ops.recon_loss(kind, _t(lg, dev), _t(x, dev), want_grad=True)
    assert normwise(rs.cpu().numpy(), rs_ref) < 1e-5
    assert normwise(g.cpu().numpy(), g_ref) < 1e-5
    rs2, g2 = ops.recon_loss(kind, _t(lg, dev), _t(x, dev), want_grad=False)
    assert g2 is None and torch.equal(rs, rs2)


def test_recon_extreme_logits(dev):
    from mvae_b200 import ops
    lg = torch.tensor([[-200., -30., 0., 30., 200., 1e4, -1e4, 88.]], device=dev)
    x = torch.tensor([[0., 1., 1., 0., 1., 1., 0., 0.]], device=dev)
    rs, g = ops.recon_loss("bce", lg, x, want_grad=True)
    ref = torch.nn.functional.binary_cross_entropy_with_logits(lg.double(), x.double(), reduction="none").sum(-1)
    assert torch.isfinite(rs).all() and abs(rs.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert torch.allclose(g, torch.sigmoid(lg) - x, atol=1e-7)


@pytest.mark.parametrize("B,C", [(4096, 3), (8192, 5), (1, 1), (100003, 7), (300, 96)])
def test_elbo_reduce(dev, oracle, B, C):
    from mvae_b200 import ops
    rng = np.random.default_rng(B * 7 + C)
    bce = rng.random(B) * 100 + 50
    kl = rng.standard_normal((B, C)) + 3
    ref = oracle.elbo(bce, kl, 0.7)
    out = ops.elbo_reduce(_t(bce, dev), _t(kl, dev), 0.7).cpu().numpy()
    np.testing.assert_allclose(out, ref, rtol=1e-5)


# ------------------------------------------------------------------------------------------ K7: optimizers
def test_adam_matches_torch(dev):
    from mvae_b200 import ops
    torch.manual_seed(0)
    n = 100_003
    p0 = torch.randn(n, device=dev)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-3)
    p = p0.clone()
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn(n, device=dev) * (10.0 ** (step - 3))
        p_ref.grad = g.clone()
        opt.step()
        ops.adam_step(p, g, m, v, 1e-3, step)
        assert torch.allclose(p, p_ref.detach(), rtol=1e-6, atol=1e-7), step
    q = p.clone()
    gq = torch.randn(n, device=dev)
    ops.sgd_step(q, gq, 1e-4)
    assert torch.allclose(q, p - 1e-4 * gq, rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------------------------------ split planes
@pytest.mark.parametrize("R,K,planes", [(4096, 784, 2), (100, 13, 2), (33, 400, 3), (7, 8, 1)])
def test_split_planes(dev, R, K, planes):
    from mvae_b200 import ops
    g = torch.Generator(device=dev).manual_seed(R + K)
    x = torch.randn(R, K, device=dev, generator=g) * 3
    d = ops.PlaneBuf(R, K, planes, dev)
    dt = ops.PlaneBuf(K, R, planes, dev)
    ops.split_planes(x, d, dt)
    # plane p is bf16(x - sum of earlier planes): bit-exact against torch's round-to-nearest-even conversion
    rem = x.clone()
    for p in range(planes):
        h = rem.to(torch.bfloat16)
        assert torch.equal(d.t[p, :, :K], h), p
        assert torch.equal(dt.t[p, :, :R], h.t()), p
        rem = rem - h.float()
    rel = (d.to_float() - x).abs().max() / x.abs().max()
    assert rel.item() < {1: 2**-8, 2: 2**-16, 3: 2**-23}[planes]


# ------------------------------------------------------------------------------------------------ tcgen05 GEMM
def _planes_of(x, dev, planes=2, ones_col=False):
    from mvae_b200 import ops
    buf = ops.PlaneBuf(x.shape[0], x.shape[1], planes, dev, ones_col=ones_col)
    ops.split_planes(x, buf)
    return buf


GEMM_TOL = 3e-5  # two bf16 planes per operand, three products: ~2^-16 relative per term


@pytest.mark.parametrize("M,N,K", [(4096, 400, 784), (4096, 784, 400), (4096, 12, 400), (4096, 400, 8), (300, 50, 70),
                                   (128, 16, 64), (1, 1, 1), (130, 257, 129)])
def test_gemm_forward_kmajor(dev, M, N, K):
    """y = x W^T + b with K-major operands (nn.Linear forward, ffnn_vae.py:48,56-57; component.py:64,69)."""
    from mvae_b200 import _lib, ops
    g = torch.Generator(device=dev).manual_seed(M * 3 + N * 5 + K)
    x = torch.randn(M, K, device=dev, generator=g)
    W = torch.randn(N, K, device=dev, generator=g) / K**0.5
    b = torch.randn(N, device=dev, generator=g)
    out = torch.full((M, N), float("nan"), device=dev)
    ops.gemm(_planes_of(x, dev), _planes_of(W, dev), M, N, K, bias=b, out_f32=out)
    ref = x.double() @ W.double().t() + b.double()
    err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    assert err < GEMM_TOL, err


@pytest.mark.parametrize("M,N,K", [(4096, 400, 784), (4096, 8, 400), (4096, 400, 12), (200, 100, 30)])
def test_gemm_dgrad_b_mnmajor(dev, M, N, K):
    """gx = gy W with W [K=out, N=in] read MN-major (no transposed copy of the weight)."""
    from mvae_b200 import _lib, ops
    g = torch.Generator(device=dev).manual_seed(M + N * 11 + K)
    gy = torch.randn(M, K, device=dev, generator=g)
    W = torch.randn(K, N, device=dev, generator=g)
    out = torch.full((M, N), float("nan"), device=dev)
    ops.gemm(_planes_of(gy, dev), _planes_of(W, dev), M, N, K, b_major=_lib.MN_MAJOR, out_f32=out)
    ref = gy.double() @ W.double()
    err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    assert err < GEMM_TOL, err


@pytest.mark.parametrize("B,OUT,IN,split", [(4096, 784, 400, 1), (4096, 400, 784, 4), (4096, 12, 400, 8),
                                            (4096, 400, 8, 2), (1000, 70, 33, 3)])
def test_gemm_wgrad_both_mnmajor(dev, B, OUT, IN, split):
    """gW = gy^T x and gb = column sums of gy, both operands read MN-major from the row-major [B, features]
    buffers; the bias gradient comes from the ones column of the input planes (column IN)."""
    from mvae_b200 import _lib, ops
    g = torch.Generator(device=dev).manual_seed(B + OUT * 13 + IN)
    gy = torch.randn(B, OUT, device=dev, generator=g)
    x = torch.randn(B, IN, device=dev, generator=g)
    gW = torch.zeros(OUT, IN, device=dev)
    gb = torch.zeros(OUT, device=dev)
    ops.gemm(_planes_of(gy, dev), _planes_of(x, dev, ones_col=True), OUT, IN + 1, B, a_major=_lib.MN_MAJOR,
             b_major=_lib.MN_MAJOR, split_k=split, out_f32=gW, out_col=gb, col_split=IN)
    refW = gy.double().t() @ x.double()
    refb = gy.double().sum(0)
    assert ((gW.double() - refW).abs().max() / refW.abs().max()).item() < GEMM_TOL
    assert ((gb.double() - refb).abs().max() / refb.abs().max()).item() < GEMM_TOL


def test_gemm_epilogues(dev, oracle):
    from mvae_b200 import _lib, ops
    g = torch.Generator(device=dev).manual_seed(5)
    M, N, K = 520, 400, 96
    x = torch.randn(M, K, device=dev, generator=g)
    W = torch.randn(N, K, device=dev, generator=g) / K**0.5
    b = torch.randn(N, device=dev, generator=g)
    xp, Wp = _planes_of(x, dev), _planes_of(W, dev)
    pre = x.double() @ W.double().t() + b.double()
    # bias + relu -> planes with a ones column that must survive the epilogue
    hp = ops.PlaneBuf(M, N, 2, dev, ones_col=True)
    h32 = torch.empty(M, N, device=dev)
    ops.gemm(xp, Wp, M, N, K, epilogue=_lib.EPI_BIAS_RELU, bias=b, out_planes=hp, out_f32=h32)
    ref = pre.clamp(min=0)
    assert ((h32.double() - ref).abs().max() / ref.abs().max()).item() < GEMM_TOL
    assert ((hp.to_float().double() - ref).abs().max() / ref.abs().max()).item() < GEMM_TOL
    assert torch.equal(hp.t[0, :, N].float(), torch.ones(M, device=dev)) and hp.t[1, :, N].abs().max().item() == 0
    assert torch.equal(hp.t[0, :, :N] > 0, h32 > 0)
    # relu mask from plane 0 of the activation
    gy = torch.randn(M, K, device=dev, generator=g)
    outp = ops.PlaneBuf(M, N, 2, dev)
    ops.gemm(_planes_of(gy, dev), Wp, M, N, K, epilogue=_lib.EPI_RELU_MASK, mask=hp, out_planes=outp)
    refm = (gy.double() @ W.double().t()) * (ref > 0)
    assert ((outp.to_float().double() - refm).abs().max() / refm.abs().max()).item() < GEMM_TOL
    # BCE / NLL row sums + dloss/dlogits planes (+ logits)
    for kind, epi in (("bce", _lib.EPI_BCE_ROWSUM), ("nll", _lib.EPI_NLL_ROWSUM)):
        tgt = (torch.rand(M, N, device=dev, generator=g) < 0.3).float() if kind == "bce" else torch.randn(
            M, N, device=dev, generator=g)
        rows = torch.zeros(M, device=dev)
        gl = ops.PlaneBuf(M, N, 2, dev)
        logits = torch.empty(M, N, device=dev)
        ops.gemm(xp, Wp, M, N, K, epilogue=epi, bias=b, aux=tgt, rowsum=rows, out_planes=gl, out_f32=logits)
        rs_ref, g_ref = oracle.recon(kind, pre.cpu().numpy(), tgt.double().cpu().numpy(), want_grad=True)
        assert normwise(logits.cpu().numpy(), pre.cpu().numpy()) < GEMM_TOL
        assert normwise(rows.cpu().numpy(), rs_ref) < 1e-5
        assert normwise(gl.to_float().cpu().numpy(), g_ref) < 5e-5


def test_gemm_rejects_bad_arguments(dev):
    from mvae_b200 import _lib, ops
    a = ops.PlaneBuf(64, 64, 2, dev)
    out = torch.empty(64, 64, device=dev)
    with pytest.raises(_lib.MvaeError):
        ops.gemm(a, a, 64, 64, 64, epilogue=_lib.EPI_BCE_ROWSUM, out_f32=out)       # no targets
    with pytest.raises(_lib.MvaeError):
        ops.gemm(a, a, 64, 64, 64, epilogue=_lib.EPI_BIAS_RELU, split_k=2, out_f32=out)  # split-K needs STORE
    with pytest.raises(_lib.MvaeError):
        ops.gemm(a, a, 64, 64, 0, out_f32=out)


# ------------------------------------------------------------------------------------------ skinny dense layers
@pytest.mark.parametrize("B,K,N", [(4096, 400, 12), (8229, 400, 60), (1000, 37, 8), (5, 400, 1)])
def test_skinny_rowdot(dev, B, K, N):
    """Heads forward (component.py:64,69) and dgrad into z: exact fp32 row dots from fp32 or 3-plane operands."""
    from mvae_b200 import ops
    g = torch.Generator(device=dev).manual_seed(B + K + N)
    a = torch.randn(B, K, device=dev, generator=g)
    W = torch.randn(N, K, device=dev, generator=g) / K**0.5
    b = torch.randn(N, device=dev, generator=g)
    ref = a.double() @ W.double().t() + b.double()
    out = torch.full((B, N), float("nan"), device=dev)
    ops.skinny_rowdot(a, W, K, 1, K=K, N=N, bias=b, out=out)
    assert ((out.double() - ref).abs().max() / ref.abs().max()).item() < 2e-6
    ap = _planes_of(a, dev, planes=3, ones_col=True)
    out2 = torch.full((B, N), float("nan"), device=dev)
    ops.skinny_rowdot((ap, 3), W, K, 1, K=K, N=N, bias=b, out=out2)
    assert ((out2.double() - ref).abs().max() / ref.abs().max()).item() < 2e-6
    # transposed weight access: out = a Wt with Wt [K, N] stored row-major
    Wt = W.t().contiguous()
    out3 = torch.full((B, N), float("nan"), device=dev)
    ops.skinny_rowdot(a, Wt, 1, N, K=K, N=N, bias=None, out=out3)
    assert ((out3.double() - (ref - b.double())).abs().max() / ref.abs().max()).item() < 2e-6


@pytest.mark.parametrize("B,K,N", [(4096, 8, 400), (1000, 34, 400), (77, 12, 130), (4096, 60, 400)])
def test_skinny_expand(dev, B, K, N):
    """fc_d0 forward (ffnn_vae.py:56) with relu -> planes, and the relu-masked dgrad into h."""
    from mvae_b200 import ops
    g = torch.Generator(device=dev).manual_seed(B * 3 + K + N)
    a = torch.randn(B, K, device=dev, generator=g)
    W = torch.randn(N, K, device=dev, generator=g)
    b = torch.randn(N, device=dev, generator=g)
    pre = a.double() @ W.double().t() + b.double()
    outp = ops.PlaneBuf(B, N, 2, dev, ones_col=True)
    out32 = torch.empty(B, N, device=dev)
    ops.skinny_expand(a, W, K, 1, K=K, N=N, bias=b, act=ops.ACT_RELU, out_planes=outp, out_f32=out32)
    ref = pre.clamp(min=0)
    assert ((out32.double() - ref).abs().max() / ref.abs().max()).item() < 2e-6
    assert ((outp.to_float().double() - ref).abs().max() / ref.abs().max()).item() < 2e-5
    assert torch.equal(outp.t[0, :, N].float(), torch.ones(B, device=dev))  # ones column untouched
    ga = torch.randn(B, K, device=dev, generator=g)
    Wt = W.t().contiguous()  # [K, N]
    gout = ops.PlaneBuf(B, N, 2, dev)
    ops.skinny_expand(ga, Wt, 1, N, K=K, N=N, act=ops.ACT_MASK, mask=outp, out_planes=gout)
    refm = (ga.double() @ Wt.double()) * (out32 > 0)
    assert ((gout.to_float().double() - refm).abs().max() / refm.abs().max()).item() < 2e-5


@pytest.mark.parametrize("B,S,Wd", [(4096, 8, 400), (4096, 12, 400), (8229, 60, 400), (333, 3, 50)])
def test_skinny_wgrad(dev, B, S, Wd):
    """Weight / bias gradients of the skinny layers: batch reductions with the bias riding along."""
    from mvae_b200 import ops
    g = torch.Generator(device=dev).manual_seed(B + S * 7 + Wd)
    small = torch.randn(B, S, device=dev, generator=g)
    wide = torch.randn(B, Wd, device=dev, generator=g)
    # (a) fc_d0 style: out[w, s], bias of the wide side from the implicit ones
    out = torch.zeros(Wd, S, device=dev)
    brow = torch.zeros(Wd, device=dev)
    ops.skinny_wgrad(small, S, (_planes_of(wide, dev), 2), Wd, out, 1, S, small_ones=True, out_row=brow)
    ref = wide.double().t() @ small.double()
    assert ((out.double() - ref).abs().max() / ref.abs().max()).item() < 3e-5
    assert ((brow.double() - wide.double().sum(0)).abs().max() / wide.double().sum(0).abs().max()).item() < 3e-5
    # (b) heads style: out[s, w], bias of the small side from the wide operand's ones column
    wp = _planes_of(wide, dev, planes=3, ones_col=True)
    out2 = torch.zeros(S, Wd, device=dev)
    bcol = torch.zeros(S, device=dev)
    ops.skinny_wgrad(small, S, (wp, 2), Wd + 1, out2, Wd, 1, out_col=bcol, col_split=Wd)
    ref2 = small.double().t() @ wide.double()
    assert ((out2.double() - ref2).abs().max() / ref2.abs().max()).item() < 3e-5
    assert ((bcol.double() - small.double().sum(0)).abs().max() / small.double().sum(0).abs().max()).item() < 1e-5
    # fp32 wide operand: exact fp32 accumulation
    out3 = torch.zeros(S, Wd, device=dev)
    ops.skinny_wgrad(small, S, wide, Wd, out3, Wd, 1)
    assert ((out3.double() - ref2).abs().max() / ref2.abs().max()).item() < 2e-6


# ------------------------------------------------------------------------------------------ fused latent block
@pytest.mark.parametrize("sig,B,H,scalar", [("h2,s2,e2", 4096, 400, False), ("h6,h6,s6,s6,e6", 1003, 400, False),
                                            ("p2", 17, 64, False), ("h3,s5,p4,e1", 300, 136, True),
                                            ("e2", 128, 400, False)])
def test_latent_fused_matches_separate_kernels(dev, oracle, sig, B, H, scalar):
    """mvae_latent_forward / _backward (heads + manifold chain + fc_d0 in one launch per direction) against the float64
    oracle for the manifold part and float64 matmuls for the dense parts; ragged batches, scalar parametrization."""
    from mvae_b200 import ops
    desc = ops.make_desc(sig, scalar_parametrization=scalar)
    odesc = oracle.make_desc(sig, scalar_parametrization=scalar)
    P, Sn, Sd, C = desc.ld_ml, desc.ld_eps, desc.ld_z, desc.C
    g = torch.Generator(device=dev).manual_seed(B + H)
    h = torch.randn(B, H, device=dev, generator=g).clamp(min=0)           # relu output: about half the units are off
    Wh = torch.randn(P, H, device=dev, generator=g) / H**0.5
    bh = torch.randn(P, device=dev, generator=g) * 0.1
    Wd0 = torch.randn(H, Sd, device=dev, generator=g) / Sd**0.5
    bd0 = torch.randn(H, device=dev, generator=g) * 0.1
    eps = torch.randn(B, Sn, device=dev, generator=g)
    R = torch.full((C,), 1.7, device=dev)
    ml = torch.full((B, P), float("nan"), device=dev)
    z = torch.full((B, Sd), float("nan"), device=dev)
    kl = torch.full((B, C), float("nan"), device=dev)
    ddp = ops.PlaneBuf(B, H, 2, dev, ones_col=True)
    ops.latent_forward(desc, h, Wh, bh, eps, R, Wd0, bd0, ml, z, kl, ddp)
    ml_ref = h.double() @ Wh.double().t() + bh.double()
    assert ((ml.double() - ml_ref).abs().max() / ml_ref.abs().max()).item() < 2e-6
    ref = oracle.pm_forward(odesc, ml.double().cpu().numpy(), eps.double().cpu().numpy(), R.double().cpu().numpy())
    assert normwise(z.cpu().numpy(), ref["z"]) < FWD_TOL
    err = np.abs(kl.cpu().numpy() - ref["kl"]) / np.maximum(1.0, np.abs(ref["kl"]))
    assert np.quantile(err, 0.99) < 5e-5
    dd_ref = (z.double() @ Wd0.double().t() + bd0.double()).clamp(min=0)
    assert ((ddp.to_float().double() - dd_ref).abs().max() / dd_ref.abs().max()).item() < 2e-5
    assert torch.equal(ddp.t[0, :, H].float(), torch.ones(B, device=dev))  # ones column untouched
    # backward
    gdd = torch.randn(B, H, device=dev, generator=g) * (dd_ref > 0)
    gdd = gdd.float().contiguous()
    ghp = ops.PlaneBuf(B, H, 2, dev)
    gWd0, gbd0 = torch.zeros(H, Sd, device=dev), torch.zeros(H, device=dev)
    gWh, gbh, gR = torch.zeros(P, H, device=dev), torch.zeros(P, device=dev), torch.zeros(C, device=dev)
    ops.latent_backward(desc, gdd, h, Wh, Wd0, ml, eps, R, z, 0.7, ghp, gWd0, gbd0, gWh, gbh, gR)
    gdd64 = gdd.double()
    gz_ref = (gdd64 @ Wd0.double()).cpu().numpy()
    gml_ref, gR_ref = oracle.pm_backward(odesc, ml.double().cpu().numpy(), eps.double().cpu().numpy(),
                                         R.double().cpu().numpy(), gz_ref, None, 0.7)
    gml64 = torch.from_numpy(gml_ref).to(dev)
    gh_ref = (gml64 @ Wh.double()) * (h > 0)
    assert ((ghp.to_float().double() - gh_ref).abs().max() / gh_ref.abs().max()).item() < 1e-4
    for got, want in ((gWd0, gdd64.t() @ z.double()), (gbd0, gdd64.sum(0)), (gWh, gml64.t() @ h.double()),
                      (gbh, gml64.sum(0))):
        assert ((got.double() - want).norm() / want.norm()).item() < 1e-4
    assert np.all(np.abs(gR.cpu().numpy() - gR_ref) <= 2e-4 * np.maximum(1.0, np.abs(gml_ref).sum()))


@pytest.mark.parametrize("sig,B,H", [("h2,s2,e2", 4096, 400), ("h2", 1003, 64), ("p2,e3", 17, 136)])
def test_latent_forward_takes_the_step_prologue_along(dev, sig, B, H):
    """mvae_latent_forward_ex: the noise the kernel draws itself is, bit for bit, what mvae_step_prologue draws for the
    same seed / device counter (so the fused step and the two-launch step sample the same eps), the spans handed to
    it are zero afterwards, and every output equals the plain launch fed with that noise."""
    from mvae_b200 import ops
    desc = ops.make_desc(sig)
    P, Sn, Sd, C = desc.ld_ml, desc.ld_eps, desc.ld_z, desc.C
    g = torch.Generator(device=dev).manual_seed(7 * B + H)
    h = torch.randn(B, H, device=dev, generator=g).clamp(min=0)
    Wh = torch.randn(P, H, device=dev, generator=g) / H**0.5
    bh = torch.randn(P, device=dev, generator=g) * 0.1
    Wd0 = torch.randn(H, Sd, device=dev, generator=g) / Sd**0.5
    bd0 = torch.randn(H, device=dev, generator=g) * 0.1
    R = torch.full((C,), 2.5, device=dev)
    seed, ctr = 0x5EED1234ABCD, torch.full((1,), 41, device=dev, dtype=torch.int64)
    eps_ref = torch.full((B, Sn), float("nan"), device=dev)
    ops.step_prologue(eps_ref, seed, ctr, [])
    assert torch.isfinite(eps_ref).all()

    def run(eps, **kw):
        out = [torch.full((B, P), float("nan"), device=dev), torch.full((B, Sd), float("nan"), device=dev),
               torch.full((B, C), float("nan"), device=dev), ops.PlaneBuf(B, H, 2, dev)]
        ops.latent_forward(desc, h, Wh, bh, eps, R, Wd0, bd0, out[0], out[1], out[2], out[3], **kw)
        return out

    plain = run(eps_ref)
    eps_drawn = torch.full((B, Sn), float("nan"), device=dev)
    z1 = torch.full((1001,), 3.0, device=dev)
    z2 = torch.full((B,), -1.0, device=dev)
    z3 = torch.full((7,), 5.0, device=dev)[1:]   # not 16-byte aligned
    fused = run(eps_drawn, draw=(seed, ctr), zero=[z1, z2, z3])
    assert torch.equal(eps_drawn, eps_ref)
    assert not z1.any() and not z2.any() and not z3.any()
    for a, b in zip(plain[:3], fused[:3]):
        assert torch.equal(a, b)
    assert torch.equal(plain[3].t, fused[3].t)
    # zero fills alone (noise supplied)
    z1.fill_(2.0)
    only_zero = run(eps_ref, zero=[z1])
    assert not z1.any() and torch.equal(only_zero[1], plain[1])
