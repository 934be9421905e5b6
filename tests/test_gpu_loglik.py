"""GPU parity tests of FusedFeedForwardVAE.log_likelihood (the IWAE estimate of vae.py:82-123) and its kernels
(mvae_iwae_latent / mvae_iwae_reduce / mvae_iwae_cov_norm, the logits GEMM with targets indexed modulo B) against the
golden vectors the reference itself produced (tests/golden/loglik_*.npz) and against the CPU oracle at larger shapes."""
import numpy as np
import pytest

from helpers import load_golden, loglik_golden_names, normwise

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-4  # BASELINE.json north_star: 1e-4 relative (normwise per tensor)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mvae_b200 import _lib
    _lib.lib()
    return torch.device("cuda:0")


def _build(meta, params, dev):
    from mvae_b200 import components, data, vae
    comps = components.parse_components(meta["sig"], False)
    ds = data.GenericDataset(1, meta["in_dim"], meta["recon"])
    model = vae.FusedFeedForwardVAE(meta["h_dim"], comps, ds, meta["scalar_parametrization"], device=dev)
    sd = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in params.items()}
    model.load_state_dict(sd, strict=True)
    return model


@pytest.mark.parametrize("name", loglik_golden_names())
def test_log_likelihood_matches_reference(dev, name):
    g, meta = load_golden(name)
    params = {k[len("param."):]: v for k, v in g.items() if k.startswith("param.")}
    model = _build(meta, params, dev)
    x = torch.from_numpy(g["x"].astype(np.float32)).to(dev)
    eps = torch.from_numpy(g["eps"].astype(np.float32)).to(dev)
    ll, mi, cov = model.log_likelihood(x, n=meta["n"], eps=eps)
    assert ll.shape == mi.shape == (x.shape[0],) and cov.dim() == 0
    assert normwise(ll.cpu().numpy(), g["log_p_x"]) < TOL
    assert normwise(mi.cpu().numpy(), g["mi"]) < TOL
    assert abs(float(cov) - float(g["cov_norm"])) < TOL * float(g["cov_norm"])
    # through compute_batch_stats (vae.py:142-145): batch sums, as BatchStats reports them (stats.py:177-186)
    rep, _, logits = model.forward(x, eps=eps[0])
    torch.manual_seed(0)
    stats = model.compute_batch_stats(x, logits, rep, beta=1.0, likelihood_n=3).convert_to_float()
    assert np.isfinite(stats.log_likelihood) and np.isfinite(stats.mutual_info) and stats.cov_norm >= 0
    assert set(stats.to_print()) == {"bce", "kl", "elbo", "ll", "mi", "cov_norm", "beta"}


@pytest.mark.parametrize("sig,B,D,H,recon,n,chunk", [
    ("h2,s2,e2", 512, 784, 400, "bce", 12, 1 << 17),     # cfg2 model: one chunk
    ("h2,s2,e2", 300, 784, 400, "bce", 11, 1024),        # ragged rows, 3 samples per chunk + ragged last chunk
    ("h6,h6,s6,s6,e6", 256, 784, 400, "bce", 6, 512),    # cfg3 model
    ("p2", 257, 50, 400, "nll", 9, 1 << 17),             # cfg4 (BDP-shaped), Gaussian NLL
    ("h2,s3,p4,e5", 64, 100, 64, "bce", 5, 1 << 17),     # mixed dimensions -> wide instantiations
])
def test_log_likelihood_vs_oracle(dev, oracle, sig, B, D, H, recon, n, chunk):
    from mvae_b200 import components, data, vae
    torch.manual_seed(0)
    model = vae.FusedFeedForwardVAE(H, components.parse_components(sig, False), data.GenericDataset(B, D, recon),
                                    False, device=dev)
    model.iwae_chunk_rows = chunk
    with torch.no_grad():
        for i, c in enumerate(model.components):
            _, rp = c.radius_parameter()
            if rp is not None:
                rp.fill_(0.8 + 0.5 * i)
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(B, D, generator=g) < 0.1307).float() if recon == "bce" else torch.randn(B, D, generator=g)
    eps = torch.randn(n, B, model.desc.ld_eps, generator=g)
    params = {k: v.detach().cpu().double().numpy() for k, v in model.state_dict().items()}
    ll, mi, cov = model.log_likelihood(x.to(dev), n=n, eps=eps.to(dev))
    ref = oracle.OracleVAE(sig, D, H, recon, False).log_likelihood(params, x.double().numpy(), eps.double().numpy())
    assert normwise(ll.cpu().numpy(), ref["log_p_x"]) < TOL
    assert normwise(mi.cpu().numpy(), ref["mi"]) < TOL
    assert abs(float(cov) - ref["cov_norm"]) < TOL * ref["cov_norm"]


def test_iwae_kernels_vs_oracle(dev, oracle):
    """The three kernels on their own: z / log q - log p per (sample, row) incl. the Monte-Carlo Euclidean term,
    sum_s z, the streaming logsumexp against numpy, and bit-exact layout (sample s of the [ns, B, .] outputs equals a
    single-sample launch on eps[s])."""
    from mvae_b200 import ops
    sig, B, ns = "h2,s2,p3,e2,e5", 77, 6
    desc, odesc = ops.make_desc(sig), oracle.make_desc(sig)
    g = torch.Generator().manual_seed(3)
    ml = torch.randn(B, desc.ld_ml, generator=g)
    eps = torch.randn(ns, B, desc.ld_eps, generator=g)
    R = torch.tensor([1.2, 0.9, 2.0, 1.0, 1.0])
    z = torch.empty(ns, B, desc.ld_z, device=dev)
    diff = torch.empty(ns, B, device=dev)
    zsum = torch.zeros(B, desc.ld_z, device=dev)
    ops.iwae_latent(desc, ml.to(dev), eps.to(dev), R.to(dev), z, diff, zsum)
    for s in range(ns):
        f = oracle.pm_forward(odesc, ml.double().numpy(), eps[s].double().numpy(), R.double().numpy(),
                              want=("z", "logq", "logp"))
        assert normwise(z[s].cpu().numpy(), f["z"]) < TOL
        assert normwise(diff[s].cpu().numpy(), (f["logq"] - f["logp"]).sum(-1)) < TOL
        z1 = torch.empty(1, B, desc.ld_z, device=dev)
        d1 = torch.empty(1, B, device=dev)
        ops.iwae_latent(desc, ml.to(dev), eps[s:s + 1].to(dev).contiguous(), R.to(dev), z1, d1, None)
        assert torch.equal(z1[0], z[s]) and torch.equal(d1[0], diff[s])
    assert normwise(zsum.cpu().numpy(), z.sum(0).cpu().double().numpy()) < 1e-6
    # wrapped-normal components: log q - log p is the training kernel's KL term (same device code)
    kl = ops.pm_forward(desc, ml.to(dev), eps[0].to(dev).contiguous(), R.to(dev))["kl"]
    f0 = oracle.pm_forward(odesc, ml.double().numpy(), eps[0].double().numpy(), R.double().numpy(), want=("logq", "logp"))
    assert normwise(kl[:, :3].cpu().numpy(), (f0["logq"] - f0["logp"])[:, :3]) < TOL
    # streaming logsumexp (values spread over a wide range; n not a multiple of the 8 sample groups)
    n = 37
    recon = (torch.rand(n, B, generator=g) * 300 + 50).to(dev)
    d = (torch.randn(n, B, generator=g) * 20).to(dev)
    ll, mi = ops.iwae_reduce(recon, d)
    ref_ll = torch.logsumexp((-recon - d).double(), 0) - np.log(n)
    ref_mi = torch.logsumexp(d.double(), 0) - np.log(n)
    assert normwise(ll.cpu().numpy(), ref_ll.cpu().numpy()) < 1e-6
    assert normwise(mi.cpu().numpy(), ref_mi.cpu().numpy()) < 1e-6
    # cov_norm against float64 numpy (vae.py:119-121 literally)
    D = 300
    x = (torch.rand(B, D, generator=g) < 0.2).float()
    zs = z.cpu().double().numpy()
    xc = x.double().numpy() - x.double().numpy().mean(0, keepdims=True)
    cov = np.einsum("bd,sbj->sdj", xc, zs - zs.mean(1, keepdims=True)).mean(0)
    got = ops.iwae_cov_norm(x.to(dev), zsum, ns)
    assert abs(float(got) - np.linalg.norm(cov)) < 1e-5 * np.linalg.norm(cov)


def test_gemm_targets_modulo_rows(dev):
    """BCE epilogue with aux_rows: row m of the [n*B, H] activations is scored against x[m % B] — the same as
    n separate GEMMs of B rows each."""
    from mvae_b200 import _lib as L
    from mvae_b200 import ops
    B, n, H, D = 200, 3, 64, 96
    g = torch.Generator().manual_seed(5)
    a = torch.randn(n * B, H, generator=g).to(dev)
    W = (torch.randn(D, H, generator=g) * 0.2).to(dev)
    bias = torch.randn(D, generator=g).to(dev)
    x = (torch.rand(B, D, generator=g) < 0.3).float().to(dev)
    ap, Wp = ops.PlaneBuf(n * B, H, 2, dev), ops.PlaneBuf(D, H, 2, dev)
    ops.split_planes(a, ap)
    ops.split_planes(W, Wp)
    rs = torch.zeros(n * B, device=dev)
    ops.gemm(ap, Wp, n * B, D, H, epilogue=L.EPI_BCE_ROWSUM, bias=bias, aux=x, aux_rows=B, rowsum=rs)
    for s in range(n):
        ap1 = ops.PlaneBuf(B, H, 2, dev)
        ops.split_planes(a[s * B:(s + 1) * B].contiguous(), ap1)
        rs1 = torch.zeros(B, device=dev)
        ops.gemm(ap1, Wp, B, D, H, epilogue=L.EPI_BCE_ROWSUM, bias=bias, aux=x, rowsum=rs1)
        assert normwise(rs1.cpu().numpy(), rs[s * B:(s + 1) * B].cpu().numpy()) < 1e-6
    logits = a.double() @ W.double().t() + bias.double()
    ref = torch.nn.functional.binary_cross_entropy_with_logits(logits, x.double().repeat(n, 1), reduction="none").sum(-1)
    assert normwise(rs.cpu().numpy(), ref.cpu().numpy()) < TOL
