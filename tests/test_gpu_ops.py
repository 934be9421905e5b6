"""GPU parity tests of the standalone `Manifold` / `WrappedNormal` operator API (exp / log maps,
parallel transport, geodesic distance, Moebius addition, model conversions, log-dets, rsample / log_prob) through the
host mirror classes (mvae_b200/manifolds.py, distributions.py -> mvae_manifold_op / mvae_wn_*), against the golden
vectors the reference's own ops produced in float64 (tests/golden/ops_<letter>.npz, generate_golden.py:gen_ops)."""
import numpy as np
import pytest

from helpers import load_golden, normwise

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-4      # BASELINE.json north_star: 1e-4 relative (normwise per tensor)
TOL_INV = 3e-4  # inverse maps / log-dets restated literally (acos / acosh of a dot product: float32 cancellation)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mvae_b200 import _lib
    _lib.lib()
    return torch.device("cuda:0")


def _manifold(letter, R, dev):
    from mvae_b200 import manifolds as M
    Rt = torch.tensor(float(R), device=dev)
    if letter == "e":
        return M.Euclidean()
    cls = {"h": M.Hyperboloid, "s": M.Sphere, "p": M.PoincareBall, "d": M.StereographicallyProjectedSphere}[letter]
    return cls(lambda: Rt)


@pytest.mark.parametrize("letter", ["h", "s", "p", "d", "e"])
def test_manifold_ops_match_reference(dev, letter):
    g, _ = load_golden("ops_" + letter)
    t = lambda k: torch.from_numpy(np.ascontiguousarray(g[k], dtype=np.float32)).to(dev)
    close = lambda got, key, tol=TOL: normwise(got.cpu().numpy(), g[key]) < tol
    man = _manifold(letter, float(g["R"]), dev)
    x, y, v = t("x"), t("y"), t("v")
    assert close(man.exp_map_mu0(t("t1")), "x")
    assert close(man.exp_map_mu0(t("t2")), "y")
    assert close(man.inverse_exp_map_mu0(x), "inv_exp_map_mu0_x", TOL_INV)
    z, (u, vv) = man.sample_projection_mu0(v, x)
    assert close(z, "sp_z") and close(u, "sp_u") and close(vv, "sp_v")
    # bit-exact layout work: v is passed through untouched
    assert torch.equal(vv, v)
    if letter != "e":
        iu, iv = man.inverse_sample_projection_mu0(t("sp_z"), x)
        assert close(iu, "isp_u", TOL_INV) and close(iv, "isp_v", TOL_INV)
        # round trip: the tangent vector that was projected comes back
        assert normwise(iv.cpu().numpy(), g["v"]) < TOL_INV
    tangent = t("tangent_mu0") if letter in "hs" else v
    pt = man.parallel_transport_mu0(tangent, x)
    assert close(pt, "pt_mu0")
    assert close(man.inverse_parallel_transport_mu0(t("pt_mu0"), x), "inv_pt_mu0")
    assert close(man.exp_map(t("pt_mu0"), x), "exp_map")
    assert close(man.inverse_exp_map(y, x), "inv_exp_map", TOL_INV)
    d_xy = man.distance(x, y)
    assert d_xy.shape == (x.shape[0], 1)
    assert close(d_xy, "distance", TOL_INV)
    # symmetry and identity of indiscernibles (the reference's own property tests, tests/mvae/ops/test_poincare.py:74-79)
    assert normwise(man.distance(y, x).cpu().numpy(), g["distance"]) < TOL_INV
    if letter in "hs":
        ld = man.logdet(None, None, None, (t("sp_u"),))
        assert close(ld, "logdet_u", TOL_INV)
        proj = man.to_poincare(x) if letter == "h" else man.to_projected(x)
        assert close(proj, "to_poincare")
    if letter in "pd":
        assert close(man.mobius_add(x, y), "mobius_add")
        lifted = man.to_lorentz(x) if letter == "p" else man.to_spherical(x)
        assert close(lifted, "from_poincare")
        ld = man.logdet(x, t("scale"), t("sp_z"), (t("sp_u"), v))
        assert close(ld, "logdet_zmu", TOL_INV)


@pytest.mark.parametrize("letter", ["h", "s", "p", "d"])
def test_wrapped_normal_matches_reference(dev, letter):
    from mvae_b200.distributions import WrappedNormal
    g, _ = load_golden("ops_" + letter)
    t = lambda k: torch.from_numpy(np.ascontiguousarray(g[k], dtype=np.float32)).to(dev)
    man = _manifold(letter, float(g["R"]), dev)
    x, scale = t("x"), t("scale")
    q = WrappedNormal(x, scale, man)
    eps = (t("v") / scale).contiguous()
    z, (u, v) = q.rsample_with_parts(eps=eps)
    assert normwise(z.cpu().numpy(), g["wn_z"]) < TOL
    assert normwise(u.cpu().numpy(), g["wn_u"]) < TOL
    assert normwise(v.cpu().numpy(), g["wn_v"]) < TOL
    lq = q.log_prob_from_parts(t("wn_z"), (t("wn_u"), t("wn_v")))
    assert normwise(lq.cpu().numpy(), g["wn_logq_parts"]) < TOL_INV
    assert normwise(q.log_prob(t("wn_z")).cpu().numpy(), g["wn_logq"]) < TOL_INV
    p = WrappedNormal(man.mu_0(x.shape, device=dev), torch.ones_like(scale), man)
    assert normwise(p.log_prob(t("wn_z")).cpu().numpy(), g["wn_logp"]) < TOL_INV
    # n samples per row (the IWAE path's shape): [n, B, d]
    zz, (uu, vv) = q.rsample_with_parts(torch.Size([3]))
    assert zz.shape == (3,) + tuple(x.shape) and vv.shape == (3,) + tuple(scale.shape)
    assert torch.isfinite(zz).all() and torch.isfinite(q.log_prob_from_parts(zz, (uu, vv))).all()


def test_component_forward_matches_fused_kernel(dev):
    """Component.forward (component.py:63-75 + reparametrize) through the standalone ops reproduces what the fused
    product-manifold kernel computes for the same head pre-activations — the two implementations of the chain
    (op-by-op restatement vs closed forms) agree, for every component letter incl. 'd'."""
    from mvae_b200 import components, ops
    torch.manual_seed(0)
    comps = components.parse_components("h3,s2,d3,p2,e2", False)
    H, B = 24, 64
    for i, c in enumerate(comps):
        c.init_layers(H, False)
        c.to(dev)
        _, rp = c.radius_parameter()
        if rp is not None:
            rp.data.fill_(1.0 + 0.3 * i)
    h = torch.randn(B, H, device=dev) * 0.5
    desc = ops.make_desc([c.kind for c in comps], [c.true_dim for c in comps])
    ml = torch.cat([torch.cat((torch.nn.functional.linear(h, c.fc_mean.weight, c.fc_mean.bias),
                               torch.nn.functional.linear(h, c.fc_logvar.weight, c.fc_logvar.bias)), -1)
                    for c in comps], -1).detach().contiguous()
    eps = torch.randn(B, desc.ld_eps, device=dev)
    R = torch.stack([c.radius_parameter()[1].detach() if c.radius_parameter()[1] is not None
                     else torch.tensor(1.0, device=dev) for c in comps]).float().contiguous()
    fused = ops.pm_forward(desc, ml, eps, R, want_mu_sigma=True)
    for i, c in enumerate(comps):
        d = desc.comp[i]
        with torch.no_grad():
            q_z, p_z, _ = c(h)
        assert normwise(q_z.loc.cpu().numpy(), fused["mu"][:, d.z_off:d.z_off + d.d].cpu().numpy()) < TOL
        assert normwise(q_z.scale.cpu().numpy(), fused["sigma"][:, d.eps_off:d.eps_off + d.n].cpu().numpy()) < TOL
        if c.letter == "e":
            continue
        z, data = q_z.rsample_with_parts(eps=eps[:, d.eps_off:d.eps_off + d.n].contiguous())
        assert normwise(z.cpu().numpy(), fused["z"][:, d.z_off:d.z_off + d.d].cpu().numpy()) < TOL
        kl = c.kl_loss(q_z, p_z, z, data)
        assert normwise(kl.cpu().numpy(), fused["kl"][:, i].cpu().numpy()) < TOL_INV
