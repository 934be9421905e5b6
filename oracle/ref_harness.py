"""Reference harness — TEST INFRASTRUCTURE ONLY (never imported by mvae_b200/).

Imports the unmodified reference from /root/reference (exists only in the build container) through
the shims in oracle/ref_shims and exposes helpers that run the reference's own hot
path with injected noise so golden vectors can be generated (tests/golden/generate_golden.py) and the
C oracle can be pinned against them.

Reference entry points exercised (all paths relative to /root/reference):
  mt/mvae/components/component.py:63-75      Component.encode
  mt/mvae/sampling/sampling_procedures.py    WrappedNormalProcedure / EuclideanNormalProcedure
  mt/mvae/distributions/wrapped_normal.py    rsample_with_parts / log_prob_from_parts / log_prob
  mt/mvae/models/vae.py:69-80,125-166        ModelVAE.forward / compute_batch_stats / train_step
"""
import contextlib
import os
import sys

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("MVAE_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mt", "mvae"))


def load_reference():
    """Put the shims and the reference on sys.path; return the ``mt`` package."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    # torch 1.1 did not validate distribution args; WrappedNormal.__init__ sets loc after super().__init__().
    torch.distributions.Distribution.set_default_validate_args(False)
    import mt  # noqa: F401
    import mt.mvae.utils  # noqa: F401
    return mt


@contextlib.contextmanager
def default_dtype(dtype):
    """The reference sets the global default dtype once (mt/examples/run.py:98-101) and allocates helper tensors
    (zeros for the tangent Normal, mu_0, ...) in it; run every reference call under the intended dtype."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        yield
    finally:
        torch.set_default_dtype(old)


@contextlib.contextmanager
def injected_noise(eps_list):
    """Make ``Normal.rsample`` consume the given standard-normal draws in order
    (reference draws once per component in component order: wrapped_normal.py:72, vae.py:75)."""
    import torch.distributions.normal as tdn
    queue = list(eps_list)
    orig = tdn._standard_normal

    def fake(shape, dtype, device):
        e = queue.pop(0)
        assert tuple(e.shape) == tuple(shape), (tuple(e.shape), tuple(shape))
        return e.to(dtype=dtype, device=device)

    tdn._standard_normal = fake
    try:
        yield
    finally:
        tdn._standard_normal = orig


class _Dataset:
    """Minimal VaeDataset stand-ins carrying only the reconstruction losses of
    mt/data/image_reconstruction.py:81-82 (BCE with logits) and mt/data/synthetic.py:161-162 (unit Gaussian NLL)."""

    def __init__(self, mt, kind, in_dim):
        from mt.data import VaeDataset
        import torch.nn.functional as F
        from torch.distributions import Normal

        class BCE(VaeDataset):

            def __init__(self):
                super().__init__(batch_size=1, in_dim=in_dim, img_dims=None)

            def reconstruction_loss(self, x_mb_, x_mb):
                return F.binary_cross_entropy_with_logits(x_mb_, x_mb, reduction="none")

        class NLL(VaeDataset):

            def __init__(self):
                super().__init__(batch_size=1, in_dim=in_dim, img_dims=None)

            def reconstruction_loss(self, x_mb_, x_mb):
                return -Normal(x_mb_, torch.ones_like(x_mb_)).log_prob(x_mb)

        self.ds = BCE() if kind == "bce" else NLL()


def build_model(model_sig, in_dim, h_dim, fixed_curvature, scalar_parametrization, recon, seed, dtype,
                architecture="ff"):
    """architecture "ff": FeedForwardVAE (ffnn_vae.py:27-60); "conv": ConvolutionalVAE (conv_vae.py:28-79; in_dim 3072,
    h_dim 8192 are fixed by its layers)."""
    mt = load_reference()
    from mt.mvae import utils
    from mt.mvae.models import ConvolutionalVAE, FeedForwardVAE
    with default_dtype(dtype):
        torch.manual_seed(seed)
        comps = utils.parse_components(model_sig, fixed_curvature)
        cls = ConvolutionalVAE if architecture == "conv" else FeedForwardVAE
        model = cls(h_dim, comps, _Dataset(mt, recon, in_dim).ds, scalar_parametrization)
        return model.to(torch.device("cpu")).to(dtype)


def comp_type_letter(component) -> str:
    return type(component).__name__.lower()[0] if "Stereo" not in type(component).__name__ else "d"


def draw_eps(model, B, seed, dtype, n_samples=None):
    g = torch.Generator().manual_seed(seed)
    out = []
    for c in model.components:
        shape = (B, c.true_dim) if n_samples is None else (n_samples, B, c.true_dim)
        out.append(torch.randn(shape, generator=g, dtype=torch.float64).to(dtype))
    return out


def ref_model_step(model, x, eps_list, beta):
    """One forward + ELBO + backward of the reference model (no optimizer step).
    Returns a dict of numpy arrays: intermediates, per-sample stats, and all parameter gradients."""
    with default_dtype(x.dtype):
        return _ref_model_step(model, x, eps_list, beta)


def _ref_model_step(model, x, eps_list, beta):
    model.zero_grad()
    h = model.encode(x)
    out = {"x": x, "h": h}
    ms, ls = [], []
    for c in model.components:
        ms.append(c.fc_mean(h))
        ls.append(c.fc_logvar(h))
    with injected_noise(eps_list):
        reparametrized, concat_z, x_ = model(x)
    stats = model.compute_batch_stats(x, x_, reparametrized, beta=beta, likelihood_n=0)
    loss = -stats.elbo
    loss.backward()
    out.update({
        "m": torch.cat(ms, -1),
        "l": torch.cat(ls, -1),
        "eps": torch.cat(eps_list, -1),
        "mu": torch.cat([r.q_z.loc for r in reparametrized], -1),
        "sigma": torch.cat([r.q_z.scale for r in reparametrized], -1),
        "z": concat_z,
        "logits": x_,
        "bce": stats._bce,
        "kl": torch.stack(stats._component_kl, -1),
        "elbo": stats.elbo,
        "bce_sum": stats.bce,
        "kl_sum": stats.kl,
    })
    res = {k: v.detach().cpu().numpy() for k, v in out.items()}
    for name, p in model.named_parameters():
        res["param." + name] = p.detach().cpu().numpy()
        res["grad." + name] = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().cpu().numpy()
    return res


def ref_log_likelihood(model, x, eps_list, n):
    """The reference's own ModelVAE.log_likelihood (vae.py:82-123) with the draws injected:
    eps_list[i] is [n, B, n_i] for component i (one Normal.rsample(sample_shape) per component, in order)."""
    with default_dtype(x.dtype), torch.no_grad(), injected_noise(eps_list):
        ll, mi, cov = model.log_likelihood(x, n=n)
    return {"log_p_x": ll.detach().cpu().numpy(), "mi": mi.detach().cpu().numpy(),
            "cov_norm": np.asarray(cov.detach().cpu().numpy())}


def ref_product_manifold(model_sig, m, l, eps, radii, gz, gkl, scalar_parametrization=False):
    """Reference L0-L2 path for a product manifold from head pre-activations.

    m [B, sum n], l [B, sum n | C], eps [B, sum n] torch tensors (leaves are created here);
    radii: list of floats (one per component; ignored for 'e').
    Returns forward outputs (mu, sigma, z, u, kl, logq, logp) and gradients of
    L = sum(gz*z) + sum(gkl*kl) with respect to m, l and every radius parameter."""
    load_reference()
    with default_dtype(m.dtype):
        return _ref_product_manifold(model_sig, m, l, eps, radii, gz, gkl, scalar_parametrization)


def _ref_product_manifold(model_sig, m, l, eps, radii, gz, gkl, scalar_parametrization):
    from mt.mvae import utils
    dtype = m.dtype
    comps = utils.parse_components(model_sig, fixed_curvature=False)
    for c in comps:
        c.init_layers(4, scalar_parametrization)
    m = m.clone().requires_grad_(True)
    l = l.clone().requires_grad_(True)
    Rs = []
    mo = lo = 0
    mus, sigmas, zs, us, kls, logqs, logps = [], [], [], [], [], [], []
    import torch.nn.functional as F
    eps_list = []
    for i, c in enumerate(comps):
        n = c.true_dim
        ln = 1 if scalar_parametrization else n
        for pname in ("_nradius", "_pradius", "_curvature"):  # 'u': radii[i] is the curvature
            if hasattr(c, pname):
                getattr(c, pname).data = torch.tensor(float(radii[i]), dtype=dtype)
                Rs.append(getattr(c, pname))
        if not (hasattr(c, "_nradius") or hasattr(c, "_pradius") or hasattr(c, "_curvature")):
            Rs.append(None)
        mi = m[:, mo:mo + n]
        li = l[:, lo:lo + ln]
        ei = eps[:, mo:mo + n]
        mo += n
        lo += ln
        # component.py:63-75 with the Linear layers factored out
        mu = c.manifold.exp_map_mu0(mi)
        std = F.softplus(li) + 1e-5
        q_z, p_z = c.reparametrize(mu, std)
        with injected_noise([ei]):
            z, data = q_z.rsample_with_parts()
        kl = c.kl_loss(q_z, p_z, z, data)
        mus.append(mu)
        sigmas.append(q_z.scale if hasattr(q_z, "scale") else std)
        zs.append(z)
        kls.append(kl)
        if data is not None:
            us.append(data[0])
            logqs.append(q_z.log_prob_from_parts(z, data))
            logps.append(p_z.log_prob(z))
        else:
            us.append(torch.zeros_like(z))
            logqs.append(q_z.log_prob(z))
            logps.append(p_z.log_prob(z))
    z = torch.cat(zs, -1)
    kl = torch.stack(kls, -1)
    L = (gz * z).sum() + (gkl * kl).sum()
    L.backward()
    gR = [float(r.grad) if (r is not None and r.grad is not None) else 0.0 for r in Rs]
    res = {
        "mu": torch.cat(mus, -1),
        "sigma": torch.cat(sigmas, -1),
        "z": z,
        "u": torch.cat(us, -1),
        "kl": kl,
        "logq": torch.stack(logqs, -1),
        "logp": torch.stack(logps, -1),
        "gm": m.grad,
        "gl": l.grad,
    }
    res = {k: v.detach().cpu().numpy() for k, v in res.items()}
    res["gR"] = np.asarray(gR, dtype=np.float64)
    return res
