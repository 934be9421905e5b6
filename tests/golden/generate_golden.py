"""Generate the golden fixtures under tests/golden/ by running the REFERENCE ITSELF (unmodified, from
/root/reference, through oracle/ref_shims) in this container.  The fixtures are what pins the C oracle
(tests/test_oracle_golden.py) and, through it, the CUDA path.  Re-run:  python tests/golden/generate_golden.py

Fixtures (all small, float64 unless the key ends in _f32):
  pm_<name>.npz     product-manifold path from head pre-activations (component.py:63-75 ... sampling_procedures.py:101-116)
                    inputs m,l,eps,radii,gz,gkl; outputs mu,sigma,z,u,kl,logq,logp and autograd gm,gl,gR;
                    plus the reference's own float32 run of the same inputs (suffix _f32)
  model_<name>.npz  whole ModelVAE.forward + compute_batch_stats + backward (vae.py:69-80,125-160) on a tiny FeedForwardVAE
  ops_<letter>.npz  standalone Manifold ops (ops/*.py) for h,s,p,e at R=2
  loglik_<name>.npz ModelVAE.log_likelihood (vae.py:82-123: IWAE estimate, mutual information, cov_norm) of a tiny
                    FeedForwardVAE with the n x B draws injected
  kat.json          known-answer vectors re-derived here
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref_harness as rh  # noqa: E402

rh.load_reference()


def _inputs(sig_dims, B, seed, scale_m, scalar):
    g = torch.Generator().manual_seed(seed)
    sn = sum(sig_dims)
    C = len(sig_dims)
    sl = C if scalar else sn
    m = torch.randn(B, sn, generator=g, dtype=torch.float64) * scale_m
    l = torch.randn(B, sl, generator=g, dtype=torch.float64)
    eps = torch.randn(B, sn, generator=g, dtype=torch.float64)
    return m, l, eps, g


def gen_pm(name, sig, radii, B=24, seed=0, scale_m=1.0, scalar=False):
    from mt.mvae import utils
    comps = utils.parse_components(sig, False)
    dims = [c.true_dim for c in comps]
    amb = sum(c.dim for c in comps)
    m, l, eps, g = _inputs(dims, B, seed, scale_m, scalar)
    gz = torch.randn(B, amb, generator=g, dtype=torch.float64)
    gkl = torch.randn(B, len(dims), generator=g, dtype=torch.float64)
    out = {"m": m.numpy(), "l": l.numpy(), "eps": eps.numpy(), "gz": gz.numpy(), "gkl": gkl.numpy(),
           "radii": np.asarray(radii, dtype=np.float64)}
    r64 = rh.ref_product_manifold(sig, m, l, eps, radii, gz, gkl, scalar)
    out.update(r64)
    f = lambda t: t.to(torch.float32)
    r32 = rh.ref_product_manifold(sig, f(m), f(l), f(eps), radii, f(gz), f(gkl), scalar)
    out.update({k + "_f32": v for k, v in r32.items()})
    meta = {"sig": sig, "scalar_parametrization": bool(scalar)}
    np.savez_compressed(os.path.join(HERE, f"pm_{name}.npz"), meta=json.dumps(meta), **out)
    print("pm", name, sig, {k: v.shape for k, v in r64.items() if hasattr(v, "shape")})


def gen_model(name, sig, in_dim, h_dim, B, recon, fixed_curvature, radius, scalar=False, seed=3, beta=1.0):
    for dtype, sfx in ((torch.float64, ""), (torch.float32, "_f32")):
        model = rh.build_model(sig, in_dim, h_dim, fixed_curvature, scalar, recon, seed, torch.float64)
        for ci, c in enumerate(model.components):
            for pn in ("_nradius", "_pradius"):
                if hasattr(c, pn):
                    getattr(c, pn).data.fill_(radius)
            if hasattr(c, "_curvature"):  # 'u': alternate hyperbolic / spherical choices
                c._curvature.data.fill_((-1.0) ** ci * 0.7)
        model = model.to(dtype)
        g = torch.Generator().manual_seed(seed + 1)
        if recon == "bce":
            x = (torch.rand(B, in_dim, generator=g, dtype=torch.float64) < 0.3).to(dtype)
        else:
            x = torch.randn(B, in_dim, generator=g, dtype=torch.float64).to(dtype)
        eps = rh.draw_eps(model, B, seed + 2, dtype)
        res = rh.ref_model_step(model, x, eps, beta)
        if sfx == "":
            out = dict(res)
        else:
            out.update({k + sfx: v for k, v in res.items() if not k.startswith("param.") and k not in ("x", "eps")})
    meta = {"sig": sig, "in_dim": in_dim, "h_dim": h_dim, "recon": recon, "fixed_curvature": fixed_curvature,
            "scalar_parametrization": bool(scalar), "beta": beta}
    np.savez_compressed(os.path.join(HERE, f"model_{name}.npz"), meta=json.dumps(meta), **out)
    print("model", name, sig, "elbo", float(out["elbo"]))


CONV_SUBSAMPLE = 1024  # gradient entries kept per parameter tensor of the conv fixture (evenly strided)


def conv_stride(numel):
    return max(1, numel // CONV_SUBSAMPLE)


def gen_conv(name, sig, B, radius, seed=41, beta=1.0):
    """ConvolutionalVAE (conv_vae.py:28-79; BASELINE cfg5: CIFAR-shaped 3 x 32 x 32 inputs, BCE on real-valued targets,
    image_reconstruction.py:142-143).  The model has 2.1 M parameters: they are NOT stored — the fixture keeps the seed
    (torch's default initialisation is a deterministic function of it; tests rebuild the same parameters and check a
    digest) — and of every gradient tensor an evenly strided subsample plus its norms."""
    model = rh.build_model(sig, 3072, 8192, False, False, "bce", seed, torch.float64, architecture="conv")
    for c in model.components:
        for pn in ("_nradius", "_pradius"):
            if hasattr(c, pn):
                getattr(c, pn).data.fill_(radius)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.rand(B, 3072, generator=g, dtype=torch.float64)
    eps = rh.draw_eps(model, B, seed + 2, torch.float64)
    res = rh.ref_model_step(model, x, eps, beta)
    out = {k: v for k, v in res.items() if not k.startswith(("param.", "grad.")) and k not in ("h", "m", "l")}
    for k, v in res.items():
        if k.startswith("param."):
            flat = v.reshape(-1)
            out["pdigest." + k[6:]] = np.asarray([flat.sum(), np.abs(flat).sum(), flat[0], flat[-1]])
        if k.startswith("grad."):
            flat = v.reshape(-1)
            out["gsub." + k[5:]] = flat[::conv_stride(flat.size)].copy()
            out["gnorm." + k[5:]] = np.asarray([np.linalg.norm(flat), np.abs(flat).max()])
    meta = {"sig": sig, "in_dim": 3072, "h_dim": 8192, "recon": "bce", "fixed_curvature": False, "seed": seed,
            "scalar_parametrization": False, "beta": beta, "radius": radius, "subsample": CONV_SUBSAMPLE}
    np.savez_compressed(os.path.join(HERE, f"conv_{name}.npz"), meta=json.dumps(meta), **out)
    print("conv", name, sig, "elbo", float(out["elbo"]), "bytes", os.path.getsize(os.path.join(HERE, f"conv_{name}.npz")))


def gen_convs():
    gen_conv("h2_s2_e2", "h2,s2,e2", B=6, radius=1.0)
    gen_conv("p2_d2_h3", "p2,d2,h3", B=4, radius=1.5, seed=43, beta=0.7)


def gen_loglik(name, sig, in_dim, h_dim, B, n, recon, radius, scalar=False, seed=11):
    out = {}
    for dtype, sfx in ((torch.float64, ""), (torch.float32, "_f32")):
        model = rh.build_model(sig, in_dim, h_dim, False, scalar, recon, seed, torch.float64)
        for ci, c in enumerate(model.components):
            for pn in ("_nradius", "_pradius"):
                if hasattr(c, pn):
                    getattr(c, pn).data.fill_(radius)
            if hasattr(c, "_curvature"):  # 'u': alternate hyperbolic / spherical choices
                c._curvature.data.fill_((-1.0) ** ci * 0.7)
        model = model.to(dtype)
        g = torch.Generator().manual_seed(seed + 1)
        if recon == "bce":
            x = (torch.rand(B, in_dim, generator=g, dtype=torch.float64) < 0.3).to(dtype)
        else:
            x = torch.randn(B, in_dim, generator=g, dtype=torch.float64).to(dtype)
        eps = rh.draw_eps(model, B, seed + 2, dtype, n_samples=n)
        res = rh.ref_log_likelihood(model, x, eps, n)
        if sfx == "":
            out.update(res)
            out["x"] = x.numpy()
            out["eps"] = torch.cat(eps, -1).numpy()
            for pname, prm in model.named_parameters():
                out["param." + pname] = prm.detach().numpy()
        else:
            out.update({k + sfx: v for k, v in res.items()})
    meta = {"sig": sig, "in_dim": in_dim, "h_dim": h_dim, "recon": recon, "scalar_parametrization": bool(scalar), "n": n}
    np.savez_compressed(os.path.join(HERE, f"loglik_{name}.npz"), meta=json.dumps(meta), **out)
    print("loglik", name, sig, "sum ll", float(out["log_p_x"].sum()), "sum mi", float(out["mi"].sum()), "cov_norm",
          float(out["cov_norm"]))


def gen_ops(letter, n=3, B=16, R=2.0, seed=5):
    from mt.mvae.ops import Hyperboloid, Sphere, PoincareBall, Euclidean
    from mt.mvae.ops import hyperbolics as H, spherical as S, poincare as P, spherical_projected as SP
    from mt.mvae.ops import StereographicallyProjectedSphere
    from mt.mvae.distributions import WrappedNormal
    with rh.default_dtype(torch.float64):
        Rt = torch.tensor(R, dtype=torch.float64)
        man = {"h": lambda: Hyperboloid(lambda: Rt), "s": lambda: Sphere(lambda: Rt),
               "p": lambda: PoincareBall(lambda: Rt), "e": lambda: Euclidean(),
               "d": lambda: StereographicallyProjectedSphere(lambda: Rt)}[letter]()
        g = torch.Generator().manual_seed(seed)
        t1 = torch.randn(B, n, generator=g, dtype=torch.float64) * 0.8
        t2 = torch.randn(B, n, generator=g, dtype=torch.float64) * 0.8
        v = torch.randn(B, n, generator=g, dtype=torch.float64) * 0.7
        scale = torch.rand(B, n, generator=g, dtype=torch.float64) + 0.3
        x = man.exp_map_mu0(t1)   # point 1
        y = man.exp_map_mu0(t2)   # point 2
        out = {"t1": t1, "t2": t2, "v": v, "scale": scale, "x": x, "y": y, "R": torch.tensor(R)}
        out["inv_exp_map_mu0_x"] = man.inverse_exp_map_mu0(x)
        z, (u, vv) = man.sample_projection_mu0(v, x)
        out["sp_z"], out["sp_u"], out["sp_v"] = z, u, vv
        if letter != "e":  # Euclidean.inverse_sample_projection_mu0 `raise`s its result (euclidean.py:55-56)
            iu, iv = man.inverse_sample_projection_mu0(z, x)
            out["isp_u"], out["isp_v"] = iu, iv
        mod = {"h": H, "s": S, "p": P}.get(letter)
        if letter in ("h", "s"):
            tv = torch.cat((torch.zeros(B, 1, dtype=torch.float64), v), -1)  # tangent at mu0
            out["tangent_mu0"] = tv
            pt = man.parallel_transport_mu0(tv, x)
            out["pt_mu0"] = pt
            out["inv_pt_mu0"] = man.inverse_parallel_transport_mu0(pt, x)
            out["exp_map"] = mod.exp_map(pt, x, radius=Rt)
            out["inv_exp_map"] = mod.inverse_exp_map(y, x, radius=Rt)
            out["logdet_u"] = mod._logdet(u, Rt)
            if letter == "h":
                out["distance"] = Rt * H.acosh(-H.lorentz_product(x, y, keepdim=True) / (Rt**2))
                out["to_poincare"] = H.lorentz_to_poincare(x, Rt)
                out["lorentz_product"] = H.lorentz_product(x, y, keepdim=True)
            else:
                nd = torch.sum(x * y, dim=-1, keepdim=True) / Rt**2
                out["distance"] = Rt * torch.acos(torch.clamp(nd, min=-1., max=1.))
                out["to_poincare"] = S.spherical_to_projected(x, Rt)
        elif letter == "p":
            pt = man.parallel_transport_mu0(v, x)
            out["pt_mu0"] = pt
            out["inv_pt_mu0"] = man.inverse_parallel_transport_mu0(pt, x)
            out["exp_map"] = P.exp_map(pt, x, radius=Rt)
            out["inv_exp_map"] = P.inverse_exp_map(y, x, radius=Rt)
            out["distance"] = P.poincare_distance(x, y, radius=Rt)
            out["mobius_add"] = P.pm.mobius_add(x, y, c=P._c(Rt))
            out["from_poincare"] = P.poincare_to_lorentz(x, Rt)
            out["logdet_zmu"] = man.logdet(x, scale, z, (u, vv))
        elif letter == "d":
            pt = man.parallel_transport_mu0(v, x)
            out["pt_mu0"] = pt
            out["inv_pt_mu0"] = man.inverse_parallel_transport_mu0(pt, x)
            out["exp_map"] = SP.exp_map(pt, x, radius=Rt)
            out["inv_exp_map"] = SP.inverse_exp_map(y, x, radius=Rt)
            out["distance"] = SP.spherical_projected_distance(x, y, K=SP._c(Rt))
            out["mobius_add"] = SP.mob_add(x, y, SP._c(Rt))
            out["from_poincare"] = SP.projected_to_spherical(x, Rt)
            out["logdet_zmu"] = man.logdet(x, scale, z, (u, vv))
        else:
            from mt.mvae.ops import euclidean as E
            out["pt_mu0"] = man.parallel_transport_mu0(v, x)
            out["inv_pt_mu0"] = man.inverse_parallel_transport_mu0(v, x)
            out["exp_map"] = E.exp_map(v, x)
            out["inv_exp_map"] = E.inverse_exp_map(y, x)
            out["distance"] = 2 * torch.norm(x - y, dim=-1, p=2, keepdim=True)
        if letter != "e":
            q = WrappedNormal(x, scale, man)
            with rh.injected_noise([v / scale]):
                zq, data = q.rsample_with_parts()
            out["wn_z"], out["wn_u"], out["wn_v"] = zq, data[0], data[1]
            out["wn_logq_parts"] = q.log_prob_from_parts(zq, data)
            out["wn_logq"] = q.log_prob(zq)
            p = WrappedNormal(man.mu_0(x.shape), torch.ones_like(scale), man)
            out["wn_logp"] = p.log_prob(zq)
    np.savez_compressed(os.path.join(HERE, f"ops_{letter}.npz"), **{k: t.detach().numpy() for k, t in out.items()})
    print("ops", letter, sorted(out))


def gen_kat():
    s = torch.tensor([0.8, 1.3], dtype=torch.float64)
    v = torch.tensor([0.3, -0.7], dtype=torch.float64)
    m = torch.tensor([[0.5, 0.25]], dtype=torch.float64)
    l = torch.log(torch.expm1(s - 1e-5))[None]
    eps = (v / s)[None]
    kat = {"m": m[0].tolist(), "l": l[0].tolist(), "eps": eps[0].tolist(), "R": 2.0, "sigma": s.tolist(),
           "v": v.tolist()}
    for sig in ("h2", "s2", "p2", "e2"):
        amb = 3 if sig[0] in "hs" else 2
        r = rh.ref_product_manifold(sig, m, l, eps, [2.0], torch.zeros(1, amb, dtype=torch.float64),
                                    torch.ones(1, 1, dtype=torch.float64))
        kat[sig] = {k: np.asarray(val).reshape(-1).tolist() for k, val in r.items()}
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("kat written")


def gen_logliks():
    gen_loglik("h2_s2_e2_bce", "h2,s2,e2", in_dim=20, h_dim=16, B=12, n=7, recon="bce", radius=1.0)
    gen_loglik("h2_p2_nll", "h2,p2", in_dim=10, h_dim=16, B=9, n=5, recon="nll", radius=2.0)
    gen_loglik("cfg3_small_bce", "h6,h6,s6,s6,e6", in_dim=24, h_dim=32, B=8, n=6, recon="bce", radius=10.0)
    gen_loglik("scalar_s3_p2_e3_bce", "s3,p2,e3", in_dim=18, h_dim=16, B=10, n=4, recon="bce", radius=1.5, scalar=True)


def gen_ds():
    """'d' (stereographically projected sphere, spherical_projected.py) fixtures."""
    gen_pm("d2_s2_d3_e2", "d2,s2,d3,e2", [1.0, 1.3, 2.0, 0.0], seed=21, scale_m=0.7)
    gen_pm("d6_p2_R", "d6,p2", [3.0, 1.5], seed=22, scale_m=0.8)
    gen_pm("scalar_d2_h2", "d2,h2", [0.9, 1.1], seed=23, scalar=True, scale_m=0.5)
    gen_pm("d1_d40", "d1,d40", [1.0, 4.0], B=8, seed=24, scale_m=0.3)
    gen_model("d2_h2_e2_bce", "d2,h2,e2", in_dim=20, h_dim=16, B=12, recon="bce", fixed_curvature=False, radius=1.5,
              seed=25)
    gen_loglik("d2_s2_bce", "d2,s2", in_dim=16, h_dim=16, B=9, n=5, recon="bce", radius=1.2, seed=26)
    gen_ops("d")


def gen_us():
    """'u' (universal, universal.py) fixtures: negative, positive and ~zero curvature choices."""
    gen_pm("u2_u3_u2_h2", "u2,u3,u2,h2", [-0.8, 1.3, 5e-7, 1.0], seed=31, scale_m=0.5)
    gen_pm("u6_u6", "u6,u6", [0.05, -2.0], seed=32, scale_m=0.6)
    gen_pm("scalar_u2_u2", "u2,u2", [-0.3, 0.4], seed=33, scalar=True, scale_m=0.5)
    gen_model("u2_u2_e2_bce", "u2,u2,e2", in_dim=20, h_dim=16, B=12, recon="bce", fixed_curvature=False, radius=1.0,
              seed=35)
    gen_loglik("u2_u2_bce", "u2,u2", in_dim=16, h_dim=16, B=9, n=5, recon="bce", radius=1.0, seed=36)


if __name__ == "__main__":
    if "--only-u" in sys.argv:
        gen_us()
        sys.exit(0)
    if "--only-d" in sys.argv:
        gen_ds()
        sys.exit(0)
    if "--only-conv" in sys.argv:  # added later: leaves the other committed fixtures byte-identical
        gen_convs()
        sys.exit(0)
    if "--only-loglik" in sys.argv:  # added later: leaves the other committed fixtures byte-identical
        gen_logliks()
        sys.exit(0)
    gen_kat()
    gen_logliks()
    gen_ds()
    gen_us()
    gen_convs()
    gen_pm("h2_s2_e2_R1", "h2,s2,e2", [1.0, 1.0, 0.0])
    gen_pm("cfg3_R10", "h6,h6,s6,s6,e6", [10.0, 10.0, 10.0, 10.0, 0.0], seed=1)
    gen_pm("cfg3_Rmixed", "h6,h6,s6,s6,e6", [1.5, 0.7, 2.0, 1.0, 0.0], seed=2, scale_m=0.6)
    gen_pm("p2_h2_R2", "p2,h2", [2.0, 2.0], seed=3)
    gen_pm("scalar_h2_s2_p3_e2", "h2,s2,p3,e2", [1.3, 0.8, 2.0, 0.0], seed=4, scalar=True)
    gen_pm("dims1", "h1,s1,p1,e1", [1.0, 1.0, 1.0, 0.0], seed=5)
    gen_pm("big_h40_s40_p40", "h40,s40,p40", [1.0, 1.0, 3.0], B=6, seed=6, scale_m=0.25)
    gen_pm("tiny_m", "h2,s2,p2", [1.0, 1.0, 1.0], seed=7, scale_m=1e-4)
    gen_model("h2_s2_e2_bce", "h2,s2,e2", in_dim=20, h_dim=16, B=12, recon="bce", fixed_curvature=False, radius=1.0)
    gen_model("e2_fixed_bce", "e2", in_dim=20, h_dim=16, B=12, recon="bce", fixed_curvature=True, radius=1.0)
    gen_model("h2_p2_nll", "h2,p2", in_dim=10, h_dim=16, B=12, recon="nll", fixed_curvature=False, radius=2.0)
    gen_model("cfg3_small_bce", "h6,h6,s6,s6,e6", in_dim=24, h_dim=32, B=8, recon="bce", fixed_curvature=False,
              radius=10.0, beta=0.5)
    for letter in "hspe":
        gen_ops(letter)
