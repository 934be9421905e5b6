"""CPU tests: the C oracle against the golden vectors produced by the reference itself
(tests/golden/generate_golden.py) and the known-answer vectors of SURVEY.md App. C.2."""
# NOTE on the Poincare / projected-sphere / universal fixtures (helpers.shim_pinned): the reference delegates that
# arithmetic to geoopt==0.1.0, which is absent here; the fixtures were generated through oracle/ref_shims/geoopt, a
# restatement from memory.  They pin the oracle to the reference's formulas and call structure; geoopt's guard constants
# are the shim's ("parity unpinned" where those clamps bind).
import json
import os

import numpy as np
import pytest

from helpers import (GOLDEN, load_golden, loglik_golden_names, model_golden_names, normwise, pack_ml, pm_golden_names, radii_array,
                     unpack_gml)

F64_TOL = 1e-9   # oracle(float64) vs reference(float64): same operations, different summation order only
F32_TOL = 2e-4   # oracle(float32) vs reference(float32): both carry float32 rounding; normwise


@pytest.mark.parametrize("name", pm_golden_names())
def test_pm_forward_backward_f64(oracle, name):
    g, meta = load_golden(name)
    desc = oracle.make_desc(meta["sig"], scalar_parametrization=meta["scalar_parametrization"])
    ml = pack_ml(desc, g["m"], g["l"])
    R = radii_array(g["radii"], np.float64)
    f = oracle.pm_forward(desc, ml, g["eps"], R, want=("z", "kl", "mu", "sigma", "u", "logq", "logp"))
    for k in ("z", "kl", "mu", "sigma", "u", "logq", "logp"):
        assert normwise(f[k], g[k]) < F64_TOL, k
    gml, gR = oracle.pm_backward(desc, ml, g["eps"], R, g["gz"], g["gkl"])
    gm, gl = unpack_gml(desc, gml)
    assert normwise(gm, g["gm"]) < F64_TOL
    assert normwise(gl, g["gl"]) < F64_TOL
    assert normwise(gR, g["gR"]) < 1e-8


@pytest.mark.parametrize("name", pm_golden_names())
def test_pm_forward_backward_f32(oracle, name):
    """float32 oracle vs the reference's own float32 run AND vs the float64 truth (normwise)."""
    g, meta = load_golden(name)
    desc = oracle.make_desc(meta["sig"], scalar_parametrization=meta["scalar_parametrization"])
    f32 = np.float32
    ml = pack_ml(desc, g["m"].astype(f32), g["l"].astype(f32))
    R = radii_array(g["radii"], f32)
    f = oracle.pm_forward(desc, ml, g["eps"].astype(f32), R, want=("z", "kl", "mu", "sigma"))
    for k in ("z", "kl", "mu", "sigma"):
        # where the reference's own float32 run is further than F32_TOL from its float64 run (small radii: the
        # Poincare log-det cancels), a float32 restatement is held to a small multiple of that distance
        tol = max(F32_TOL, 3 * normwise(g[k + "_f32"], g[k]))
        assert normwise(f[k], g[k + "_f32"]) < tol, k
        assert normwise(f[k], g[k]) < tol, k
    gml, gR = oracle.pm_backward(desc, ml, g["eps"].astype(f32), R, g["gz"].astype(f32), g["gkl"].astype(f32))
    gm, gl = unpack_gml(desc, gml)
    # the reference's float32 backward is itself only ~1e-3 normwise accurate on these inputs
    ref_err_m = normwise(g["gm_f32"], g["gm"])
    ref_err_l = normwise(g["gl_f32"], g["gl"])
    assert normwise(gm, g["gm"]) < max(5e-4, 3 * ref_err_m)
    assert normwise(gl, g["gl"]) < max(5e-4, 3 * ref_err_l)


def test_known_answer_vectors(oracle):
    """SURVEY.md App. C.2 (R=2, m=[.5,.25], v=[.3,-.7], sigma=[.8,1.3])."""
    with open(os.path.join(GOLDEN, "kat.json")) as fh:
        kat = json.load(fh)
    survey = {
        "h2": dict(mu=[2.078634952699212, 0.5065358953384105, 0.25326794766920524], logq=-2.8095787992285097,
                   kl=0.18559320444190153),
        "s2": dict(mu=[1.9223823036006809, 0.4935149673944569, 0.24675748369722844], logq=-2.761243314915615,
                   kl=0.14869617650860523),
        "p2": dict(mu=[0.4873735955230119, 0.24368679776150595], logq=-2.8095787992285097, kl=0.6891761111118648),
        "e2": dict(mu=[0.25, 0.125]),
    }
    for sig in ("h2", "s2", "p2", "e2"):
        desc = oracle.make_desc(sig)
        ml = np.asarray([kat["m"] + kat["l"]], dtype=np.float64)
        eps = np.asarray([kat["eps"]], dtype=np.float64)
        f = oracle.pm_forward(desc, ml, eps, np.asarray([2.0]), want=("z", "kl", "mu", "sigma", "u", "logq", "logp"))
        for k in ("mu", "z", "u", "kl", "logq", "logp"):
            np.testing.assert_allclose(f[k].reshape(-1), kat[sig][k], rtol=1e-11, atol=1e-13, err_msg=f"{sig}.{k}")
        np.testing.assert_allclose(f["mu"].reshape(-1), survey[sig]["mu"], rtol=1e-12)
        # the survey's vectors used sigma exactly [.8,1.3]; ours go through softplus^-1, so 1e-8 is the bar here
        for k in ("logq", "kl"):
            if k in survey[sig]:
                np.testing.assert_allclose(f[k].reshape(-1)[0], survey[sig][k], rtol=0, atol=2e-8)


@pytest.mark.parametrize("name", model_golden_names())
def test_model_step_f64(oracle, name):
    g, meta = load_golden(name)
    params = {k[len("param."):]: v for k, v in g.items() if k.startswith("param.")}
    vae = oracle.OracleVAE(meta["sig"], meta["in_dim"], meta["h_dim"], meta["recon"], meta["scalar_parametrization"])
    out = vae.step(params, g["x"], g["eps"], beta=meta["beta"])
    for k in ("h", "z", "kl", "mu", "sigma", "logits", "bce"):
        assert normwise(out[k], g[k]) < F64_TOL, k
    assert abs(out["elbo"] - g["elbo"]) < 1e-9 * abs(g["elbo"])
    assert abs(out["bce_sum"] - g["bce_sum"]) < 1e-9 * abs(g["bce_sum"])
    assert abs(out["kl_sum"] - g["kl_sum"]) < 1e-9 * abs(g["kl_sum"])
    for k, v in out["grads"].items():
        ref = g["grad." + k]
        if meta["fixed_curvature"] and "radius" in k:
            continue  # requires_grad=False in the reference
        assert normwise(v, ref) < 1e-8, k


def test_plain_c_linear_matches_blas(oracle):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((33, 17))
    W = rng.standard_normal((9, 17))
    b = rng.standard_normal(9)
    a = oracle.linear(x, W, b, relu=True, use_blas=True)
    c = oracle.linear(x, W, b, relu=True, use_blas=False)
    np.testing.assert_allclose(a, c, rtol=1e-12, atol=1e-12)


def test_elbo_and_recon(oracle):
    rng = np.random.default_rng(1)
    lg = rng.standard_normal((7, 13)) * 3
    x = (rng.random((7, 13)) < 0.4).astype(np.float64)
    rs, gl = oracle.recon("bce", lg, x, want_grad=True)
    ref = np.maximum(lg, 0) - lg * x + np.log1p(np.exp(-np.abs(lg)))
    np.testing.assert_allclose(rs, ref.sum(-1), rtol=1e-12)
    np.testing.assert_allclose(gl, 1 / (1 + np.exp(-lg)) - x, rtol=1e-12)
    rs, gl = oracle.recon("nll", lg, x, want_grad=True)
    np.testing.assert_allclose(rs, (0.5 * (x - lg)**2 + 0.5 * np.log(2 * np.pi)).sum(-1), rtol=1e-12)
    kl = rng.standard_normal((7, 3))
    out = oracle.elbo(rs, kl, 0.7)
    np.testing.assert_allclose(out[0], rs.sum())
    np.testing.assert_allclose(out[1], kl.sum())
    np.testing.assert_allclose(out[2], (-rs - 0.7 * kl.sum(-1)).sum())
    np.testing.assert_allclose(out[3:], kl.sum(0))


@pytest.mark.parametrize("name", loglik_golden_names())
def test_log_likelihood_vs_reference(oracle, name):
    """OracleVAE.log_likelihood against the reference's own ModelVAE.log_likelihood (vae.py:82-123) run with the
    same injected n x B draws: IWAE estimate, mutual information and cov_norm."""
    g, meta = load_golden(name)
    params = {k[len("param."):]: v for k, v in g.items() if k.startswith("param.")}
    o = oracle.OracleVAE(meta["sig"], meta["in_dim"], meta["h_dim"], meta["recon"], meta["scalar_parametrization"])
    r = o.log_likelihood(params, g["x"], g["eps"])
    assert r["log_p_x"].shape == g["log_p_x"].shape == (g["x"].shape[0],)
    assert normwise(r["log_p_x"], g["log_p_x"]) < F64_TOL
    assert normwise(r["mi"], g["mi"]) < F64_TOL
    assert abs(r["cov_norm"] - float(g["cov_norm"])) < F64_TOL * float(g["cov_norm"])
    # float32 oracle against the float64 truth and the reference's float32 run
    p32 = {k: v.astype(np.float32) for k, v in params.items()}
    r32 = o.log_likelihood(p32, g["x"].astype(np.float32), g["eps"].astype(np.float32))
    for k in ("log_p_x", "mi"):
        assert normwise(r32[k], g[k]) < max(F32_TOL, 3 * normwise(g[k + "_f32"], g[k])), k
    assert abs(r32["cov_norm"] - float(g["cov_norm"])) < 1e-4 * float(g["cov_norm"])


def test_oracle_trainer_update_is_torch_adam_and_sgd(oracle):
    """oracle.OracleTrainer (the k-step reference of tests/test_gpu_timed_path.py and scripts/dp_check.py) applies
    exactly torch.optim.Adam + SGD(1e-4) on the radii (Trainer.build_optimizer, train.py:327-360) to the oracle's
    gradients."""
    import torch
    sig, B, D, H = "h2,s2,e2", 64, 20, 16
    rng = np.random.default_rng(0)
    ov = oracle.OracleVAE(sig, D, H, "bce", False)
    params = {}
    for i in range(3):
        params[f"components.{i}.fc_mean.weight"] = rng.standard_normal((2, H)) * 0.3
        params[f"components.{i}.fc_mean.bias"] = rng.standard_normal(2) * 0.1
        params[f"components.{i}.fc_logvar.weight"] = rng.standard_normal((2, H)) * 0.3
        params[f"components.{i}.fc_logvar.bias"] = rng.standard_normal(2) * 0.1
    params["components.0._nradius"] = np.asarray(1.0)
    params["components.1._pradius"] = np.asarray(1.0)
    for nm, shp in (("fc_e0", (H, D)), ("fc_d0", (H, 8)), ("fc_logits", (D, H))):
        params[nm + ".weight"] = rng.standard_normal(shp) * 0.3
        params[nm + ".bias"] = rng.standard_normal(shp[0]) * 0.1
    tr = oracle.OracleTrainer(ov, params)
    tp = {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=True) for k, v in params.items()}
    adam = torch.optim.Adam([v for k, v in tp.items() if "radius" not in k], lr=1e-3)
    sgd = torch.optim.SGD([v for k, v in tp.items() if "radius" in k], lr=1e-4)
    for _ in range(4):
        x = (rng.random((B, D)) < 0.3).astype(np.float64)
        eps = rng.standard_normal((B, 6))
        cur = {k: v.detach().numpy().copy() for k, v in tp.items()}
        g = ov.step(cur, x, eps, beta=0.7)["grads"]
        for k, v in tp.items():
            v.grad = torch.tensor(np.asarray(g[k], dtype=np.float64)).reshape(v.shape)
        adam.step()
        sgd.step()
        tr.step(x, eps, beta=0.7)
    for k, v in tp.items():
        assert np.allclose(v.detach().numpy(), tr.params[k], rtol=1e-10, atol=1e-12), k


def test_cpu_baseline_step_matches_oracle(oracle):
    """oracle/cpu_baseline.py — what bench.py's CPU arm times — computes the same step as OracleVAE.step (float32)."""
    import torch
    import cpu_baseline
    sig, B, D, H = "h2,s2,p2,e2", 256, 40, 32
    rng = np.random.default_rng(1)
    ov = oracle.OracleVAE(sig, D, H, "bce", False)
    params = {}
    for i in range(4):
        params[f"components.{i}.fc_mean.weight"] = (rng.standard_normal((2, H)) * 0.05).astype(np.float32)
        params[f"components.{i}.fc_mean.bias"] = (rng.standard_normal(2) * 0.1).astype(np.float32)
        params[f"components.{i}.fc_logvar.weight"] = (rng.standard_normal((2, H)) * 0.05).astype(np.float32)
        params[f"components.{i}.fc_logvar.bias"] = (rng.standard_normal(2) * 0.1).astype(np.float32)
    params["components.0._nradius"] = np.asarray(1.3, dtype=np.float32)
    params["components.1._pradius"] = np.asarray(0.8, dtype=np.float32)
    params["components.2._nradius"] = np.asarray(1.1, dtype=np.float32)
    for nm, shp in (("fc_e0", (H, D)), ("fc_d0", (H, 10)), ("fc_logits", (D, H))):
        params[nm + ".weight"] = (rng.standard_normal(shp) * 0.1).astype(np.float32)
        params[nm + ".bias"] = (rng.standard_normal(shp[0]) * 0.1).astype(np.float32)
    x = (rng.random((B, D)) < 0.3).astype(np.float32)
    eps = rng.standard_normal((B, 8)).astype(np.float32)
    ref = ov.step({k: v.astype(np.float64) for k, v in params.items()}, x.astype(np.float64), eps.astype(np.float64),
                  beta=0.9)
    cpu = cpu_baseline.CpuTrainStep(sig, D, H, "bce", params)
    out = cpu.step(torch.from_numpy(x), torch.from_numpy(eps), beta=0.9, update=False)
    assert abs(out["elbo"] - ref["elbo"]) < 1e-4 * abs(ref["elbo"])   # float32 oracle (reference-order arithmetic)
    assert abs(out["kl_sum"] - ref["kl_sum"]) < 2e-3 * abs(ref["kl_sum"])
    for k_cpu, k_ref in (("fc_e0.W", "fc_e0.weight"), ("fc_e0.b", "fc_e0.bias"), ("fc_d0.W", "fc_d0.weight"),
                         ("fc_logits.W", "fc_logits.weight"), ("fc_logits.b", "fc_logits.bias")):
        assert normwise(out["grads"][k_cpu].numpy(), ref["grads"][k_ref]) < 2e-4, k_cpu
    gWh = np.concatenate([np.concatenate([ref["grads"][f"components.{i}.fc_mean.weight"],
                                          ref["grads"][f"components.{i}.fc_logvar.weight"]]) for i in range(4)])
    assert normwise(out["grads"]["Wh"].numpy(), gWh) < 2e-4
    gR = [float(ref["grads"].get(f"components.{i}.{nm}", 0.0)) for i, nm in enumerate(("_nradius", "_pradius", "_nradius"))]
    assert np.allclose(out["gR"][:3], gR, rtol=2e-3, atol=1e-3)
    # and the update moves the parameters (Adam: about lr per step) and the radii
    w0 = cpu.p["fc_e0.W"].clone()
    cpu.step(torch.from_numpy(x), torch.from_numpy(eps), beta=0.9)
    assert 1e-4 < float((cpu.p["fc_e0.W"] - w0).abs().max()) < 2e-3 and cpu.R[0] != np.float32(1.3)


@pytest.mark.parametrize("name", __import__("helpers").conv_golden_names())
def test_conv_oracle_matches_reference_golden(oracle, name):
    """oracle.OracleConvVAE (numpy restatement of ConvolutionalVAE, conv_vae.py:28-79, + the latent path) against the
    run of the reference's own model: forward tensors, statistics and (subsampled) autograd gradients at 1e-9."""
    from helpers import check_conv_digest, conv_params_from_seed
    g, meta = load_golden(name)
    params = conv_params_from_seed(meta["sig"], meta["seed"], meta["radius"])
    check_conv_digest(params, g)
    out = oracle.OracleConvVAE(meta["sig"]).step(params, g["x"], g["eps"], beta=meta["beta"])
    for k in ("z", "mu", "sigma", "kl", "logits", "bce"):
        assert normwise(out[k], g[k]) < 1e-9, k
    assert abs(out["elbo"] - g["elbo"]) < 1e-10 * abs(g["elbo"])
    stride = lambda n: max(1, n // meta["subsample"])  # noqa: E731
    for k in params:
        got = np.asarray(out["grads"][k], dtype=np.float64).reshape(-1)
        assert normwise(got[::stride(got.size)], g["gsub." + k]) < 1e-8, k
        assert abs(np.linalg.norm(got) - g["gnorm." + k][0]) < 1e-8 * max(g["gnorm." + k][0], 1e-30), k
