"""CPU tests: the C oracle against the golden vectors produced by the reference itself
(tests/golden/generate_golden.py) and the known-answer vectors of SURVEY.md App. C.2."""
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
