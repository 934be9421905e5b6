"""CPU check of the ARITHMETIC of the fused product-manifold kernels: mvae_b200/csrc/pm_math.cuh is compiled as plain
C++ (tests/host_emul/pm_emul.cpp, libm stand-ins for the one-instruction MUFU ops) and compared with the float64 oracle.
This pins the closed forms and the hand-derived reverse sweep before any GPU time is spent; the GPU parity tests
(tests/test_gpu_kernels.py) then check the real kernels, MUFU approximations included.  Test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import normwise

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "pm_emul.cpp")
HDR = os.path.join(HERE, "..", "mvae_b200", "csrc", "pm_math.cuh")
SO = os.path.join(HERE, "host_emul", "libpm_emul.so")


@pytest.fixture(scope="module")
def emul():
    if (not os.path.exists(SO)) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC], check=True)
    return ctypes.CDLL(SO)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _scale_of(sig):
    """'d' components put mu = R tan(|m|/R) m/|m| (spherical_projected.py:150-154): head outputs are kept away from the
    pole |m| = pi R / 2, where the map (and any evaluation of it, the reference's included) is ill conditioned."""
    if "d" not in sig:
        return 1.0
    nmax = max(int(tok[1:]) for tok in sig.split(","))
    return 0.35 / (nmax / 2.0) ** 0.5  # |m| ~ 0.35 sqrt(2) on average whatever the dimension (radii in the tests >= 1)


def _inputs(desc, B, seed, scale_m=1.0):
    rng = np.random.default_rng(seed)
    ml = (rng.standard_normal((B, desc.ld_ml)) * scale_m).astype(np.float32)
    eps = rng.standard_normal((B, desc.ld_eps)).astype(np.float32)
    gz = rng.standard_normal((B, desc.ld_z)).astype(np.float32)
    return ml, eps, gz


def _drop_pole_rows(oracle, sig, desc, radius, ml, eps, gz):
    """'d' only: the stereographic chart sends the antipode of mu0 to infinity, so a sample that lands near it has
    unbounded coordinates and gradients (in the reference too).  Such rows (|z| > 6 R) are conditioning, not
    arithmetic; they are left out of the comparison (tests/test_gpu_kernels.py compares the 'd' kernels with the
    reference's own autograd on the golden fixtures)."""
    if "d" not in sig:
        return ml, eps, gz
    f = oracle.pm_forward(desc, ml.astype(np.float64), eps.astype(np.float64), radius.astype(np.float64),
                          want=("z", "sigma"))
    keep = np.ones(ml.shape[0], bool)
    for i in range(desc.C):
        c = desc.comp[i]
        if c.type == oracle.TYPE_OF_LETTER["d"]:
            keep &= np.abs(f["z"][:, c.z_off:c.z_off + c.d]).max(1) < 6.0 * radius[i]
            # ... and the log-det (n-1) log|sin t| is singular at |v| = k pi R (k >= 1): gradients blow up there and
            # no float32 evaluation tracks the float64 one
            v = eps[:, c.eps_off:c.eps_off + c.n] * f["sigma"][:, c.eps_off:c.eps_off + c.n]
            t = np.linalg.norm(v, axis=1) / radius[i]
            keep &= (t < 1.5) | (np.abs(np.sin(t)) > 0.05)
    assert keep.mean() > 0.5
    return ml[keep], eps[keep], gz[keep]


def _emul_forward(emul, desc, ml, eps, radius, dyn=0):
    B = ml.shape[0]
    z = np.zeros((B, desc.ld_z), np.float32)
    kl = np.zeros((B, desc.C), np.float32)
    mu = np.zeros((B, desc.ld_z), np.float32)
    sigma = np.zeros((B, desc.ld_eps), np.float32)
    emul.pm_emul_forward(ctypes.byref(desc), ctypes.c_int64(B), _p(ml), _p(eps), _p(radius), _p(z), _p(kl), _p(mu),
                         _p(sigma), ctypes.c_int(dyn))
    return z, kl, mu, sigma


def _emul_backward(emul, desc, ml, eps, radius, gz, beta, dyn=0):
    B = ml.shape[0]
    gml = np.zeros_like(ml)
    gR = np.zeros(desc.C, np.float64)
    emul.pm_emul_backward(ctypes.byref(desc), ctypes.c_int64(B), _p(ml), _p(eps), _p(radius), _p(gz), None,
                          ctypes.c_float(beta), _p(gml), _p(gR), ctypes.c_int(dyn))
    return gml, gR


SIGS = ["h2,s2,e2", "h6,h6,s6,s6,e6", "p2", "h2", "s2", "e2", "h3,s5,p4,e1", "h8,s8,p8", "s1,h1,p1", "h12,s9,p7,e11",
        "d2", "d6,d3,s2", "d1,d8,d11"]


@pytest.mark.parametrize("sig", SIGS)
@pytest.mark.parametrize("R", [1.0, 2.0, 10.0])
@pytest.mark.parametrize("scalar", [False, True])
def test_forward_matches_oracle(oracle, emul, sig, R, scalar):
    desc = oracle.make_desc(sig, scalar_parametrization=scalar)
    radius = np.full(desc.C, R, np.float32)
    ml, eps, _ = _drop_pole_rows(oracle, sig, desc, radius, *_inputs(desc, 512, 3, _scale_of(sig)))
    ref = oracle.pm_forward(desc, ml.astype(np.float64), eps.astype(np.float64), radius.astype(np.float64),
                            want=("z", "kl", "mu", "sigma"))
    z, kl, mu, sigma = _emul_forward(emul, desc, ml, eps, radius)
    assert normwise(z, ref["z"]) < 2e-5
    assert normwise(mu, ref["mu"]) < 2e-5
    assert normwise(sigma, ref["sigma"]) < 2e-6
    # KL: per-sample values of O(1..10); conditioning of the sphere log-det near |v| = pi R is excluded below
    err = np.abs(kl - ref["kl"]) / np.maximum(1.0, np.abs(ref["kl"]))
    assert np.quantile(err, 0.99) < 2e-5
    assert np.median(err) < 2e-6


@pytest.mark.parametrize("sig", SIGS)
@pytest.mark.parametrize("R", [1.0, 2.0, 10.0])
@pytest.mark.parametrize("scalar", [False, True])
def test_backward_matches_oracle(oracle, emul, sig, R, scalar):
    desc = oracle.make_desc(sig, scalar_parametrization=scalar)
    radius = np.full(desc.C, R, np.float32)
    ml, eps, gz = _drop_pole_rows(oracle, sig, desc, radius, *_inputs(desc, 512, 5, _scale_of(sig)))
    gml_ref, gR_ref = oracle.pm_backward(desc, ml.astype(np.float64), eps.astype(np.float64),
                                         radius.astype(np.float64), gz.astype(np.float64), None, 0.7)
    gml, gR = _emul_backward(emul, desc, ml, eps, radius, gz, 0.7)
    # rows near the sphere log-det singularity (|v| ~ pi R) are ill conditioned in any fp32 evaluation: compare the
    # bulk tightly and every row loosely
    row_err = np.abs(gml - gml_ref).max(axis=1) / np.maximum(1.0, np.abs(gml_ref).max(axis=1))
    assert np.quantile(row_err, 0.98) < 5e-5
    assert np.median(row_err) < 5e-6
    assert np.all(np.abs(gR - gR_ref) <= 2e-4 * np.maximum(1.0, np.abs(gml_ref).sum()))


@pytest.mark.parametrize("sig", ["h2,s2,e2", "h6,s6,p3,e6"])
def test_dynamic_dimension_path_equals_static(oracle, emul, sig):
    desc = oracle.make_desc(sig)
    ml, eps, gz = _inputs(desc, 64, 11)
    radius = np.full(desc.C, 1.5, np.float32)
    a = _emul_forward(emul, desc, ml, eps, radius, dyn=0)
    b = _emul_forward(emul, desc, ml, eps, radius, dyn=1)
    for x, y in zip(a, b):
        np.testing.assert_allclose(x, y, rtol=2e-6, atol=2e-6)
    ga, gRa = _emul_backward(emul, desc, ml, eps, radius, gz, 1.0, dyn=0)
    gb, gRb = _emul_backward(emul, desc, ml, eps, radius, gz, 1.0, dyn=1)
    np.testing.assert_allclose(ga, gb, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(gRa, gRb, rtol=1e-4, atol=1e-4)


def test_sincos_and_atan2_polynomials(emul):
    x = np.concatenate([np.linspace(-40.0, 40.0, 200001), np.linspace(-3e4, 3e4, 100001),
                        np.array([0.0, 1e-8, 1e-4, np.pi, np.pi / 2, 1e5, -2e6])]).astype(np.float32)
    s = np.zeros_like(x)
    c = np.zeros_like(x)
    emul.pm_emul_sincos(ctypes.c_int64(x.size), _p(x), _p(s), _p(c))
    xd = x.astype(np.float64)
    assert np.max(np.abs(s - np.sin(xd))) < 2.5e-7
    assert np.max(np.abs(c - np.cos(xd))) < 2.5e-7
    small = np.abs(xd) < 0.5  # relative accuracy near 0 (A = sin(t)/t)
    assert np.max(np.abs(s[small] - np.sin(xd[small])) / np.maximum(np.abs(np.sin(xd[small])), 1e-30)) < 3e-7
    rng = np.random.default_rng(0)
    y = np.abs(rng.standard_normal(200000)).astype(np.float32)
    xx = rng.standard_normal(200000).astype(np.float32)
    y[:10] = 0.0
    xx[5:15] = 0.0
    r = np.zeros_like(y)
    emul.pm_emul_atan2(ctypes.c_int64(y.size), _p(y), _p(xx), _p(r))
    assert np.max(np.abs(r - np.arctan2(y.astype(np.float64), xx.astype(np.float64)))) < 6e-7
