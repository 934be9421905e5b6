"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def shim_pinned(name_or_sig: str) -> bool:
    """True for fixtures whose reference run went through oracle/ref_shims/geoopt — a from-memory restatement of
    geoopt 0.1.0 (the wheel is absent here and on no local index): every fixture with a `p`, `d` or `u` component.
    Their formulas are pinned by the reference's own property tests; geoopt's guard CONSTANTS
    (MIN_NORM, the artanh / tanh clamps, lambda_x's clamp_min) are the shim's.  Where those clamps bind, parity is
    pinned to the shim, not to the real dependency ("parity unpinned")."""
    body = name_or_sig.split("_", 1)[1] if "_" in name_or_sig else name_or_sig
    return any(tok in ("p", "d", "u") or (tok[:1] in "pdu" and tok[1:2].isdigit())
               for tok in body.replace(",", "_").split("_"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    meta = json.loads(str(d.pop("meta"))) if "meta" in d else {}
    return d, meta


def pm_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith("pm_") and f.endswith(".npz"))


def model_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith("model_") and f.endswith(".npz"))


def loglik_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith("loglik_") and f.endswith(".npz"))


def pack_ml(desc, m, l, dtype=None):
    """Component-ordered [m | l] matrices (reference layout) -> packed ml rows [m_0|l_0|m_1|l_1|...]."""
    dtype = dtype or m.dtype
    B = m.shape[0]
    ml = np.zeros((B, desc.ld_ml), dtype=dtype)
    mo = lo = 0
    for i in range(desc.C):
        c = desc.comp[i]
        ml[:, c.m_off:c.m_off + c.n] = m[:, mo:mo + c.n]
        ml[:, c.l_off:c.l_off + c.l_n] = l[:, lo:lo + c.l_n]
        mo += c.n
        lo += c.l_n
    return ml


def unpack_gml(desc, gml):
    gm = np.concatenate([gml[:, desc.comp[i].m_off:desc.comp[i].m_off + desc.comp[i].n] for i in range(desc.C)], 1)
    gl = np.concatenate([gml[:, desc.comp[i].l_off:desc.comp[i].l_off + desc.comp[i].l_n] for i in range(desc.C)], 1)
    return gm, gl


def normwise(a, b):
    """max|a-b| / max|b| — the parity metric of SURVEY.md App. E / BASELINE.md §4."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = float(np.max(np.abs(b))) if b.size else 0.0
    return float(np.max(np.abs(a - b))) / max(den, 1e-30) if a.size else 0.0


def radii_array(radii, dtype):
    return np.asarray([r if r else 1.0 for r in radii], dtype=dtype)


def conv_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith("conv_") and f.endswith(".npz"))


def conv_params_from_seed(sig, seed, radius, fixed_curvature=False):
    """The parameters the reference's ConvolutionalVAE(8192, parse_components(sig), ...) gets under
    torch.manual_seed(seed) with float64 as default dtype (how tests/golden/generate_golden.py::gen_conv built it):
    construction order of vae.py:55-57 (components) then conv_vae.py:47-55 (e0 e1 e2 d0 d1 d2 d3).  The conv fixtures
    keep only a digest of the 2.1 M parameters; callers check it (pdigest.*)."""
    import torch
    from mvae_b200 import components
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        torch.manual_seed(seed)
        comps = components.parse_components(sig, fixed_curvature)
        params = {}
        for i, c in enumerate(comps):
            c.init_layers(8192, scalar_parametrization=False)
        total_z = sum(c.dim for c in comps)
        layers = [("e0", torch.nn.Conv2d(3, 64, 4, 2, 1)), ("e1", torch.nn.Conv2d(64, 128, 4, 2, 1)),
                  ("e2", torch.nn.Conv2d(128, 512, 4, 2, 1)), ("d0", torch.nn.Linear(total_z, 2048)),
                  ("d1", torch.nn.ConvTranspose2d(128, 256, 4, 2, 1)), ("d2", torch.nn.ConvTranspose2d(256, 64, 4, 2, 1)),
                  ("d3", torch.nn.ConvTranspose2d(64, 3, 4, 2, 1))]
        for i, c in enumerate(comps):
            name, rp = c.radius_parameter()
            if rp is not None:
                params[f"components.{i}.{name}"] = np.asarray(radius if name != "_curvature" else float(rp.detach()))
            for nm in ("fc_mean", "fc_logvar"):
                params[f"components.{i}.{nm}.weight"] = getattr(c, nm).weight.detach().numpy().copy()
                params[f"components.{i}.{nm}.bias"] = getattr(c, nm).bias.detach().numpy().copy()
        for nm, layer in layers:
            params[nm + ".weight"] = layer.weight.detach().numpy().copy()
            params[nm + ".bias"] = layer.bias.detach().numpy().copy()
        return params
    finally:
        torch.set_default_dtype(old)


def check_conv_digest(params, golden):
    for k, v in params.items():
        flat = np.asarray(v, dtype=np.float64).reshape(-1)
        want = golden["pdigest." + k]
        got = np.asarray([flat.sum(), np.abs(flat).sum(), flat[0], flat[-1]])
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12), (k, got, want)
