"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


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
