"""ctypes front end of the CPU oracle (oracle/liboracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product (mvae_b200/) never does.  See oracle/mvae_oracle_impl.h for what each function restates.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MAX_COMPONENTS = 96
TYPE_OF_LETTER = {"e": 0, "h": 1, "s": 2, "p": 3, "d": 4, "u": 5}


class Component(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("type", "n", "d", "m_off", "l_off", "l_n", "eps_off", "z_off")]


class PmDesc(ctypes.Structure):
    _fields_ = [("C", ctypes.c_int32), ("ld_ml", ctypes.c_int32), ("ld_eps", ctypes.c_int32),
                ("ld_z", ctypes.c_int32), ("comp", Component * MAX_COMPONENTS)]


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("mvae_oracle.c", "mvae_oracle_impl.h")]
    src.append(os.path.join(_HERE, "..", "include", "mvae_b200.h"))
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        assert _lib.oracle_sizeof_pm_desc() == ctypes.sizeof(PmDesc)
    return _lib


def parse_signature(sig: str):
    """'2h3,s2,e2' -> ([types], [dims]) following the reference grammar (mt/mvae/utils.py:78-140)."""
    types, dims = [], []
    for tok in sig.lower().strip().split(","):
        tok = tok.strip().split("-")[0]
        i = 0
        while i < len(tok) and tok[i].isdigit():
            i += 1
        mult = int(tok[:i]) if i else 1
        j = i
        while j < len(tok) and tok[j].isalpha():
            j += 1
        letter, dim = tok[i:j], int(tok[j:])
        for _ in range(mult):
            types.append(TYPE_OF_LETTER[letter])
            dims.append(dim)
    return types, dims


def make_desc(sig_or_types, dims=None, scalar_parametrization=False) -> PmDesc:
    if isinstance(sig_or_types, str):
        types, dims = parse_signature(sig_or_types)
    else:
        types = list(sig_or_types)
    C = len(types)
    d = PmDesc()
    rc = lib().oracle_pm_desc_init(ctypes.byref(d), C, (ctypes.c_int32 * C)(*types), (ctypes.c_int32 * C)(*dims),
                                   int(scalar_parametrization))
    if rc != 0:
        raise ValueError("bad product-manifold signature")
    return d


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "_f64", ctypes.c_double
    if dtype == np.float32:
        return "_f32", ctypes.c_float
    raise TypeError(dtype)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def pm_forward(desc: PmDesc, ml, eps, radius, want=("z", "kl")):
    """Returns dict with any of z, kl, mu, sigma, u, logq, logp."""
    ml = np.ascontiguousarray(ml)
    dt = ml.dtype
    sfx, _ = _sfx(dt)
    eps = np.ascontiguousarray(eps, dtype=dt)
    radius = np.ascontiguousarray(radius, dtype=dt)
    B = ml.shape[0]
    assert ml.shape[1] == desc.ld_ml and eps.shape[1] == desc.ld_eps and radius.shape[0] == desc.C
    shapes = {"z": desc.ld_z, "kl": desc.C, "mu": desc.ld_z, "sigma": desc.ld_eps, "u": desc.ld_z,
              "logq": desc.C, "logp": desc.C}
    out = {k: (np.zeros((B, shapes[k]), dtype=dt) if k in want else None) for k in shapes}
    fn = getattr(lib(), "oracle_pm_forward" + sfx)
    fn.restype = None
    fn(ctypes.byref(desc), ctypes.c_int64(B), _p(ml), _p(eps), _p(radius), _p(out["z"]), _p(out["kl"]),
       _p(out["mu"]), _p(out["sigma"]), _p(out["u"]), _p(out["logq"]), _p(out["logp"]))
    return {k: v for k, v in out.items() if v is not None}


def pm_backward(desc: PmDesc, ml, eps, radius, gz, gkl=None, gkl_scalar=1.0):
    """Returns (gml [B, ld_ml], gradius [C] float64)."""
    ml = np.ascontiguousarray(ml)
    dt = ml.dtype
    sfx, creal = _sfx(dt)
    eps = np.ascontiguousarray(eps, dtype=dt)
    radius = np.ascontiguousarray(radius, dtype=dt)
    gz = np.ascontiguousarray(gz, dtype=dt)
    gkl = None if gkl is None else np.ascontiguousarray(gkl, dtype=dt)
    B = ml.shape[0]
    gml = np.zeros_like(ml)
    gR = np.zeros(desc.C, dtype=np.float64)
    fn = getattr(lib(), "oracle_pm_backward" + sfx)
    fn.restype = None
    fn(ctypes.byref(desc), ctypes.c_int64(B), _p(ml), _p(eps), _p(radius), _p(gz), _p(gkl), creal(gkl_scalar),
       _p(gml), _p(gR))
    return gml, gR


def recon(kind, logits, x, want_grad=False, gscale=1.0):
    """kind 'bce' | 'nll' -> (rowsum [B], glogits or None)."""
    logits = np.ascontiguousarray(logits)
    dt = logits.dtype
    sfx, creal = _sfx(dt)
    x = np.ascontiguousarray(x, dtype=dt)
    B, D = logits.shape
    rs = np.zeros(B, dtype=dt)
    g = np.zeros_like(logits) if want_grad else None
    fn = getattr(lib(), "oracle_recon" + sfx)
    fn.restype = None
    fn(0 if kind == "bce" else 1, ctypes.c_int64(B), D, _p(logits), _p(x), _p(rs), _p(g), creal(gscale))
    return rs, g


def elbo(bce, kl, beta):
    """-> float64 array [bce_sum, kl_sum, elbo, kl_c...] (stats.py:144-202)."""
    bce = np.ascontiguousarray(bce)
    dt = bce.dtype
    sfx, creal = _sfx(dt)
    kl = np.ascontiguousarray(kl, dtype=dt)
    B, C = kl.shape
    out = np.zeros(3 + C, dtype=np.float64)
    fn = getattr(lib(), "oracle_elbo" + sfx)
    fn.restype = None
    fn(ctypes.c_int64(B), C, _p(bce), _p(kl), creal(beta), _p(out))
    return out


def linear(x, W, b, relu=False, use_blas=True):
    """y = x W^T + b (optionally relu).  use_blas routes the contraction through numpy (threaded BLAS) — same
    arithmetic, used by the timed CPU baseline; the plain-C loop is kept for cross-checking."""
    if use_blas:
        y = x @ W.T
        if b is not None:
            y += b
        if relu:
            np.maximum(y, 0, out=y)
        return y
    x = np.ascontiguousarray(x)
    dt = x.dtype
    sfx, _ = _sfx(dt)
    W = np.ascontiguousarray(W, dtype=dt)
    b = None if b is None else np.ascontiguousarray(b, dtype=dt)
    B, K = x.shape
    N = W.shape[0]
    y = np.zeros((B, N), dtype=dt)
    fn = getattr(lib(), "oracle_linear" + sfx)
    fn.restype = None
    fn(ctypes.c_int64(B), K, N, _p(x), _p(W), _p(b), _p(y), int(relu))
    return y


class OracleVAE:
    """Whole hot path (ModelVAE.forward + compute_batch_stats + backward, vae.py:69-80,125-160) on the CPU:
    numpy/BLAS for the three dense layers (ffnn_vae.py:42-60) + the C oracle for everything per-sample.
    Parameters are a dict with the reference's state_dict names."""

    def __init__(self, sig, in_dim, h_dim, recon_kind="bce", scalar_parametrization=False):
        self.types, self.dims = parse_signature(sig)
        self.desc = make_desc(self.types, self.dims, scalar_parametrization)
        self.in_dim, self.h_dim, self.recon_kind = in_dim, h_dim, recon_kind
        self.C = len(self.types)

    def heads_matrix(self, params):
        """Stack fc_mean / fc_logvar of every component in packed-ml order -> (W [P,H], b [P])."""
        Ws, bs = [], []
        for i in range(self.C):
            Ws += [params[f"components.{i}.fc_mean.weight"], params[f"components.{i}.fc_logvar.weight"]]
            bs += [params[f"components.{i}.fc_mean.bias"], params[f"components.{i}.fc_logvar.bias"]]
        return np.concatenate(Ws, 0), np.concatenate(bs, 0)

    def radii(self, params, dtype):
        r = np.ones(self.C, dtype=dtype)
        for i in range(self.C):
            for nm in ("_nradius", "_pradius", "_curvature"):  # 'u': the slot holds the raw curvature
                k = f"components.{i}.{nm}"
                if k in params:
                    r[i] = params[k]
        return r

    def log_likelihood(self, params, x, eps):
        """ModelVAE.log_likelihood (vae.py:82-123) with the draws supplied: eps [n, B, sum(n_i)] ->
        dict(log_p_x [B], mi [B], cov_norm, log_q [n,B], log_p [n,B], log_p_x_z [n,B], z [n,B,Sd])."""
        dt = x.dtype
        n, B = eps.shape[0], x.shape[0]
        Wh, bh = self.heads_matrix(params)
        R = self.radii(params, dt)
        h = linear(x, params["fc_e0.weight"], params["fc_e0.bias"], relu=True)      # vae.py:93
        ml = linear(h, Wh, bh)
        lq = np.zeros((n, B), dtype=dt)
        lp = np.zeros((n, B), dtype=dt)
        lpxz = np.zeros((n, B), dtype=dt)
        zs = np.zeros((n, B, self.desc.ld_z), dtype=dt)
        for s in range(n):
            f = pm_forward(self.desc, ml, eps[s], R, want=("z", "logq", "logp"))    # :97-104 rsample_log_probs
            lq[s], lp[s], zs[s] = f["logq"].sum(-1), f["logp"].sum(-1), f["z"]
            dd = linear(f["z"], params["fc_d0.weight"], params["fc_d0.bias"], relu=True)
            logits = linear(dd, params["fc_logits.weight"], params["fc_logits.bias"])  # :107
            lpxz[s] = -recon(self.recon_kind, logits, x)[0]                          # :109

        def lse(a):
            m = a.max(0)
            return m + np.log(np.exp(a - m).sum(0))

        log_p_x = lse(lpxz + lp - lq) - np.log(n)                                    # :113-114
        mi = lse(lq - lp) - np.log(n)                                                # :117
        xc = x - x.mean(0, keepdims=True)                                            # :119-121
        zc = zs - zs.mean(1, keepdims=True)
        cov = np.einsum("bd,sbj->dj", xc, zc) / n
        return {"log_p_x": log_p_x, "mi": mi, "cov_norm": np.sqrt((cov * cov).sum()), "log_q": lq, "log_p": lp,
                "log_p_x_z": lpxz, "z": zs}

    def step(self, params, x, eps, beta=1.0, backward=True, relu_decisions=None):
        """relu_decisions = {"h": bool [B,H], "dd": bool [B,H]} (optional): the backward pass uses these relu masks
        instead of its own (h > 0), (dd > 0).  relu'(0) is a convention and a pre-activation within float32 rounding
        of zero may fall on either side in a float32 implementation; the parity tests pass the decisions the device
        took, after checking that every differing unit really sits at the kink (|pre-activation| ~ 0).  The
        pre-activations are returned as "h_pre" / "dd_pre" for that check.  Forward values are unaffected."""
        dt = x.dtype
        Wh, bh = self.heads_matrix(params)
        R = self.radii(params, dt)
        h_pre = linear(x, params["fc_e0.weight"], params["fc_e0.bias"])
        h = np.maximum(h_pre, 0)
        ml = linear(h, Wh, bh)
        f = pm_forward(self.desc, ml, eps, R, want=("z", "kl", "mu", "sigma"))
        dd_pre = linear(f["z"], params["fc_d0.weight"], params["fc_d0.bias"])
        dd = np.maximum(dd_pre, 0)
        logits = linear(dd, params["fc_logits.weight"], params["fc_logits.bias"])
        bce, glogits = recon(self.recon_kind, logits, x, want_grad=backward)
        stats = elbo(bce, f["kl"], beta)
        out = {"h": h, "ml": ml, "z": f["z"], "kl": f["kl"], "mu": f["mu"], "sigma": f["sigma"], "logits": logits,
               "bce": bce, "h_pre": h_pre, "dd_pre": dd_pre, "bce_sum": stats[0], "kl_sum": stats[1], "elbo": stats[2], "kl_comp": stats[3:]}
        if not backward:
            return out
        g = {}
        # loss = -elbo = sum bce + beta * sum kl
        g["fc_logits.weight"] = glogits.T @ dd
        g["fc_logits.bias"] = glogits.sum(0)
        h_on = (h > 0) if relu_decisions is None else np.asarray(relu_decisions["h"], dtype=bool)
        dd_on = (dd > 0) if relu_decisions is None else np.asarray(relu_decisions["dd"], dtype=bool)
        gdd = (glogits @ params["fc_logits.weight"]) * dd_on
        g["fc_d0.weight"] = gdd.T @ f["z"]
        g["fc_d0.bias"] = gdd.sum(0)
        gz = gdd @ params["fc_d0.weight"]
        gml, gR = pm_backward(self.desc, ml, eps, R, gz, None, beta)
        gWh = gml.T @ h
        gbh = gml.sum(0)
        row = 0
        for i in range(self.C):
            c = self.desc.comp[i]
            g[f"components.{i}.fc_mean.weight"] = gWh[row:row + c.n]
            g[f"components.{i}.fc_mean.bias"] = gbh[row:row + c.n]
            row += c.n
            g[f"components.{i}.fc_logvar.weight"] = gWh[row:row + c.l_n]
            g[f"components.{i}.fc_logvar.bias"] = gbh[row:row + c.l_n]
            row += c.l_n
            for nm in ("_nradius", "_pradius", "_curvature"):
                if f"components.{i}.{nm}" in params:
                    g[f"components.{i}.{nm}"] = np.asarray(gR[i])
        gh = (gml @ Wh) * h_on
        g["fc_e0.weight"] = gh.T @ x
        g["fc_e0.bias"] = gh.sum(0)
        out["grads"] = g
        out["gz"] = gz
        out["gml"] = gml
        return out


class OracleTrainer:
    """k train steps of the reference on the CPU: OracleVAE.step (vae.py:149-160) followed by the update of
    Trainer.build_optimizer + CurvatureOptimizer.step (train.py:327-360, utils.py:174-180): torch.optim.Adam(lr) on
    every network parameter, the 2-norm clip of the "curvature"-named gradients (vae.py:161-163), SGD(lr=1e-4) on
    `_nradius` / `_pradius` / `_curvature` when the curvature optimizers step.  All arithmetic in the dtype of `params`
    (float64 for the parity tests).  Under data parallelism the global batch is simply passed as one batch (the ELBO
    and its gradients are sums over samples)."""

    def __init__(self, ovae: OracleVAE, params: dict, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, curvature_lr=1e-4,
                 curvature_step=True, fixed_curvature=False):
        self.ovae = ovae
        self.params = {k: np.array(v, copy=True) for k, v in params.items()}
        self.lr, self.betas, self.eps = lr, betas, eps
        self.curvature_lr = curvature_lr if (curvature_step and not fixed_curvature) else 0.0
        self.net_keys = [k for k in self.params if not any(s in k for s in ("radius", "curvature"))]
        self.curv_keys = [k for k in self.params if any(s in k for s in ("radius", "curvature"))]
        self.m = {k: np.zeros_like(self.params[k]) for k in self.net_keys}
        self.v = {k: np.zeros_like(self.params[k]) for k in self.net_keys}
        self.t = 0

    def step(self, x, eps, beta=1.0, relu_decisions=None):
        out = self.ovae.step(self.params, x, eps, beta=beta, relu_decisions=relu_decisions)
        g = out["grads"]
        clip_keys = [k for k in self.curv_keys if "curvature" in k and k in g]
        if clip_keys:
            total = float(np.sqrt(sum(float(g[k]) ** 2 for k in clip_keys)))
            coef = min(1.0, 1.0 / (total + 1e-6))
            for k in clip_keys:
                g[k] = np.asarray(g[k]) * coef
        self.t += 1
        b1, b2 = self.betas
        bc1, bc2 = 1.0 - b1 ** self.t, 1.0 - b2 ** self.t
        for k in self.net_keys:
            gk = np.asarray(g[k], dtype=self.params[k].dtype)
            self.m[k] = b1 * self.m[k] + (1.0 - b1) * gk
            self.v[k] = b2 * self.v[k] + (1.0 - b2) * gk * gk
            self.params[k] = self.params[k] - (self.lr / bc1) * self.m[k] / (np.sqrt(self.v[k]) / np.sqrt(bc2) + self.eps)
        if self.curvature_lr:
            for k in self.curv_keys:
                if k in g:
                    self.params[k] = self.params[k] - self.curvature_lr * np.asarray(g[k], dtype=self.params[k].dtype)
        return out


# ------------------------------------------------------------------------------------------ convolutional VAE
# ConvolutionalVAE (mt/mvae/models/conv_vae.py:28-79): three nn.Conv2d(k=4, s=2, p=1) + relu -> flatten (8192) ->
# components -> nn.Linear(z, 2048) + relu -> view(128, 4, 4) -> three nn.ConvTranspose2d(k=4, s=2, p=1) (relu between).
# The layer arithmetic is torch's (a dependency of the reference, not code under /root/reference): restated here from
# its documented definition and pinned by tests/golden/conv_*.npz, which the reference's own model produced.
def _im2col_k4s2p1(x):
    """x [B, C, H, W] -> cols [B, H/2, W/2, C, 4, 4]: cols[b,oy,ox,c,ky,kx] = xpad[b,c,2oy+ky,2ox+kx] (pad 1)."""
    B, C, H, W = x.shape
    OH, OW = H // 2, W // 2
    xp = np.zeros((B, C, H + 2, W + 2), dtype=x.dtype)
    xp[:, :, 1:-1, 1:-1] = x
    cols = np.empty((B, OH, OW, C, 4, 4), dtype=x.dtype)
    for ky in range(4):
        for kx in range(4):
            cols[:, :, :, :, ky, kx] = xp[:, :, ky:ky + 2 * OH:2, kx:kx + 2 * OW:2].transpose(0, 2, 3, 1)
    return cols


def _col2im_k4s2p1(cols):
    """Adjoint of _im2col_k4s2p1: cols [B, OH, OW, C, 4, 4] -> x [B, C, 2 OH, 2 OW] (scatter-add)."""
    B, OH, OW, C = cols.shape[:4]
    xp = np.zeros((B, C, 2 * OH + 2, 2 * OW + 2), dtype=cols.dtype)
    for ky in range(4):
        for kx in range(4):
            xp[:, :, ky:ky + 2 * OH:2, kx:kx + 2 * OW:2] += cols[:, :, :, :, ky, kx].transpose(0, 3, 1, 2)
    return xp[:, :, 1:-1, 1:-1]


def conv2d_k4s2p1(x, w, b):
    """torch.nn.Conv2d(Ci, Co, 4, 2, 1): x [B,Ci,H,W], w [Co,Ci,4,4] -> [B,Co,H/2,W/2]."""
    B, Ci, H, W = x.shape
    Co = w.shape[0]
    cols = _im2col_k4s2p1(x).reshape(B * (H // 2) * (W // 2), Ci * 16)
    y = cols @ w.reshape(Co, Ci * 16).T + b
    return y.reshape(B, H // 2, W // 2, Co).transpose(0, 3, 1, 2)


def conv2d_k4s2p1_bwd(x, w, gy):
    B, Ci, H, W = x.shape
    Co = w.shape[0]
    OH, OW = H // 2, W // 2
    cols = _im2col_k4s2p1(x).reshape(B * OH * OW, Ci * 16)
    g = gy.transpose(0, 2, 3, 1).reshape(B * OH * OW, Co)
    gw = (g.T @ cols).reshape(w.shape)
    gx = _col2im_k4s2p1((g @ w.reshape(Co, Ci * 16)).reshape(B, OH, OW, Ci, 4, 4))
    return gx, gw, g.sum(0)


def convT2d_k4s2p1(x, w, b):
    """torch.nn.ConvTranspose2d(Ci, Co, 4, 2, 1): x [B,Ci,H,W], w [Ci,Co,4,4] -> [B,Co,2H,2W]:
    out[b,co,2iy-1+ky,2ix-1+kx] += x[b,ci,iy,ix] w[ci,co,ky,kx]."""
    B, Ci, H, W = x.shape
    Co = w.shape[1]
    xm = x.transpose(0, 2, 3, 1).reshape(B * H * W, Ci)
    cols = (xm @ w.reshape(Ci, Co * 16)).reshape(B, H, W, Co, 4, 4)
    return _col2im_k4s2p1(cols) + b[None, :, None, None]


def convT2d_k4s2p1_bwd(x, w, gy):
    B, Ci, H, W = x.shape
    Co = w.shape[1]
    xm = x.transpose(0, 2, 3, 1).reshape(B * H * W, Ci)
    gcols = _im2col_k4s2p1(gy).reshape(B * H * W, Co * 16)
    gx = (gcols @ w.reshape(Ci, Co * 16).T).reshape(B, H, W, Ci).transpose(0, 3, 1, 2)
    gw = (xm.T @ gcols).reshape(w.shape)
    return gx, gw, gy.sum((0, 2, 3))


class OracleConvVAE(OracleVAE):
    """ConvolutionalVAE (conv_vae.py:28-79) + the same latent path / ELBO / backward as OracleVAE.  Inputs are
    [B, 3072] rows in (c, y, x) order (conv_vae.py:61), h_dim = 8192 (the flattened 512 x 4 x 4 encoder output)."""

    IMG = (3, 32, 32)

    def __init__(self, sig, recon_kind="bce", scalar_parametrization=False):
        super().__init__(sig, 3072, 8192, recon_kind, scalar_parametrization)

    def step(self, params, x, eps, beta=1.0, backward=True, relu_decisions=None):
        assert relu_decisions is None, "the convolutional oracle takes its own relu decisions"
        p = params
        B = x.shape[0]
        R = self.radii(p, x.dtype)
        Wh, bh = self.heads_matrix(p)
        a0 = x.reshape((B,) + self.IMG)
        pre, act = [], [a0]
        for nm in ("e0", "e1", "e2"):                                            # conv_vae.py:62-64
            pre.append(conv2d_k4s2p1(act[-1], p[nm + ".weight"], p[nm + ".bias"]))
            act.append(np.maximum(pre[-1], 0))
        h = act[-1].reshape(B, -1)                                               # :65 flatten (c, y, x)
        ml = linear(h, Wh, bh)
        f = pm_forward(self.desc, ml, eps, R, want=("z", "kl", "mu", "sigma"))
        dd_pre = linear(f["z"], p["d0.weight"], p["d0.bias"])                    # :72
        dact = [np.maximum(dd_pre, 0).reshape(B, 128, 4, 4)]                     # :73
        dpre = []
        for nm in ("d1", "d2"):                                                  # :74-75
            dpre.append(convT2d_k4s2p1(dact[-1], p[nm + ".weight"], p[nm + ".bias"]))
            dact.append(np.maximum(dpre[-1], 0))
        logits = convT2d_k4s2p1(dact[-1], p["d3.weight"], p["d3.bias"]).reshape(B, -1)   # :76-78
        bce, glogits = recon(self.recon_kind, logits, x, want_grad=backward)
        stats = elbo(bce, f["kl"], beta)
        out = {"h": h, "ml": ml, "z": f["z"], "kl": f["kl"], "mu": f["mu"], "sigma": f["sigma"], "logits": logits,
               "bce": bce, "bce_sum": stats[0], "kl_sum": stats[1], "elbo": stats[2], "kl_comp": stats[3:]}
        if not backward:
            return out
        g = {}
        gy = glogits.reshape((B,) + self.IMG)
        gy, g["d3.weight"], g["d3.bias"] = convT2d_k4s2p1_bwd(dact[2], p["d3.weight"], gy)
        for i, nm in ((1, "d2"), (0, "d1")):
            gy = gy * (dact[i + 1] > 0)
            gy, g[nm + ".weight"], g[nm + ".bias"] = convT2d_k4s2p1_bwd(dact[i], p[nm + ".weight"], gy)
        gdd = gy.reshape(B, -1) * (dd_pre > 0)
        g["d0.weight"], g["d0.bias"] = gdd.T @ f["z"], gdd.sum(0)
        gz = gdd @ p["d0.weight"]
        gml, gR = pm_backward(self.desc, ml, eps, R, gz, None, beta)
        gWh, gbh = gml.T @ h, gml.sum(0)
        row = 0
        for i in range(self.C):
            c = self.desc.comp[i]
            g[f"components.{i}.fc_mean.weight"], g[f"components.{i}.fc_mean.bias"] = gWh[row:row + c.n], gbh[row:row + c.n]
            row += c.n
            g[f"components.{i}.fc_logvar.weight"] = gWh[row:row + c.l_n]
            g[f"components.{i}.fc_logvar.bias"] = gbh[row:row + c.l_n]
            row += c.l_n
            for nm in ("_nradius", "_pradius", "_curvature"):
                if f"components.{i}.{nm}" in p:
                    g[f"components.{i}.{nm}"] = np.asarray(gR[i])
        gy = ((gml @ Wh) * (h > 0)).reshape(B, 512, 4, 4)
        for i, nm in ((2, "e2"), (1, "e1"), (0, "e0")):
            gy, g[nm + ".weight"], g[nm + ".bias"] = conv2d_k4s2p1_bwd(act[i], p[nm + ".weight"], gy)
            if i > 0:
                gy = gy * (act[i] > 0)
        out["grads"] = g
        out["gz"], out["gml"] = gz, gml
        return out
