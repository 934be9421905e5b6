"""Restatement of the part of ``geoopt==0.1.0``'s ``geoopt.manifolds.poincare.math`` that the
reference calls (reference call sites: mt/mvae/ops/poincare.py:100,117,121,129,137,145,149,154,163,181;
mt/mvae/ops/spherical_projected.py:113,137; tests use parallel_transport, gyration, lambda_x).

TEST INFRASTRUCTURE ONLY. geoopt is a third-party dependency that is not vendored under
/root/reference and cannot be installed here (no network). Formulas follow the published
Poincare-ball gyrovector calculus (Ganea et al. 2018) as shipped in geoopt 0.1.0; the guard
constants (MIN_NORM, artanh / tanh clamps) are from memory => "parity unpinned" where they bind.
"""
import torch

MIN_NORM = 1e-15


def tanh(x):
    return x.clamp(-15, 15).tanh()


class _Artanh(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x):
        x = x.clamp(-1 + 1e-5, 1 - 1e-5)
        ctx.save_for_backward(x)
        return (torch.log1p(x) - torch.log1p(-x)) * 0.5

    @staticmethod
    def backward(ctx, g):
        x, = ctx.saved_tensors
        return g / (1 - x**2)


def artanh(x):
    return _Artanh.apply(x)


def _sq(x, dim=-1, keepdim=True):
    return x.pow(2).sum(dim=dim, keepdim=keepdim)


def lambda_x(x, c=1.0, keepdim=False, dim=-1):
    return 2 / (1 - c * _sq(x, dim, keepdim)).clamp_min(MIN_NORM)


def mobius_add(x, y, c=1.0, dim=-1):
    x2 = _sq(x, dim)
    y2 = _sq(y, dim)
    xy = (x * y).sum(dim=dim, keepdim=True)
    num = (1 + 2 * c * xy + c * y2) * x + (1 - c * x2) * y
    denom = 1 + 2 * c * xy + c**2 * x2 * y2
    return num / denom.clamp_min(MIN_NORM)


def expmap(x, u, c=1.0, dim=-1):
    sqrt_c = c**0.5
    u_norm = u.norm(dim=dim, p=2, keepdim=True).clamp_min(MIN_NORM)
    second = tanh(sqrt_c / 2 * lambda_x(x, c=c, keepdim=True, dim=dim) * u_norm) * u / (sqrt_c * u_norm)
    return mobius_add(x, second, c=c, dim=dim)


def expmap0(u, c=1.0, dim=-1):
    sqrt_c = c**0.5
    u_norm = u.norm(dim=dim, p=2, keepdim=True).clamp_min(MIN_NORM)
    return tanh(sqrt_c * u_norm) * u / (sqrt_c * u_norm)


def logmap(x, y, c=1.0, dim=-1):
    sub = mobius_add(-x, y, c=c, dim=dim)
    sub_norm = sub.norm(dim=dim, p=2, keepdim=True).clamp_min(MIN_NORM)
    lam = lambda_x(x, c=c, keepdim=True, dim=dim)
    sqrt_c = c**0.5
    return 2 / sqrt_c / lam * artanh(sqrt_c * sub_norm) * sub / sub_norm


def logmap0(y, c=1.0, dim=-1):
    sqrt_c = c**0.5
    y_norm = y.norm(dim=dim, p=2, keepdim=True).clamp_min(MIN_NORM)
    return y / y_norm / sqrt_c * artanh(sqrt_c * y_norm)


def gyration(a, b, u, c=1.0, dim=-1):
    a2 = _sq(a, dim)
    b2 = _sq(b, dim)
    ab = (a * b).sum(dim=dim, keepdim=True)
    au = (a * u).sum(dim=dim, keepdim=True)
    bu = (b * u).sum(dim=dim, keepdim=True)
    c2 = c**2
    A = -c2 * au * b2 + c * bu + 2 * c2 * ab * bu
    B = -c2 * bu * a2 - c * au
    D = 1 + 2 * c * ab + c2 * a2 * b2
    return u + 2 * (A * a + B * b) / D.clamp_min(MIN_NORM)


def parallel_transport(x, y, v, c=1.0, dim=-1):
    return gyration(y, -x, v, c=c, dim=dim) * lambda_x(x, c=c, keepdim=True, dim=dim) / lambda_x(
        y, c=c, keepdim=True, dim=dim)


def parallel_transport0(y, v, c=1.0, dim=-1):
    return v * (1 - c * _sq(y, dim)).clamp_min(MIN_NORM)


def parallel_transport0back(x, v, c=1.0, dim=-1):
    return v / (1 - c * _sq(x, dim)).clamp_min(MIN_NORM)


def dist(x, y, c=1.0, keepdim=False, dim=-1):
    sqrt_c = c**0.5
    d = artanh(sqrt_c * mobius_add(-x, y, c=c, dim=dim).norm(dim=dim, p=2, keepdim=keepdim))
    return d * 2 / sqrt_c


def project(x, c=1.0, dim=-1):
    norm = x.norm(dim=dim, keepdim=True, p=2).clamp_min(MIN_NORM)
    maxnorm = (1 - 1e-3) / (c**0.5)
    return torch.where(norm > maxnorm, x / norm * maxnorm, x)
