"""Host-side mirror of the reference's `Manifold` interface (mt/mvae/ops/manifold.py:22-75) and its concrete
manifolds (mt/mvae/ops/{hyperbolics.py:26-55, spherical.py:26-55, poincare.py:28-89, euclidean.py:24-59,
spherical_projected.py:28-92}).

Every method runs a device kernel of libmvae_b200.so through mvae_manifold_op (include/mvae_b200.h); inputs are
float32 CUDA tensors of shape [..., dim].  These standalone ops are forward-only (no autograd): gradients of the
training path are produced by the fused product-manifold backward kernel, not by differentiating these calls.
"""
from typing import Any, Callable, Tuple

import torch
from torch import Tensor

from . import _lib as L
from . import ops


class Manifold:
    """mt/mvae/ops/manifold.py:22-60."""
    kind: int = -1

    def __init__(self, n: int = None) -> None:
        self._n = n  # true dimension, inferred from the last axis of the arguments when None

    # -- helpers
    def _radius_param(self):
        return None

    def _true_dim(self, width: int, ambient: bool) -> int:
        if self.kind in (L.HYPERBOLOID, L.SPHERE) and ambient:
            return width - 1
        return width

    def _op(self, op: int, x: Tensor, y: Tensor = None, ambient_x: bool = True) -> Tensor:
        n = self._true_dim(x.shape[-1], ambient_x)
        rp = self._radius_param()
        return ops.manifold_op(op, self.kind, n, x, y, None if rp is None else rp.detach().reshape(1).float())

    # -- interface
    def exp_map_mu0(self, x: Tensor) -> Tensor:
        return self._op(L.OP_EXP_MAP_MU0, x, ambient_x=False)

    def inverse_exp_map_mu0(self, x: Tensor) -> Tensor:
        return self._op(L.OP_INV_EXP_MAP_MU0, x)

    def exp_map(self, x: Tensor, at_point: Tensor) -> Tensor:
        return self._op(L.OP_EXP_MAP, x, at_point)

    def inverse_exp_map(self, x: Tensor, at_point: Tensor) -> Tensor:
        return self._op(L.OP_INV_EXP_MAP, x, at_point)

    def parallel_transport_mu0(self, x: Tensor, dst: Tensor) -> Tensor:
        return self._op(L.OP_PT_MU0, x, dst)

    def inverse_parallel_transport_mu0(self, x: Tensor, src: Tensor) -> Tensor:
        return self._op(L.OP_INV_PT_MU0, x, src)

    def distance(self, x: Tensor, y: Tensor) -> Tensor:
        """Geodesic distance [..., 1] (poincare.py:96-105; tests/mvae/ops/test_hyperbolics.py:46, test_spherical.py:45,
        test_euclidean.py:41 for the other models)."""
        return self._op(L.OP_DISTANCE, x, y)

    def mu_0(self, shape: torch.Size, **kwargs: Any) -> Tensor:
        raise NotImplementedError

    def sample_projection_mu0(self, x: Tensor, at_point: Tensor) -> Tuple[Tensor, Tuple[Tensor, Tensor]]:
        n = x.shape[-1]
        rp = self._radius_param()
        ones = torch.ones_like(x)
        z, u, v = ops.wn_rsample(self.kind, n, at_point.reshape(-1, at_point.shape[-1]), ones.reshape(-1, n),
                                 x.reshape(-1, n).contiguous(), None if rp is None else rp.detach().reshape(1).float())
        lead = x.shape[:-1]
        return z.reshape(*lead, -1), (u.reshape(*lead, -1), x)

    def inverse_sample_projection_mu0(self, x_proj: Tensor, at_point: Tensor) -> Tuple[Tensor, Tensor]:
        u = self.inverse_exp_map(x_proj, at_point)
        v = self.inverse_parallel_transport_mu0(u, at_point)
        if self.kind in (L.HYPERBOLOID, L.SPHERE):
            v = v[..., 1:]
        elif self.kind in (L.POINCARE, L.PROJ_SPHERE):
            # poincare.py:160-164 / spherical_projected.py:182-186: v_ * lambda_x; the inverse PT kernel returns
            # v_ * lambda_x / 2
            v = 2.0 * v
        return u, v

    def logdet(self, mu: Tensor, std: Tensor, z: Tensor, data: Tuple[Tensor, ...]) -> Tensor:
        raise NotImplementedError

    @property
    def radius(self) -> Tensor:
        raise NotImplementedError

    @property
    def curvature(self) -> Tensor:
        raise NotImplementedError


class RadiusManifold(Manifold):
    """mt/mvae/ops/manifold.py:63-75: radius = clamp(relu(param), 1e-8, 1e8)."""

    def __init__(self, radius: Callable[[], Tensor]):
        super().__init__()
        self._radius = radius

    def _radius_param(self):
        return self._radius()

    @property
    def radius(self) -> Tensor:
        return torch.clamp(torch.relu(self._radius()), min=1e-8, max=1e8)

    @property
    def curvature(self) -> Tensor:
        return 1. / self.radius.pow(2)


class Hyperboloid(RadiusManifold):
    """mt/mvae/ops/hyperbolics.py:26-55."""
    kind = L.HYPERBOLOID

    def mu_0(self, shape: torch.Size, **kwargs: Any) -> Tensor:
        out = torch.zeros(shape, **kwargs)
        out[..., 0] = self.radius.to(out.dtype)
        return out

    def logdet(self, mu: Tensor, std: Tensor, z: Tensor, data: Tuple[Tensor, ...]) -> Tensor:
        return self._op(L.OP_LOGDET, data[0]).squeeze(-1)

    def to_poincare(self, x: Tensor) -> Tensor:
        """lorentz_to_poincare (hyperbolics.py:151-152)."""
        return self._op(L.OP_TO_POINCARE, x)

    @property
    def curvature(self) -> Tensor:
        return -super().curvature


class Sphere(RadiusManifold):
    """mt/mvae/ops/spherical.py:26-55."""
    kind = L.SPHERE

    def mu_0(self, shape: torch.Size, **kwargs: Any) -> Tensor:
        out = torch.zeros(shape, **kwargs)
        out[..., 0] = self.radius.to(out.dtype)
        return out

    def logdet(self, mu: Tensor, std: Tensor, z: Tensor, data: Tuple[Tensor, ...]) -> Tensor:
        return self._op(L.OP_LOGDET, data[0]).squeeze(-1)

    def to_projected(self, x: Tensor) -> Tensor:
        """spherical_to_projected (spherical.py:132-133)."""
        return self._op(L.OP_TO_POINCARE, x)


class PoincareBall(RadiusManifold):
    """mt/mvae/ops/poincare.py:28-89 (geoopt 0.1.0 math restated in the kernels)."""
    kind = L.POINCARE

    def mu_0(self, shape: torch.Size, **kwargs: Any) -> Tensor:
        return torch.zeros(shape, **kwargs)

    def mobius_add(self, x: Tensor, y: Tensor) -> Tensor:
        return self._op(L.OP_MOBIUS_ADD, x, y)

    def mobius_scalar_mul(self, r: Tensor, x: Tensor) -> Tensor:
        """r (x)_c x — no call site in the reference: parity unpinned."""
        return self._op(L.OP_MOBIUS_SCALAR_MUL, x, r.reshape(*x.shape[:-1], 1).float().contiguous())

    def to_lorentz(self, x: Tensor) -> Tensor:
        """poincare_to_lorentz (poincare.py:167-170)."""
        return self._op(L.OP_FROM_POINCARE, x)

    def logdet(self, mu: Tensor, std: Tensor, z: Tensor, data: Tuple[Tensor, ...]) -> Tensor:
        # poincare.py:84-89: log-det through the Lorentz model; equals (n-1)(log R + log(sinh r / r)), r = dist(mu, z)/R
        n = z.shape[-1]
        r = self.distance(mu, z).squeeze(-1) / self.radius
        return (n - 1) * (torch.log(self.radius) + torch.log(torch.sinh(r) / r))

    @property
    def curvature(self) -> Tensor:
        return -super().curvature


class StereographicallyProjectedSphere(RadiusManifold):
    """mt/mvae/ops/spherical_projected.py:28-92 ('d': the sphere of radius R in stereographic coordinates; geoopt's
    mobius_add at c = -1/R^2)."""
    kind = L.PROJ_SPHERE

    def mu_0(self, shape: torch.Size, **kwargs: Any) -> Tensor:
        return torch.zeros(shape, **kwargs)

    def mobius_add(self, x: Tensor, y: Tensor) -> Tensor:
        """mob_add (spherical_projected.py:107-113)."""
        return self._op(L.OP_MOBIUS_ADD, x, y)

    def to_spherical(self, x: Tensor) -> Tensor:
        """projected_to_spherical (spherical_projected.py:190-195)."""
        return self._op(L.OP_FROM_POINCARE, x)

    def logdet(self, mu: Tensor, std: Tensor, z: Tensor, data: Tuple[Tensor, ...]) -> Tensor:
        # spherical_projected.py:58-92: S._logdet of the sphere's log map of z at mu
        lead = z.shape[:-1]
        mu_b = mu.expand(*lead, mu.shape[-1]).reshape(-1, mu.shape[-1]).contiguous()
        return self._op(L.OP_LOGDET, z.reshape(-1, z.shape[-1]).contiguous(), mu_b).reshape(lead)


class Universal(Manifold):
    """mt/mvae/ops/universal.py:27-86: the sign of the learnable curvature picks the manifold on every call —
    PoincareBall (kappa < -eps), StereographicallyProjectedSphere (kappa > eps) or Euclidean — with
    radius = relu(1 / sqrt(|kappa|)).  Standalone ops delegate to the chosen manifold (one host read of kappa, like the
    reference's `if`); the fused kernels take the same branch on the device (MVAE_UNIVERSAL)."""
    kind = L.UNIVERSAL

    def __init__(self, curvature: Callable[[], Tensor], eps: float = 1e-6) -> None:
        super().__init__()
        self._curvature = curvature
        self.eps = eps
        self._manifolds = {-1: PoincareBall(lambda: self.radius), 0: Euclidean(),
                           1: StereographicallyProjectedSphere(lambda: self.radius)}

    @property
    def curvature(self) -> Tensor:
        return self._curvature()

    @property
    def radius(self) -> Tensor:
        return torch.relu(1.0 / torch.sqrt(torch.clamp(self.curvature.detach().abs(), min=1e-9)))

    @property
    def _choice(self) -> int:
        k = float(self._curvature().detach())
        return -1 if k < -self.eps else (1 if k > self.eps else 0)

    @property
    def manifold(self) -> Manifold:
        return self._manifolds[self._choice]

    def _radius_param(self):
        return self.manifold._radius_param()

    def exp_map_mu0(self, x: Tensor) -> Tensor:
        return self.manifold.exp_map_mu0(x)

    def inverse_exp_map_mu0(self, x: Tensor) -> Tensor:
        return self.manifold.inverse_exp_map_mu0(x)

    def exp_map(self, x: Tensor, at_point: Tensor) -> Tensor:
        return self.manifold.exp_map(x, at_point)

    def inverse_exp_map(self, x: Tensor, at_point: Tensor) -> Tensor:
        return self.manifold.inverse_exp_map(x, at_point)

    def parallel_transport_mu0(self, x: Tensor, dst: Tensor) -> Tensor:
        return self.manifold.parallel_transport_mu0(x, dst)

    def inverse_parallel_transport_mu0(self, x: Tensor, src: Tensor) -> Tensor:
        return self.manifold.inverse_parallel_transport_mu0(x, src)

    def distance(self, x: Tensor, y: Tensor) -> Tensor:
        return self.manifold.distance(x, y)

    def mu_0(self, shape: torch.Size, **kwargs: Any) -> Tensor:
        return self.manifold.mu_0(shape, **kwargs)

    def sample_projection_mu0(self, x: Tensor, at_point: Tensor) -> Tuple[Tensor, Tuple[Tensor, Tensor]]:
        return self.manifold.sample_projection_mu0(x, at_point)

    def inverse_sample_projection_mu0(self, x_proj: Tensor, at_point: Tensor) -> Tuple[Tensor, Tensor]:
        return self.manifold.inverse_sample_projection_mu0(x_proj, at_point)

    def logdet(self, mu: Tensor, std: Tensor, z: Tensor, data: Tuple[Tensor, ...]) -> Tensor:
        return self.manifold.logdet(mu, std, z, data)


class Euclidean(Manifold):
    """mt/mvae/ops/euclidean.py:24-59 (note exp_map_mu0(x) = x/2)."""
    kind = L.EUCLIDEAN

    @property
    def radius(self) -> Tensor:
        return 0

    @property
    def curvature(self) -> Tensor:
        return 0

    def mu_0(self, shape: torch.Size, **kwargs: Any) -> Tensor:
        return torch.zeros(shape, **kwargs)

    def logdet(self, mu: Tensor, std: Tensor, z: Tensor, data: Tuple[Tensor, ...]) -> Tensor:
        return torch.zeros_like(mu)
