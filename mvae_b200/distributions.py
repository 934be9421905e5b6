"""Host-side mirror of the reference's distributions on the hot path:
`WrappedNormal` (mt/mvae/distributions/wrapped_normal.py:25-107) and `EuclideanNormal`
(mt/mvae/distributions/wrapped_distributions.py:23-42).  Sampling and densities run in the kernels behind
mvae_wn_rsample / mvae_wn_log_prob_from_parts / mvae_wn_log_prob (include/mvae_b200.h); forward-only."""
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L
from . import ops
from .manifolds import Manifold


class VaeDistribution:
    """wrapped_distributions.py:23-36."""

    def rsample_with_parts(self, shape: torch.Size = torch.Size()) -> Tuple[Tensor, Optional[Tuple[Tensor, ...]]]:
        z = self.rsample(shape)
        return z, None

    def log_prob_from_parts(self, z: Tensor, data: Optional[Tuple[Tensor, ...]]) -> Tensor:
        return self.log_prob(z)

    def rsample_log_prob(self, shape: torch.Size = torch.Size()) -> Tuple[Tensor, Tensor]:
        z, data = self.rsample_with_parts(shape)
        return z, self.log_prob_from_parts(z, data)


class EuclideanNormal(torch.distributions.Normal, VaeDistribution):
    """wrapped_distributions.py:39-42: Normal whose log_prob sums over the last axis."""

    def log_prob(self, value: Tensor) -> Tensor:
        return super().log_prob(value).sum(dim=-1)


class WrappedNormal(VaeDistribution):
    """WrappedNormal(loc, scale, manifold): loc [B, d] on the manifold, scale [B, n] or [B, 1] (isotropic)."""
    has_rsample = True

    def __init__(self, loc: Tensor, scale: Tensor, manifold: Manifold) -> None:
        self.dim = loc.shape[-1]
        tangent_dim = self.dim if manifold.kind in (L.POINCARE, L.PROJ_SPHERE, L.EUCLIDEAN) else self.dim - 1
        if scale.shape[-1] > 1 and scale.shape[-1] != tangent_dim:
            raise ValueError("Invalid scale dimension: neither isotropic nor elliptical.")
        if scale.shape[-1] == 1:
            scale = scale.expand(*scale.shape[:-1], tangent_dim)
        assert loc.shape[:-1] == scale.shape[:-1]
        self.loc = loc.contiguous()
        self.scale = scale.contiguous()
        self.manifold = manifold
        self.device = loc.device
        self._n = tangent_dim

    @property
    def mean(self) -> Tensor:
        return self.loc

    @property
    def stddev(self) -> Tensor:
        return self.scale

    def _radius(self):
        rp = self.manifold._radius_param()
        return None if rp is None else rp.detach().reshape(1).float()

    def rsample_with_parts(self, shape: torch.Size = torch.Size(), eps: Optional[Tensor] = None):
        lead = tuple(shape) + tuple(self.loc.shape[:-1])
        loc = self.loc.expand(*lead, self.dim).reshape(-1, self.dim).contiguous()
        scale = self.scale.expand(*lead, self._n).reshape(-1, self._n).contiguous()
        if eps is None:
            eps = torch.randn(loc.shape[0], self._n, device=self.device, dtype=torch.float32)
        z, u, v = ops.wn_rsample(self.manifold.kind, self._n, loc, scale, eps.reshape(-1, self._n).contiguous(),
                                 self._radius())
        return z.reshape(*lead, self.dim), (u.reshape(*lead, self.dim), v.reshape(*lead, self._n))

    def rsample(self, sample_shape: torch.Size = torch.Size()) -> Tensor:
        return self.rsample_with_parts(sample_shape)[0]

    def log_prob_from_parts(self, z: Tensor, data: Optional[Tuple[Tensor, ...]]) -> Tensor:
        if data is None:
            raise ValueError("Additional data cannot be empty for WrappedNormal.")
        lead = z.shape[:-1]
        loc = self.loc.expand(*lead, self.dim).reshape(-1, self.dim).contiguous()
        scale = self.scale.expand(*lead, self._n).reshape(-1, self._n).contiguous()
        lp = ops.wn_log_prob_from_parts(self.manifold.kind, self._n, loc, scale, z.reshape(-1, self.dim).contiguous(),
                                        data[0].reshape(-1, self.dim).contiguous(),
                                        data[1].reshape(-1, self._n).contiguous(), self._radius())
        return lp.reshape(lead)

    def log_prob(self, z: Tensor) -> Tensor:
        lead = z.shape[:-1]
        loc = self.loc.expand(*lead, self.dim).reshape(-1, self.dim).contiguous()
        scale = self.scale.expand(*lead, self._n).reshape(-1, self._n).contiguous()
        lp = ops.wn_log_prob(self.manifold.kind, self._n, loc, scale, z.reshape(-1, self.dim).contiguous(),
                             self._radius())
        return lp.reshape(lead)
