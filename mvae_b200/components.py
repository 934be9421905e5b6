"""Host-side mirror of the reference's latent `Component`s (mt/mvae/components/component.py:30-242) and of
`parse_components` (mt/mvae/utils.py:30-48,78-140).  Same attributes (dim, true_dim, mean_dim, manifold, fc_mean,
fc_logvar, _nradius / _pradius) and the same state_dict keys, so reference checkpoints load and the reference's
Trainer.build_optimizer ("nradius" / "pradius" name filters, train.py:327-360) keeps working.

In training the per-component work is NOT done here: FusedFeedForwardVAE runs all components of the product
manifold in one fused kernel.  `Component.forward` exists for API compatibility (evaluation, visualisation)."""
import re
from typing import List, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from . import _lib as L
from .distributions import EuclideanNormal, WrappedNormal
from .manifolds import (Euclidean, Hyperboloid, Manifold, PoincareBall, Sphere, StereographicallyProjectedSphere,
                        Universal)


class Component(torch.nn.Module):
    letter = "?"
    kind = -1

    def __init__(self, dim: int, fixed_curvature: bool) -> None:
        super().__init__()
        self.dim = dim
        self.fixed_curvature = fixed_curvature
        self.manifold: Manifold = None
        self.fc_mean: torch.nn.Linear = None
        self.fc_logvar: torch.nn.Linear = None
        self.scalar_parametrization = False

    def init_layers(self, in_dim: int, scalar_parametrization: bool) -> None:
        """component.py:47-56 (same construction order => same default initialisation under a given seed)."""
        self.manifold = self.create_manifold()
        self.scalar_parametrization = scalar_parametrization
        self.fc_mean = torch.nn.Linear(in_dim, self.mean_dim)
        self.fc_logvar = torch.nn.Linear(in_dim, 1 if scalar_parametrization else self.true_dim)

    @property
    def device(self) -> torch.device:
        return self.fc_mean.weight.device

    def radius_parameter(self):
        """(name, parameter) of the component's curvature parameter: a raw radius, or the raw curvature of 'u'."""
        for name in ("_nradius", "_pradius", "_curvature"):
            if hasattr(self, name):
                return name, getattr(self, name)
        return None, None

    def effective_kind(self) -> int:
        """Manifold the component currently lives on (differs from `kind` only for 'u')."""
        return self.kind

    def encode(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        """component.py:63-75."""
        z_mean = F.linear(x, self.fc_mean.weight, self.fc_mean.bias)
        z_mean_h = self.manifold.exp_map_mu0(z_mean.float().contiguous())
        std = F.softplus(F.linear(x, self.fc_logvar.weight, self.fc_logvar.bias)) + 1e-5
        return z_mean_h, std

    def reparametrize(self, z_mean: Tensor, std: Tensor):
        """WrappedNormalProcedure / EuclideanNormalProcedure.reparametrize (sampling_procedures.py:93-99,147-151)."""
        if self.effective_kind() == L.EUCLIDEAN:
            return EuclideanNormal(z_mean, std), EuclideanNormal(torch.zeros_like(z_mean), torch.ones_like(std))
        man = self.manifold.manifold if self.kind == L.UNIVERSAL else self.manifold
        q_z = WrappedNormal(z_mean, std, man)
        mu_0 = man.mu_0(z_mean.shape, device=z_mean.device, dtype=z_mean.dtype)
        p_z = WrappedNormal(mu_0, torch.ones_like(q_z.scale), man)
        return q_z, p_z

    def forward(self, x: Tensor):
        z_params = self.encode(x)
        q_z, p_z = self.reparametrize(*z_params)
        return q_z, p_z, z_params

    def kl_loss(self, q_z, p_z, z: Tensor, data) -> Tensor:
        """sampling_procedures.py:101-116 (log q - log p, one-sample MC) and :153-155 (analytic for Euclidean)."""
        if self.effective_kind() == L.EUCLIDEAN:
            return torch.distributions.kl.kl_divergence(q_z, p_z).sum(dim=-1)
        return q_z.log_prob_from_parts(z, data) - p_z.log_prob(z)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(R^{self.dim})"

    def _shortcut(self) -> str:
        return f"{self.letter}{self.true_dim}"

    def summary_name(self, comp_idx: int) -> str:
        return f"comp_{comp_idx:03d}_{self._shortcut()}"

    def summaries(self, comp_idx: int, q_z, prefix: str = "train") -> dict:
        """component.py:95-100: histograms Trainer logs under --train_statistics (train.py:204-206)."""
        name = prefix + "/" + self.summary_name(comp_idx)
        return {name + "/mean/norm": torch.norm(q_z.mean, p=2, dim=-1),
                name + "/stddev/norm": torch.norm(q_z.stddev, p=2, dim=-1)}

    def create_manifold(self) -> Manifold:
        raise NotImplementedError

    @property
    def true_dim(self) -> int:
        raise NotImplementedError

    @property
    def mean_dim(self) -> int:
        return self.true_dim


class HyperbolicComponent(Component):
    """component.py:114-130: ambient dim = n + 1, learnable `_nradius`."""
    letter, kind = "h", L.HYPERBOLOID

    def __init__(self, dim: int, fixed_curvature: bool, radius: float = 1.0) -> None:
        super().__init__(dim + 1, fixed_curvature)
        self._nradius = torch.nn.Parameter(torch.tensor(radius), requires_grad=not fixed_curvature)

    def create_manifold(self) -> Manifold:
        return Hyperboloid(lambda: self._nradius)

    @property
    def true_dim(self) -> int:
        return self.dim - 1


class PoincareComponent(Component):
    """component.py:133-149."""
    letter, kind = "p", L.POINCARE

    def __init__(self, dim: int, fixed_curvature: bool, radius: float = 1.0) -> None:
        super().__init__(dim, fixed_curvature)
        self._nradius = torch.nn.Parameter(torch.tensor(radius), requires_grad=not fixed_curvature)

    def create_manifold(self) -> Manifold:
        return PoincareBall(lambda: self._nradius)

    @property
    def true_dim(self) -> int:
        return self.dim


class SphericalComponent(Component):
    """component.py:152-167: ambient dim = n + 1, learnable `_pradius`."""
    letter, kind = "s", L.SPHERE

    def __init__(self, dim: int, fixed_curvature: bool, radius: float = 1.0) -> None:
        super().__init__(dim + 1, fixed_curvature)
        self._pradius = torch.nn.Parameter(torch.tensor(radius), requires_grad=not fixed_curvature)

    def create_manifold(self) -> Manifold:
        return Sphere(lambda: self._pradius)

    @property
    def true_dim(self) -> int:
        return self.dim - 1


class StereographicallyProjectedSphereComponent(Component):
    """component.py:170-189: the sphere in stereographic coordinates (dim = n), learnable `_pradius`."""
    letter, kind = "d", L.PROJ_SPHERE

    def __init__(self, dim: int, fixed_curvature: bool, radius: float = 1.0) -> None:
        super().__init__(dim, fixed_curvature)
        self._pradius = torch.nn.Parameter(torch.tensor(radius), requires_grad=not fixed_curvature)

    def create_manifold(self) -> Manifold:
        return StereographicallyProjectedSphere(lambda: self._pradius)

    @property
    def true_dim(self) -> int:
        return self.dim


class UniversalComponent(Component):
    """component.py:225-242: learnable `_curvature` (kappa); the manifold is the Poincare ball, the projected sphere or
    the Euclidean plane according to its sign (universal.py), the sampling procedure follows
    (UniversalSamplingProcedure, sampling_procedures.py:184-206)."""
    letter, kind = "u", L.UNIVERSAL

    def __init__(self, dim: int, fixed_curvature: bool, curvature: float = 0.0, eps: float = 1e-6) -> None:
        super().__init__(dim, fixed_curvature)
        self._curvature = torch.nn.Parameter(torch.tensor(curvature), requires_grad=not fixed_curvature)
        self._eps = eps

    def create_manifold(self) -> Manifold:
        return Universal(lambda: self._curvature, eps=self._eps)

    def effective_kind(self) -> int:
        return {-1: L.POINCARE, 0: L.EUCLIDEAN, 1: L.PROJ_SPHERE}[self.manifold._choice]

    @property
    def true_dim(self) -> int:
        return self.dim


class EuclideanComponent(Component):
    """component.py:192-203: always fixed curvature."""
    letter, kind = "e", L.EUCLIDEAN

    def __init__(self, dim: int, fixed_curvature: bool = True) -> None:
        super().__init__(dim, fixed_curvature=True)

    def create_manifold(self) -> Manifold:
        return Euclidean()

    @property
    def true_dim(self) -> int:
        return self.dim


space_creator_map = {"h": HyperbolicComponent, "s": SphericalComponent, "d": StereographicallyProjectedSphereComponent,
                     "p": PoincareComponent, "e": EuclideanComponent, "u": UniversalComponent}


def parse_component_str(space_str: str) -> Tuple[int, str, int]:
    """mt/mvae/utils.py:78-100: '[mult]<letter><dim>' (a '-<sampling>' suffix is accepted when it names the default
    wrapped-normal procedure)."""
    space_str = space_str.split("-")[0]
    m = re.fullmatch(r"(\d*)([a-z])(\d+)", space_str)
    if not m:
        raise ValueError(f"Cannot parse component '{space_str}'.")
    return int(m.group(1) or 1), m.group(2), int(m.group(3))


def parse_components(arg: str, fixed_curvature: bool) -> List[Component]:
    """mt/mvae/utils.py:103-140.  Letters h, s, d, p, e, u ('c' is outside this hot path: SURVEY.md §8f)."""
    arg = arg.lower().strip()
    if not arg:
        return []
    components: List[Component] = []
    for space_str in (s.strip() for s in arg.split(",")):
        mult, letter, dim = parse_component_str(space_str)
        if mult < 1:
            raise ValueError(f"Space multiplier has to be at least 1, was: '{mult}'.")
        if dim < 1:
            raise ValueError(f"Dimension has to be at least 1, was: '{dim}'.")
        if letter not in space_creator_map:
            raise NotImplementedError(f"Unknown / unsupported latent space type '{letter}'.")
        for _ in range(mult):
            components.append(space_creator_map[letter](dim, fixed_curvature))
    return components


def canonical_name(components: List[Component]) -> str:
    """mt/mvae/utils.py canonical_name: sorted, multiplicities merged ('3h3,2s2,1e1,e2' -> 'e1,e2,3h3,2s2')."""
    counts = {}
    for c in components:
        key = (c.letter, c.true_dim)
        counts[key] = counts.get(key, 0) + 1
    parts = []
    for (letter, dim), k in sorted(counts.items()):
        parts.append(f"{k if k > 1 else ''}{letter}{dim}")
    return ",".join(parts)
