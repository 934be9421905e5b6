"""FusedFeedForwardVAE — drop-in for the reference's `FeedForwardVAE` / `ModelVAE`
(mt/mvae/models/ffnn_vae.py:27-60, mt/mvae/models/vae.py:41-166) whose whole train_step runs in hand-written
sm_100a kernels (libmvae_b200.so):

  encode   fc_e0 + relu                      tcgen05 GEMM, BIAS_RELU epilogue -> split-bf16 planes of h
  heads    all fc_mean / fc_logvar at once   tcgen05 GEMM (one concatenated weight [sum(n)+sum(l_n), H])
  latent   Component.encode + reparametrize + rsample_with_parts + kl_loss for every component
                                             ONE fused product-manifold kernel (mvae_pm_forward)
  decode   fc_d0 + relu, fc_logits           tcgen05 GEMMs; the BCE / Gaussian-NLL row sums and dloss/dlogits are
                                             fused into the logits GEMM epilogue (logits never reach HBM in training)
  ELBO     BatchStats (stats.py:144-202)     warp-shuffle reduction (mvae_elbo_reduce)
  backward dgrad / wgrad                     the same GEMM kernel reading the forward buffers MN-major; bias gradients
                                             ride along as an extra "ones" column; relu masks in the epilogue;
                                             the latent backward is mvae_pm_backward (recompute from inputs)
  update   Adam on one flat bucket + SGD on the radii (train.py:327-360)   mvae_adam_step / mvae_sgd_step

Parameters keep the reference's names and shapes (state_dict round-trips) but live as views
into one flat fp32 buffer; gradients are views into one flat bucket [net grads | radius grads | ELBO stats] so that
data-parallel training needs a single SUM all-reduce per step (mvae_b200/parallel.py).
"""
import ctypes
import os
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib as L
from . import ops
from .components import Component
from .distributions import EuclideanNormal, WrappedNormal

# ---------------------------------------------------------------------------------------------- component adapters
# The model accepts the mirror components of mvae_b200.components AND the reference's own Component objects
# (mt/mvae/components/component.py:30-242, as built by mt.mvae.utils.parse_components): the reference's Trainer
# singles components out with isinstance checks against ITS classes (radius warm-up, train.py:189-194), which only
# hold for its own objects.  Everything the kernels need is read through these three functions.
_KIND_OF_CLASS = {"HyperbolicComponent": L.HYPERBOLOID, "SphericalComponent": L.SPHERE,
                  "PoincareComponent": L.POINCARE, "StereographicallyProjectedSphereComponent": L.PROJ_SPHERE,
                  "EuclideanComponent": L.EUCLIDEAN, "UniversalComponent": L.UNIVERSAL}
_SUPPORTED_PROCEDURES = ("WrappedNormalProcedure", "EuclideanNormalProcedure", "UniversalSamplingProcedure")


def kind_of(c) -> int:
    """Manifold kind (mvae_manifold) of a mirror or reference component."""
    k = getattr(c, "kind", None)
    if isinstance(k, int) and k >= 0:
        return k
    for cls in type(c).__mro__:
        if cls.__name__ in _KIND_OF_CLASS:
            proc = getattr(c, "_sampling_procedure_type", None)
            if proc is not None and proc.__name__ not in _SUPPORTED_PROCEDURES:
                raise NotImplementedError(f"sampling procedure {proc.__name__} is outside the fused hot path "
                                          "(wrapped / Euclidean normal only, SURVEY.md §8)")
            return _KIND_OF_CLASS[cls.__name__]
    raise NotImplementedError(f"component type {type(c).__name__} is outside the fused hot path")


def radius_parameter_of(c):
    """(name, parameter) of the component's curvature parameter: a raw radius, or the raw curvature of 'u'."""
    for name in ("_nradius", "_pradius", "_curvature"):
        if hasattr(c, name):
            return name, getattr(c, name)
    return None, None


def effective_kind_of(c) -> int:
    """Manifold the component currently lives on (differs from kind_of only for 'u': universal.py:64-74)."""
    k = kind_of(c)
    if k != L.UNIVERSAL:
        return k
    kappa, eps = float(c._curvature.detach()), float(getattr(c, "_eps", 1e-6))
    return L.POINCARE if kappa < -eps else (L.PROJ_SPHERE if kappa > eps else L.EUCLIDEAN)


class Reparametrized:
    """mt/mvae/models/vae.py:29-35.  `data` (= (u, v) of rsample_with_parts) is recomputed on demand."""

    def __init__(self, q_z, p_z, z: Tensor, kl: Tensor, eps: Optional[Tensor]) -> None:
        self.q_z = q_z
        self.p_z = p_z
        self.z = z
        self.kl = kl
        self._eps = eps
        self._data = None

    @property
    def data(self):
        if not hasattr(self.q_z, "manifold"):  # EuclideanNormal (mirror or reference): no parts
            return None
        if self._data is None:
            v = self._eps * self.q_z.scale
            _, (u, _) = self.q_z.manifold.sample_projection_mu0(v.contiguous(), self.q_z.loc)
            self._data = (u, v)
        return self._data


class LazyTensor:
    """A tensor computed on first use.  train_step returns the reference's Outputs triple (vae.py:166); the training
    kernels never materialise the logits x_mb_ (the reconstruction loss is fused into the GEMM that would produce
    them), so the third element is this proxy: touching it decodes the step's latent sample once."""

    def __init__(self, fn) -> None:
        object.__setattr__(self, "_fn", fn)
        object.__setattr__(self, "_t", None)

    def get(self) -> Tensor:
        if self._t is None:
            object.__setattr__(self, "_t", self._fn())
        return self._t

    def __getattr__(self, name):
        return getattr(self.get(), name)

    def __getitem__(self, idx):
        return self.get()[idx]

    def __len__(self) -> int:
        return len(self.get())

    def __repr__(self) -> str:
        return "LazyTensor(" + ("pending" if self._t is None else repr(self._t)) + ")"

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        unwrap = lambda a: a.get() if isinstance(a, LazyTensor) else a  # noqa: E731
        return func(*[unwrap(a) for a in args], **{k: unwrap(v) for k, v in (kwargs or {}).items()})


class LazyReparametrized:
    """List[Reparametrized] of the last train step, built on first use (the training kernels keep q_z's loc / scale in
    registers; Trainer._train_epoch reads them only under --train_statistics, train.py:200-206)."""

    def __init__(self, fn) -> None:
        self._fn, self._items = fn, None

    def _get(self) -> List[Reparametrized]:
        if self._items is None:
            self._items = self._fn()
        return self._items

    def __iter__(self):
        return iter(self._get())

    def __len__(self) -> int:
        return len(self._get())

    def __getitem__(self, idx):
        return self._get()[idx]


class BatchStatsFloat:
    """mt/mvae/stats.py:115-142: the per-step scalars, read back with ONE device->host copy."""

    def __init__(self, bce: float, kl: float, elbo: float, component_kl: List[float], beta: float,
                 log_likelihood=None, mutual_info=None, cov_norm=None) -> None:
        self.bce, self.kl, self.elbo = bce, kl, elbo
        self.log_likelihood, self.mutual_info, self.cov_norm = log_likelihood, mutual_info, cov_norm
        self.component_kl = component_kl
        self.beta = beta

    def to_print(self):
        return {"bce": self.bce, "kl": self.kl, "elbo": self.elbo,
                "ll": 0.0 if self.log_likelihood is None else self.log_likelihood,
                "mi": 0.0 if self.mutual_info is None else self.mutual_info,
                "cov_norm": 0.0 if self.cov_norm is None else self.cov_norm, "beta": self.beta}

    def summaries(self, stats, prefix: str = "train/batch") -> None:
        stats.add_scalar(prefix + "/bce", self.bce)
        stats.add_scalar(prefix + "/kl", self.kl)
        stats.add_scalar(prefix + "/elbo", self.elbo)
        if self.log_likelihood is not None:
            stats.add_scalar(prefix + "/log_likelihood", self.log_likelihood)
        if self.mutual_info is not None:
            stats.add_scalar(prefix + "/mutual_info", self.mutual_info)
        if self.cov_norm is not None:
            stats.add_scalar(prefix + "/cov_norm", self.cov_norm)


class BatchStats:
    """mt/mvae/stats.py:144-212 over the device vector [bce_sum, kl_sum, elbo, kl_c...] of mvae_elbo_reduce."""

    def __init__(self, stats_vec: Tensor, beta: float, bce_rows: Optional[Tensor] = None,
                 kl_rows: Optional[Tensor] = None, likelihood=None) -> None:
        self._vec = stats_vec
        self._beta = beta
        self._bce = bce_rows
        self._component_kl = None if kl_rows is None else [kl_rows[:, i] for i in range(kl_rows.shape[1])]
        # likelihood = (log_p_x [B], mi [B], cov_norm [1]) of log_likelihood(); reported summed over the batch
        # (stats.py:177-186).  The two sums are one more launch of the ELBO reduction kernel.
        self._log_likelihood, self._mutual_info, self._cov_norm = likelihood if likelihood is not None else (None,) * 3
        self._lsum = None
        if likelihood is not None:
            self._lsum = ops.elbo_reduce(self._log_likelihood, self._mutual_info.reshape(-1, 1), 0.0)

    bce = property(lambda self: self._vec[0])
    kl = property(lambda self: self._vec[1])
    elbo = property(lambda self: self._vec[2])
    beta = property(lambda self: self._beta)
    log_likelihood = property(lambda self: None if self._lsum is None else self._lsum[0])
    mutual_info = property(lambda self: None if self._lsum is None else self._lsum[1])
    cov_norm = property(lambda self: None if self._cov_norm is None else self._cov_norm.reshape(()))

    @property
    def component_kl(self) -> List[Tensor]:
        return [self._vec[3 + i] for i in range(self._vec.numel() - 3)]

    def convert_to_float(self) -> BatchStatsFloat:
        h = self._vec.detach().cpu().tolist()  # one D2H copy (the reference does 3 + C .item() syncs)
        if self._lsum is None:
            return BatchStatsFloat(h[0], h[1], h[2], h[3:], self._beta)
        ll, mi = self._lsum[:2].cpu().tolist()
        return BatchStatsFloat(h[0], h[1], h[2], h[3:], self._beta, ll, mi, float(self._cov_norm.reshape(-1)[0].item()))


def _recon_kind_of(dataset) -> str:
    kind = getattr(dataset, "recon_kind", None)
    if kind in ("bce", "nll"):
        return kind
    name = type(dataset).__name__.lower()
    if "bdp" in name:
        return "nll"  # mt/data/synthetic.py:161-162
    return "bce"      # mt/data/image_reconstruction.py:81-82, :142-143 (MNIST / Omniglot / CIFAR)


class _Workspace:
    """Per-batch-size activation buffers (allocated once, reused every step; addresses stay fixed for CUDA graphs)."""

    def __init__(self, m: "FusedFeedForwardVAE", B: int) -> None:
        dev, D, H, P, Sn, Sd, C = m.device, m.in_dim, m.h_dim, m.desc.ld_ml, m.desc.ld_eps, m.desc.ld_z, m.desc.C
        f = dict(device=dev, dtype=torch.float32)
        self.B = B
        # two input slots: train_epoch() copies batch i+1 host->device into one while the kernels of step i read the
        # other (the CUDA graphs are keyed by slot, the buffers never move)
        self.xbuf = [torch.zeros(B, D, **f), torch.zeros(B, D, **f)]
        self.slot = 0
        # uint8 image batches (binarised on the device, mvae_binarize): raw pixels per slot, allocated on first use;
        # u8 = the current batch arrived as uint8; bin_ctr = Philox step counter of the dynamic binarisation
        self.x8buf = None
        self._u8 = [False, False]  # per input slot (train_epoch stages batch i+1 while step i is being enqueued)
        self.bin_ctr = m._bin_ctr  # one counter per model: no two steps (of any batch size) share uniforms
        self.eps = torch.zeros(B, Sn, **f)
        self.xp = ops.PlaneBuf(B, D, m.input_planes, dev, ones_col=True)
        self.hp = ops.PlaneBuf(B, H, 3, dev, ones_col=True)   # 3 planes: feeds the heads at fp32 accuracy
        self.ml = torch.zeros(B, P, **f)
        self.z = torch.zeros(B, Sd, **f)
        self.kl = torch.zeros(B, C, **f)
        self.mu = torch.zeros(B, Sd, **f)
        self.sigma = torch.zeros(B, Sn, **f)
        self.ddp = ops.PlaneBuf(B, H, 2, dev, ones_col=True)
        self.bce = torch.zeros(B, **f)
        self.logits = None  # allocated on first forward() that needs them
        self.gLp = ops.PlaneBuf(B, D, 2, dev)
        # gdd feeds the reverse sweep of the manifold chain (ill conditioned near the log-det singularities): when it is
        # consumed as planes (latent_gemm) it carries 3 of them
        self.gddp = ops.PlaneBuf(B, H, 3 if m.latent_gemm else 2, dev)
        self.gz = torch.zeros(B, Sd, **f)
        self.gml = torch.zeros(B, P, **f)
        self.ghp = ops.PlaneBuf(B, H, 2, dev)
        self.flag = torch.zeros(1, device=dev, dtype=torch.int32)
        # fused latent block: the fc_e0 / logits-dgrad epilogues write h and gdd as fp32 for it (no planes needed)
        self.h32 = torch.zeros(B, H, **f)
        self.gdd32 = torch.zeros(B, H, **f)
        # wide product manifolds: the latent dense layers run on the tensor cores (FusedFeedForwardVAE.latent_gemm)
        if m.latent_gemm:
            self.zp = ops.PlaneBuf(B, Sd, 3, dev, ones_col=True)  # 3 planes: z feeds fc_d0 + relu at fp32 accuracy
            self.gmlp = ops.PlaneBuf(B, P, 3, dev)  # 3 planes: the head weight gradients are read at fp32 accuracy

    adopted = None  # a device-resident input batch used in place (FusedFeedForwardVAE.adopt_device_inputs)

    @property
    def x(self) -> Tensor:
        return self.adopted if self.adopted is not None else self.xbuf[self.slot]

    @property
    def u8(self) -> bool:
        return self._u8[self.slot]

    @property
    def x8(self) -> Tensor:
        if self.x8buf is None:
            self.x8buf = [torch.zeros(self.B, self.xbuf[0].shape[1], device=self.xbuf[0].device, dtype=torch.uint8)
                          for _ in range(2)]
        return self.x8buf[self.slot]


class FusedFeedForwardVAE(nn.Module):

    def __init__(self, h_dim: int, components: List[Component], dataset, scalar_parametrization: bool,
                 device="cuda", input_planes: Optional[int] = None) -> None:
        """Same signature as FeedForwardVAE (ffnn_vae.py:29-30) plus the device.  `input_planes`: bf16 planes carrying
        the input batch — 1 when the dataset declares `binary_inputs` (dynamically binarised images are exact in
        bf16), else 3 (fp32 accuracy for real-valued inputs)."""
        super().__init__()
        self.device = torch.device(device)
        self.h_dim = h_dim
        self.in_dim = dataset.in_dim
        self.recon_kind = _recon_kind_of(dataset)
        self.scalar_parametrization = scalar_parametrization
        self.input_planes = input_planes or (1 if getattr(dataset, "binary_inputs", False) else 3)
        self.components = nn.ModuleList(components)
        self.total_z_dim = sum(c.dim for c in components)
        self._kinds = [kind_of(c) for c in components]  # refuses components / procedures outside the hot path
        # construction order of the reference (vae.py:55-57 then ffnn_vae.py:35-40) => identical default init per seed
        for c in components:
            c.init_layers(h_dim, scalar_parametrization=scalar_parametrization)
        self._build_layers()
        self.desc = L.make_desc(self._kinds, [c.true_dim for c in components], scalar_parametrization)
        assert self.desc.ld_z == self.total_z_dim
        if self.desc.ld_ml > 64 or self.desc.ld_z > 64:
            raise NotImplementedError("product manifolds with more than 64 head outputs / latent coordinates")
        self._ws = {}
        self._graphs = {}
        self._gemm_tiles = {}
        self._eps_override: Optional[Tensor] = None
        self._flat = None
        self.check_finite = False
        # Philox key of the step's noise: taken from torch's global generator AFTER every layer has been initialised
        # (torch.manual_seed governs the noise like it does for the reference's Normal.rsample, without disturbing the
        # default initialisation, which must match the reference's per seed)
        self.noise_seed = int(torch.randint(0, 2**62, (1,)).item())
        # (on a CPU device this is host-side bookkeeping only — every kernel entry point refuses non-CUDA tensors)
        self._flatten()

    # ------------------------------------------------------------------------------------------ parameter storage
    def _build_layers(self) -> None:
        """ffnn_vae.py:35-40 (after the components: same construction order => same default initialisation)."""
        self.fc_e0 = nn.Linear(self.in_dim, self.h_dim)
        self.fc_d0 = nn.Linear(self.total_z_dim, self.h_dim)
        self.fc_logits = nn.Linear(self.h_dim, self.in_dim)

    def _head_params(self):
        heads_w, heads_b = [], []
        for i, c in enumerate(self.components):
            heads_w += [(f"components.{i}.fc_mean.weight", c.fc_mean.weight),
                        (f"components.{i}.fc_logvar.weight", c.fc_logvar.weight)]
            heads_b += [(f"components.{i}.fc_mean.bias", c.fc_mean.bias),
                        (f"components.{i}.fc_logvar.bias", c.fc_logvar.bias)]
        return heads_w + heads_b

    def _net_params(self) -> List[Tuple[str, nn.Parameter]]:
        rest = [("fc_e0.weight", self.fc_e0.weight), ("fc_e0.bias", self.fc_e0.bias),
                ("fc_d0.weight", self.fc_d0.weight), ("fc_d0.bias", self.fc_d0.bias),
                ("fc_logits.weight", self.fc_logits.weight), ("fc_logits.bias", self.fc_logits.bias)]
        return self._head_params() + rest

    def plane_targets(self):
        """[(offset in the flat buffer, rows, PlaneBuf, fp32 weight view)] of every GEMM weight whose split-bf16 planes
        the optimizer kernels refresh after an update."""
        t = [(self._slices["fc_e0.weight"][0], self.h_dim, self.We0p, self.fc_e0.weight.data),
             (self._slices["fc_logits.weight"][0], self.in_dim, self.Wlp, self.fc_logits.weight.data)]
        if self.latent_gemm:
            t += [(self._slices["components.0.fc_mean.weight"][0], self.desc.ld_ml, self.Whp, self.Wh),
                  (self._slices["fc_d0.weight"][0], self.h_dim, self.Wd0p, self.fc_d0.weight.data)]
        return t

    def dp_early_begin(self) -> int:
        """Offset from which the flat gradient bucket is complete EARLY in the backward pass (data parallel: that tail
        is exchanged on a side stream under the rest of the backward pass): here fc_logits, the last parameters."""
        return self._slices["fc_logits.weight"][0]

    def _master_perm(self, name: str):
        """Permutation p such that the flat buffer stores param.permute(p) contiguously (None: the reference's order)."""
        return None

    def flat_sizes(self) -> Tuple[int, int]:
        """(floats of the flat parameter buffer, floats of the gradient / statistics bucket)."""
        return self._n_net, self._n_net + 2 * self.desc.C + 3

    def _flatten(self, storage=None) -> None:
        """Re-home every parameter as a view into one flat fp32 buffer (and its gradient into one flat bucket)."""
        dev = self.device
        net = [t[:2] for t in self._net_params()]
        C = self.desc.C
        # Offsets: the head weights and the head biases stay contiguous blocks ([P, H] and [P]); every other tensor
        # (and the head-bias block) starts on a 16-byte boundary so that the kernels can use 128-bit accesses.
        # The few padding floats are zero parameters with zero gradients.
        n_heads = 2 * len(self.components)
        offsets, off = [], 0
        for i, (name, p) in enumerate(net):
            if i == n_heads or i >= 2 * n_heads:
                off = (off + 3) // 4 * 4
            offsets.append(off)
            off += p.numel()
        n_net = (off + 3) // 4 * 4
        if storage is None:
            flat = torch.zeros(n_net, device=dev, dtype=torch.float32)
            bucket = torch.zeros(n_net + C + 3 + C, device=dev, dtype=torch.float32)
        else:  # buffers provided by the caller (peer-mapped memory of the data-parallel step, parallel.attach_p2p)
            flat, bucket = storage
            assert flat.numel() == n_net and bucket.numel() == n_net + C + 3 + C
            assert flat.data_ptr() % 16 == 0 and bucket.data_ptr() % 16 == 0
            flat.zero_()
            bucket.zero_()
        rflat = torch.ones(C, device=dev, dtype=torch.float32)
        self._slices = {}
        self._grad_views = {}
        for (name, p), off in zip(net, offsets):
            n = p.numel()
            perm = self._master_perm(name)
            if perm is None:
                flat[off:off + n].copy_(p.data.reshape(-1).to(dev, torch.float32))
                p.data = flat[off:off + n].view(p.shape)
                p.grad = bucket[off:off + n].view(p.shape)
            else:
                # the kernels want this tensor in another memory order (e.g. conv filters with the channel innermost):
                # the flat buffer holds that order, the parameter keeps the reference's SHAPE as a permuted view of it
                # (state_dict / load_state_dict / torch optimizers see the reference's tensor)
                inv = [perm.index(i) for i in range(len(perm))]
                src = p.data.detach().to(dev, torch.float32).permute(perm).contiguous()
                flat[off:off + n].copy_(src.reshape(-1))
                p.data = flat[off:off + n].view(src.shape).permute(inv)
                p.grad = bucket[off:off + n].view(src.shape).permute(inv)
            self._slices[name] = (off, n)
            self._grad_views[name] = p.grad
        self._radius_mask = torch.zeros(C, device=dev, dtype=torch.float32)
        for i, c in enumerate(self.components):
            _, rp = radius_parameter_of(c)
            if rp is not None:
                rflat[i] = rp.data.to(dev, torch.float32)
                rp.data = rflat[i]
                if rp.requires_grad:
                    rp.grad = bucket[n_net + i]
                    self._radius_mask[i] = 1.0
        # universal components: the reference clips the 2-norm of their "curvature"-named gradients to 1 in every
        # train step (vae.py:161-163); _clip_mask selects them in the radius-gradient vector
        self._clip_mask = None
        if any(k == L.UNIVERSAL for k in self._kinds):
            self._clip_mask = torch.tensor([1.0 if k == L.UNIVERSAL else 0.0 for k in self._kinds], device=dev)
        self._radius_params = [radius_parameter_of(c)[1] for c in self.components]
        self._any_fixed_radius = any((rp is not None) and (not rp.requires_grad) for rp in self._radius_params)
        self._flat, self._rflat, self._bucket, self._n_net = flat, rflat, bucket, n_net
        self._gnet = bucket[:n_net]
        self._gradius = bucket[n_net:n_net + C]
        self._stats = bucket[n_net + C:]
        self._stats_report = self._stats  # where the step's (rank-summed) statistics are read from
        # what one device->host copy brings back per step: the statistics, followed (peer-memory data parallel) by the
        # sticky error word of mvae_dp_step
        self._stats_wire = self._stats_report
        self._planes_stale = True
        self._ws = {}
        self._graphs = {}
        self._gemm_tiles = {}
        self._bin_ctr = torch.zeros(1, device=dev, dtype=torch.int64)
        self._radius_ptrs = [rflat.data_ptr() + 4 * i for i in range(C)]
        self._bind_views(flat, bucket)

    def _bind_views(self, flat: Tensor, bucket: Tensor) -> None:
        """Layer-specific views into the flat buffers + the split-bf16 planes of the GEMM weights."""
        dev = self.device
        H, D, P, Sd = self.h_dim, self.in_dim, self.desc.ld_ml, self.desc.ld_z

        def view(buf, name, shape):
            o, n = self._slices[name]
            return buf[o:o + n].view(shape)

        o0, _ = self._slices["components.0.fc_mean.weight"]
        self.Wh, self.gWh = flat[o0:o0 + P * H].view(P, H), bucket[o0:o0 + P * H].view(P, H)
        b0, _ = self._slices["components.0.fc_mean.bias"]
        self.bh, self.gbh = flat[b0:b0 + P], bucket[b0:b0 + P]
        self.gWe0, self.gbe0 = view(bucket, "fc_e0.weight", (H, D)), view(bucket, "fc_e0.bias", (H,))
        self.gWd0, self.gbd0 = view(bucket, "fc_d0.weight", (H, Sd)), view(bucket, "fc_d0.bias", (H,))
        self.gWl, self.gbl = view(bucket, "fc_logits.weight", (D, H)), view(bucket, "fc_logits.bias", (D,))
        # split-bf16 planes of the four weight matrices (refreshed after every optimizer step)
        # Operands that produce the argument of a NON-SMOOTH function (the two relus, the manifold maps with their
        # clamps and singular log-dets) carry 3 planes (~2^-24, fp32 accuracy): with 2 planes (~2^-16) relu decisions
        # near zero flip ~30x more often than in a true fp32 run.  Smooth consumers (logits -> BCE, every dgrad / wgrad)
        # read 2 planes.
        # The heads and fc_d0 are "skinny" layers computed in exact fp32 on the CUDA cores (no planes at all).
        self.We0p = ops.PlaneBuf(H, D, 3, dev)
        self.Wlp = ops.PlaneBuf(D, H, 2, dev)
        # heads + manifold chain + fc_d0 as one kernel per direction (mvae_latent_forward / _backward)
        self.fused_latent = (H % 8 == 0) and P <= 64 and Sd <= 64 and os.environ.get("MVAE_FUSED_LATENT", "1") != "0"
        # Wide products (cfg3: 60 head outputs, 34 latent coordinates): the per-CTA weight-gradient reductions of the
        # fused block and the CUDA-core skinny kernels both scale with P * B; there the heads and fc_d0 (forward, dgrad,
        # wgrad) go through the tcgen05 GEMM instead — 3 operand planes where the result feeds a non-smooth function —
        # around the standalone product-manifold kernels (measured at cfg3: 600 us fused / 460 us skinny -> ~130 us).
        self.latent_gemm = (P > 16 or Sd > 16) and os.environ.get("MVAE_LATENT_GEMM", "1") != "0"
        if self.latent_gemm:
            self.fused_latent = False
            self.Whp = ops.PlaneBuf(P, H, 3, dev)
            self.Wd0p = ops.PlaneBuf(H, Sd, 3, dev)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        p = self._net_params()[0][1]
        if p.device.type == "cuda" and (self._flat is None or p.data.untyped_storage().data_ptr() !=
                                        self._flat.untyped_storage().data_ptr()):
            self.device = p.device
            self._flatten()
        return out

    def to(self, *args, **kwargs):
        out = super().to(*args, **kwargs)
        return out

    def load_state_dict(self, state_dict, strict: bool = True):
        res = super().load_state_dict(state_dict, strict=strict)  # copies in place into the flat views
        self._planes_stale = True
        return res

    def refresh_weight_planes(self) -> None:
        ops.split_planes(self.fc_e0.weight.data, self.We0p)
        ops.split_planes(self.fc_logits.weight.data, self.Wlp)
        if self.latent_gemm:
            ops.split_planes(self.Wh, self.Whp)
            ops.split_planes(self.fc_d0.weight.data, self.Wd0p)
        self._planes_stale = False

    def _sync_radii(self) -> None:
        """Trainer._train_epoch REBINDS `c._pradius.data = ones_like(...) * (11 - epoch)` during its first ten epochs
        (train.py:189-194): a fresh tensor the kernels — which read the flat radius vector — would never see.  A rebind
        is detected by address; the value moves into the flat vector and the parameter is re-homed there.  In-place
        writes (`fill_`, optimizers) land in the flat vector directly.  Called before every kernel sequence."""
        for i, rp in enumerate(self._radius_params):
            if rp is not None and rp.data_ptr() != self._radius_ptrs[i]:
                self._rflat[i:i + 1].copy_(rp.data.detach().reshape(1).to(self._rflat.device, torch.float32))
                rp.data = self._rflat[i]

    def mark_parameters_changed(self) -> None:
        """Call after modifying parameters outside of train_step (external optimizers are handled automatically)."""
        self._planes_stale = True

    def _workspace(self, B: int) -> _Workspace:
        ws = self._ws.get(B)
        if ws is None:
            ws = self._ws[B] = _Workspace(self, B)
        return ws

    # ------------------------------------------------------------------------------------------ kernel sequences
    def _comm_stream(self) -> "torch.cuda.Stream":
        st = getattr(self, "_comm", None)
        if st is None:
            st = self._comm = torch.cuda.Stream(device=self.device)
        return st

    def _side_stream(self) -> "torch.cuda.Stream":
        st = getattr(self, "_side", None)
        if st is None:
            st = self._side = torch.cuda.Stream(device=self.device)
        return st

    def _forward_kernels(self, ws: _Workspace, beta: float, train: bool, want_mu_sigma: bool, logits: Optional[Tensor],
                         draw_eps: bool = False):
        B, D, H, P, Sd = ws.B, self.in_dim, self.h_dim, self.desc.ld_ml, self.desc.ld_z
        if self._planes_stale:
            self.refresh_weight_planes()
        fused = self.fused_latent and not want_mu_sigma
        zero = [ws.bce, ws.ml if self.latent_gemm else None, self._bucket[:self._n_net + self.desc.C] if train else None]
        # The head of the step — eps ~ N(0, I) (Philox, offset = the model's device step counter), zero fill of the
        # reconstruction row sums and of the gradient bucket (optimizer.zero_grad(), vae.py:151).  With the fused
        # latent block that kernel takes it along (mvae_latent_forward_ex): nothing ahead of it accumulates into those
        # buffers, and as a launch of its own on a side branch it makes the latent kernel the node that JOINS the branch
        # — a full launch latency behind fc_e0 instead of a programmatic edge (scripts/step_timeline.py: the latent
        # kernel starts 3.3 us earlier and spends 2.5 us of that on the draw).  MVAE_FOLD_PROLOGUE=0 keeps the branch.
        fold = fused and self.fold_prologue
        main, side = torch.cuda.current_stream(self.device), self._side_stream()
        if not fold:
            # Fork: a side stream (a parallel branch of the step's CUDA graph) next to split(x) + fc_e0: ONE launch,
            # + zero fill of ml (the heads GEMM accumulates its K slices into it)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ops.step_prologue(ws.eps if draw_eps else None, self.noise_seed, self._bin_ctr, zero)
        ws.drew_eps = draw_eps
        self._input_planes(ws, train)
        ws.has_mu_sigma = want_mu_sigma
        if fused:
            self._gemm("e0_fwd32", ws.xp, self.We0p, B, H, D, epilogue=L.EPI_BIAS_RELU, bias=self.fc_e0.bias.data,
                       out_f32=ws.h32)
        else:
            self._gemm("e0_fwd", ws.xp, self.We0p, B, H, D, epilogue=L.EPI_BIAS_RELU, bias=self.fc_e0.bias.data,
                       out_planes=ws.hp)
        ws.fused = fused
        if not fold:
            main.wait_stream(side)  # join: eps drawn, bce / gradient bucket zeroed
        if fused:
            # heads + manifold chain + fc_d0/relu in ONE kernel: ml, z, kl kept for the backward pass / statistics
            ops.latent_forward(self.desc, ws.h32, self.Wh, self.bh, ws.eps, self._rflat, self.fc_d0.weight.data,
                               self.fc_d0.bias.data, ws.ml, ws.z, ws.kl, ws.ddp,
                               flag=ws.flag if self.check_finite else None,
                               draw=(self.noise_seed, self._bin_ctr) if (fold and draw_eps) else None,
                               zero=zero if fold else ())
        elif self.latent_gemm:
            # wide product: heads and fc_d0 on the tensor cores (3 planes each side: fp32 accuracy ahead of the
            # manifold maps / the relu), the standalone product-manifold kernel between them
            self._heads_gemm(ws)
            out = {"z": ws.z, "kl": ws.kl, "mu": ws.mu, "sigma": ws.sigma}
            ops.pm_forward(self.desc, ws.ml, ws.eps, self._rflat, want_mu_sigma=want_mu_sigma,
                           flag=ws.flag if self.check_finite else None, out=out)
            ops.split_planes(ws.z, ws.zp)
            self._gemm("d0_fwd", ws.zp, self.Wd0p, B, H, Sd, epilogue=L.EPI_BIAS_RELU, bias=self.fc_d0.bias.data,
                       out_planes=ws.ddp)
        else:
            # heads: N = P is tiny -> CUDA-core row dots in exact fp32 (h read from its 3 planes)
            ops.skinny_rowdot((ws.hp, 3), self.Wh, H, 1, K=H, N=P, bias=self.bh, out=ws.ml)
            out = {"z": ws.z, "kl": ws.kl, "mu": ws.mu, "sigma": ws.sigma}
            ops.pm_forward(self.desc, ws.ml, ws.eps, self._rflat, want_mu_sigma=want_mu_sigma,
                           flag=ws.flag if self.check_finite else None, out=out)
            # fc_d0: K = total_z_dim is tiny -> CUDA-core expansion in exact fp32, relu, planes of dd for the logits GEMM
            ops.skinny_expand(ws.z, self.fc_d0.weight.data, Sd, 1, K=Sd, N=H, bias=self.fc_d0.bias.data,
                              act=ops.ACT_RELU, out_planes=ws.ddp)
        epi = L.EPI_BCE_ROWSUM if self.recon_kind == "bce" else L.EPI_NLL_ROWSUM
        self._gemm("logits_fwd", ws.ddp, self.Wlp, B, D, H, epilogue=epi, bias=self.fc_logits.bias.data, aux=ws.x, rowsum=ws.bce,
                 out_planes=ws.gLp if train else None, out_f32=logits)
        if not train:  # training: the reduction runs beside the backward GEMMs (_backward_kernels)
            ops.elbo_reduce(ws.bce, ws.kl, beta, out=self._stats)

    def _backward_kernels(self, ws: _Workspace, beta: float, early: bool = True, advance: Optional[bool] = None):
        """`early`: let a data-parallel optimizer exchange + update fc_logits on a third stream as soon as its gradient
        is complete (FusedCurvatureOptimizer.step_early), and advance the step counter behind the Philox draws —
        False for the warm-up pass ahead of a graph capture, which no optimizer step follows and which must leave the
        model's state (counters, exchanged parameters) untouched."""
        advance = early if advance is None else advance
        B, D, H, P, Sd, C = ws.B, self.in_dim, self.h_dim, self.desc.ld_ml, self.desc.ld_z, self.desc.C
        MN = L.MN_MAJOR
        # Fork: the ELBO reduction and the weight gradient of fc_logits depend only on the forward pass; they run as a
        # parallel branch beside  logits dgrad -> latent backward -> fc_e0 wgrad  (each of these GEMMs fills about half
        # of the SMs on its own).  The gradient bucket was zeroed by the forward pass's side branch.
        main, side = torch.cuda.current_stream(self.device), self._side_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if advance and (ws.u8 or ws.drew_eps):
                ops.counter_add(self._bin_ctr)  # the next step's noise / binarisation draws are fresh
            ops.elbo_reduce(ws.bce, ws.kl, beta, out=self._stats)
            # fc_logits: gW = gL^T dd (+ bias from the ones column of dd)
            self._gemm("logits_wgrad", ws.gLp, ws.ddp, D, H + 1, B, a_major=MN, b_major=MN, split_k=0, out_f32=self.gWl,
                       out_col=self.gbl, col_split=H)
        # gdd = (gL W) * 1[dd > 0]
        fused_latent = ws.fused  # what the forward pass of this step ran (train_statistics keeps mu / sigma: unfused)
        if fused_latent:
            self._gemm("logits_dgrad32", ws.gLp, self.Wlp, B, H, D, b_major=MN, epilogue=L.EPI_RELU_MASK, mask=ws.ddp,
                       out_f32=ws.gdd32)
        else:
            self._gemm("logits_dgrad", ws.gLp, self.Wlp, B, H, D, b_major=MN, epilogue=L.EPI_RELU_MASK, mask=ws.ddp,
                       out_planes=ws.gddp)
        early = early and self._early_step is not None
        if early:
            # fc_logits is final once its weight gradient (side branch) is in the bucket and the input-gradient GEMM
            # (main branch: the last reader of its weight planes) is done
            comm = self._comm_stream()
            comm.wait_stream(side)
            comm.wait_stream(main)
            with torch.cuda.stream(comm):
                self._early_step()
        if fused_latent:
            # fc_d0 dgrad + wgrad, manifold reverse sweep (d(-ELBO)/d kl = beta), heads dgrad + wgrad: ONE kernel
            ops.latent_backward(self.desc, ws.gdd32, ws.h32, self.Wh, self.fc_d0.weight.data, ws.ml, ws.eps, self._rflat,
                                ws.z, beta, ws.ghp, self.gWd0, self.gbd0, self.gWh, self.gbh, self._gradius)
            if self._any_fixed_radius:
                self._gradius.mul_(self._radius_mask)  # requires_grad=False radii (fixed curvature) get no gradient
        elif self.latent_gemm:
            # fc_d0: gW = gdd^T z (+ bias from the ones column of z) on the side branch, gz = gdd W on the main one
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self._gemm("d0_wgrad", ws.gddp, ws.zp, H, Sd + 1, B, a_major=MN, b_major=MN, split_k=0,
                           out_f32=self.gWd0, out_col=self.gbd0, col_split=Sd, a_planes=2, b_planes=2)
            self._gemm("d0_dgrad", ws.gddp, self.Wd0p, B, Sd, H, b_major=MN, out_f32=ws.gz)
            ops.pm_backward(self.desc, ws.ml, ws.eps, self._rflat, ws.gz, None, beta, gml=ws.gml, gradius=self._gradius)
            if self._any_fixed_radius:
                self._gradius.mul_(self._radius_mask)  # requires_grad=False radii (fixed curvature) get no gradient
            ops.split_planes(ws.gml, ws.gmlp)
            # heads: gWh = gml^T h (+ bias from h's ones column) on the side branch;  gh = (gml Wh) * 1[h > 0]
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self._gemm("heads_wgrad", ws.gmlp, ws.hp, P, H + 1, B, a_major=MN, b_major=MN, split_k=0,
                           out_f32=self.gWh, out_col=self.gbh, col_split=H)
            self._gemm("heads_dgrad", ws.gmlp, self.Whp, B, H, P, b_major=MN, epilogue=L.EPI_RELU_MASK, mask=ws.hp,
                       out_planes=ws.ghp, a_planes=2, b_planes=2)
        else:
            # fc_d0 (skinny): gW[h, j] = sum_b gdd[b, h] z[b, j], gb[h] = sum_b gdd[b, h];  gz = gdd W
            ops.skinny_wgrad(ws.z, Sd, (ws.gddp, 2), H, self.gWd0, 1, Sd, small_ones=True, out_row=self.gbd0)
            ops.skinny_rowdot((ws.gddp, 2), self.fc_d0.weight.data, 1, Sd, K=H, N=Sd, bias=None, out=ws.gz)
            # latent: d(-ELBO)/d kl = beta
            ops.pm_backward(self.desc, ws.ml, ws.eps, self._rflat, ws.gz, None, beta, gml=ws.gml, gradius=self._gradius)
            if self._any_fixed_radius:
                self._gradius.mul_(self._radius_mask)  # requires_grad=False radii (fixed curvature) get no gradient
            # heads (skinny): gWh[p, k] = sum_b gml[b, p] h[b, k] (+ bias from h's ones column);  gh = (gml Wh) * 1[h > 0]
            ops.skinny_wgrad(ws.gml, P, (ws.hp, 2), H + 1, self.gWh, H, 1, out_col=self.gbh, col_split=H)
            ops.skinny_expand(ws.gml, self.Wh, 1, H, K=P, N=H, act=ops.ACT_MASK, mask=ws.hp, out_planes=ws.ghp)
        # fc_e0 (no dgrad into x)
        self._gemm("e0_wgrad", ws.ghp, ws.xp, H, D + 1, B, a_major=MN, b_major=MN, split_k=0, out_f32=self.gWe0,
                 out_col=self.gbe0, col_split=D, b_planes=2)
        if early and getattr(self._early_step.__self__, "early_is_independent", False) \
                and self._early_step.__self__._early_done:
            # The launch that ends the step touches nothing the side branches produce (ELBO statistics, fc_logits'
            # gradient and update): it follows fc_e0's weight gradient directly, and the branches are joined AFTER it
            # (_join_pending).  Joined here, the final launch was a three-way join node of the graph — a full launch
            # latency (~3.5 us, scripts/step_timeline.py) behind the last GEMM instead of a programmatic edge.
            self._pending_join = (side, comm)
            return
        main.wait_stream(side)  # join
        if early:
            main.wait_stream(comm)

    def _join_pending(self) -> None:
        """Join the side branches whose join _backward_kernels left to the end of the step."""
        pj = getattr(self, "_pending_join", None)
        if pj is not None:
            main = torch.cuda.current_stream(self.device)
            for st in pj:
                main.wait_stream(st)
            self._pending_join = None

    # ------------------------------------------------------------------------------------------ GEMM tile policy
    # The five big GEMMs of a step are launch- and L2-bound at these shapes and their best tile (BLOCK_N, CTAs per SM,
    # split-K) depends on M, N, K and the operand layouts: each call site times a handful of candidates ONCE per batch
    # size (CUDA events, first eager step) and keeps the fastest.  MVAE_GEMM_AUTOTUNE=0 keeps the automatic policy.
    _GEMM_TILES = [None, (48, 2), (64, 2), (96, 1), (112, 1), (112, 2), (128, 1), (208, 1)]
    # (the split-K GEMMs are the weight gradients, K = batch: their main loops run at the L2 throughput cap —
    # scripts/gemm_phases.py — and a wider tile re-reads the activations less often: 128 x 208 moves ~30 % fewer bytes
    # than 128 x 112 for the same product)
    _GEMM_TILES_SPLITK = [None, (64, 2, 0), (64, 2, 4), (64, 2, 6), (80, 1, 0), (96, 1, 0), (112, 1, 0), (128, 1, 0),
                          (144, 1, 0), (176, 1, 0), (208, 1, 0), (256, 1, 0)]
    autotune_gemm = os.environ.get("MVAE_GEMM_AUTOTUNE", "1") != "0"

    def _gemm(self, site: str, a, b, M: int, N: int, K: int, **kw) -> None:
        key = (site, M, N, K)
        if key not in self._gemm_tiles:
            if self.autotune_gemm and not torch.cuda.is_current_stream_capturing():
                self._gemm_tiles[key] = self._tune_gemm(a, b, M, N, K, kw)
            else:
                return ops.gemm(a, b, M, N, K, **kw)
        ops.gemm(a, b, M, N, K, tile=self._gemm_tiles[key], **kw)

    def _heads_gemm(self, ws: _Workspace) -> None:
        """ml = h Wh^T + bh on the tensor cores (wide products), accumulated into a ZEROED ws.ml.  The head
        pre-activations feed the manifold maps, whose gradients are ill conditioned near the log-det singularities, and
        the tensor core truncates (does not round) every add into its fp32 accumulator: one CTA per 64-wide K slice,
        combined by fp32 atomics (round to nearest), keeps the number of truncating adds per output at 24 instead of
        6 * K / 16 (measured: gradient error of the sphere components 1.9e-4 -> fp32-level)."""
        ops.gemm(ws.hp, self.Whp, ws.B, self.desc.ld_ml, self.h_dim, bias=self.bh, out_f32=ws.ml,
                 split_k=(self.h_dim + 63) // 64)

    def _tune_gemm(self, a, b, M: int, N: int, K: int, kw) -> Optional[tuple]:
        acc = [t for t in (kw.get("rowsum"), kw.get("out_f32") if kw.get("split_k", 1) != 1 else None,
                           kw.get("out_col") if kw.get("split_k", 1) != 1 else None) if t is not None]
        saved = [t.clone() for t in acc]  # accumulating outputs are restored afterwards
        cands = self._GEMM_TILES_SPLITK if kw.get("split_k", 1) != 1 else self._GEMM_TILES
        best, best_ms = None, float("inf")
        n0 = ops.launch_count()
        for tile in cands:
            try:
                for _ in range(2):
                    ops.gemm(a, b, M, N, K, tile=tile, **kw)
                ms = float("inf")
                for _ in range(3):   # best of three rounds: one round's noise used to flip between near-equal tiles
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    for _ in range(8):
                        ops.gemm(a, b, M, N, K, tile=tile, **kw)
                    e.record()
                    e.synchronize()
                    ms = min(ms, s.elapsed_time(e))
            except L.MvaeError:
                continue
            if ms < best_ms:
                best, best_ms = tile, ms
        for t, sv in zip(acc, saved):
            t.copy_(sv)
        ops.add_launches(n0 - ops.launch_count())  # tuning launches are not part of any step
        return best

    def _input_planes(self, ws: _Workspace, train: bool) -> None:
        """The staged batch -> fp32 targets ws.x + the bf16 operand planes of fc_e0."""
        if ws.u8:
            # uint8 pixels -> binarised fp32 targets + the bf16 operand plane of fc_e0 in one kernel
            # (image_reconstruction.py:37-53: dynamic in training, threshold 0.5 in evaluation)
            ops.binarize(ws.x8, x=ws.x, planes=ws.xp, seed=self.binarize_seed, offset_dev=ws.bin_ctr,
                         dynamic=train or self.binarize_eval_dynamic, invert=self.binarize_invert)
            if self.input_planes > 1:  # binarised values are exact in plane 0: the residual planes are zero
                ws.xp.t[1:, :, :self.in_dim].zero_()
        else:
            ops.split_planes(ws.x, ws.xp)

    def _stage(self, ws: _Workspace, x: Tensor, eps: Optional[Tensor]) -> bool:
        """Input batch (and supplied noise) into the workspace.  Returns True when the noise has to be drawn: the
        forward kernels then do it on their side branch (inside the step's CUDA graph)."""
        self._stage_x(ws, ws.slot, x)  # H2D if x lives on the host
        if eps is None:
            eps = self._eps_override
        if eps is None:
            return True
        ws.eps.copy_(eps, non_blocking=True)
        return False

    noise_seed = 0x5EED            # Philox key of the in-kernel N(0, I) draws (set per model from torch's generator)
    # True: a float32 batch that already lives on the model's device is read IN PLACE by the step's kernels instead of
    # being copied into the workspace (saves the copy and its launch).  The tensor must stay alive and unchanged until
    # the step has run; the CUDA graph of the step is keyed by its address (one graph per distinct batch tensor, at most
    # 32 kept), so this suits a set of resident batches that is cycled through, not a fresh tensor every step.
    adopt_device_inputs = False
    binarize_seed = 0              # Philox key of the on-device dynamic binarisation
    binarize_invert = False        # ImageDynamicBinarization(invert=...) (Omniglot)
    binarize_eval_dynamic = False  # forward() / log_likelihood() use the fixed 0.5 threshold like the reference's test loader

    def _stage_x(self, ws: _Workspace, slot: int, x: Tensor) -> None:
        """Copy a batch into input slot `slot`: float batches as they are, uint8 image batches (raw grayscale pixels,
        what the dataset stores) into the slot's uint8 buffer — they are binarised on the device by the forward
        kernels.  Marks which kind the slot's consumer has to read."""
        ws.adopted = None
        if x.dtype == torch.uint8:
            ws.x8  # allocate on first use
            ws.x8buf[slot].copy_(x.reshape(ws.B, self.in_dim), non_blocking=True)
            ws._u8[slot] = True
        elif (self.adopt_device_inputs and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and
              x.shape[1] == self.in_dim and x.is_contiguous() and x.device == self.device and x.data_ptr() % 16 == 0):
            ws.adopted = x   # read in place: no device-to-device copy (the step's graph is keyed by the address)
            ws._u8[slot] = False
        else:
            ws.xbuf[slot].copy_(x.reshape(ws.B, self.in_dim), non_blocking=True)
            ws._u8[slot] = False

    # ------------------------------------------------------------------------------------------ reference API
    def encode(self, x: Tensor) -> Tensor:
        """ffnn_vae.py:42-50 -> relu(fc_e0(x)) as fp32."""
        ws = self._workspace(x.shape[0])
        self._stage_x(ws, ws.slot, x)
        if self._planes_stale:
            self.refresh_weight_planes()
        self._input_planes(ws, train=False)
        h = torch.empty(ws.B, self.h_dim, device=self.device)
        ops.gemm(ws.xp, self.We0p, ws.B, self.h_dim, self.in_dim, epilogue=L.EPI_BIAS_RELU,
                 bias=self.fc_e0.bias.data, out_planes=ws.hp, out_f32=h)
        return h

    def decode(self, concat_z: Tensor) -> Tensor:
        """ffnn_vae.py:52-60 for [B, total_z_dim] or [n, B, total_z_dim]."""
        lead = concat_z.shape[:-1]
        z2 = concat_z.reshape(-1, self.total_z_dim).float().contiguous()
        Bz = z2.shape[0]
        if self._planes_stale:
            self.refresh_weight_planes()
        ddp = ops.PlaneBuf(Bz, self.h_dim, 2, self.device)
        ops.skinny_expand(z2, self.fc_d0.weight.data, self.total_z_dim, 1, K=self.total_z_dim, N=self.h_dim,
                          bias=self.fc_d0.bias.data, act=ops.ACT_RELU, out_planes=ddp)
        out = torch.empty(Bz, self.in_dim, device=self.device)
        ops.gemm(ddp, self.Wlp, Bz, self.in_dim, self.h_dim, bias=self.fc_logits.bias.data, out_f32=out)
        return out.reshape(*lead, self.in_dim)

    def _reparametrized(self, ws: _Workspace) -> List[Reparametrized]:
        res = []
        for i, c in enumerate(self.components):
            d = self.desc.comp[i]
            loc = ws.mu[:, d.z_off:d.z_off + d.d]
            scale = ws.sigma[:, d.eps_off:d.eps_off + d.n]
            z = ws.z[:, d.z_off:d.z_off + d.d]
            if hasattr(c, "sampling_procedure"):
                # a reference component: its own procedure builds its own distribution types
                # (sampling_procedures.py:93-99,147-151)
                q_z, p_z = c.sampling_procedure.reparametrize(loc, scale)
            elif effective_kind_of(c) == L.EUCLIDEAN:
                q_z = EuclideanNormal(loc, scale)
                p_z = EuclideanNormal(torch.zeros_like(loc), torch.ones_like(scale))
            else:
                man = c.manifold.manifold if self._kinds[i] == L.UNIVERSAL else c.manifold
                q_z = WrappedNormal(loc, scale, man)
                p_z = WrappedNormal(man.mu_0(loc.shape, device=loc.device, dtype=loc.dtype), torch.ones_like(scale), man)
            res.append(Reparametrized(q_z, p_z, z, ws.kl[:, i], ws.eps[:, d.eps_off:d.eps_off + d.n]))
        return res

    @torch.no_grad()
    def forward(self, x: Tensor, eps: Optional[Tensor] = None, beta: float = 1.0):
        """vae.py:69-80 -> (List[Reparametrized], concat_z, x_).  Also leaves the ELBO statistics of this batch in
        the stats vector (see compute_batch_stats)."""
        self._sync_radii()
        x = x.to(self.device, non_blocking=True)
        ws = self._workspace(x.shape[0])
        draw = self._stage(ws, x, eps)
        if ws.logits is None:
            ws.logits = torch.empty(ws.B, self.in_dim, device=self.device)
        self._forward_kernels(ws, beta, train=False, want_mu_sigma=True, logits=ws.logits, draw_eps=draw)
        if draw:
            ops.counter_add(self._bin_ctr)  # the next call draws fresh noise
        self._last_ws = ws
        return self._reparametrized(ws), ws.z, ws.logits

    @torch.no_grad()
    def compute_batch_stats(self, x_mb: Tensor, x_mb_: Tensor, reparametrized: List[Reparametrized], beta: float,
                            likelihood_n: int = 0) -> BatchStats:
        """vae.py:125-147: bce row sums of the given logits + the per-component KL of `reparametrized` -> BatchStats."""
        if x_mb.dtype == torch.uint8:
            # raw pixels: the targets are the binarised batch (image_reconstruction.py:37-53) — the very one forward()
            # produced when it ran on this batch just before (eval.py:95-96, train.py:236-238), else the 0.5 threshold
            ws = self._ws.get(x_mb.shape[0])
            if ws is not None and ws.u8 and x_mb_.data_ptr() == (ws.logits.data_ptr() if ws.logits is not None else 0):
                x_mb = ws.x.clone()
            else:
                x_mb = ops.binarize(x_mb.to(self.device).reshape(x_mb.shape[0], -1).contiguous(), dynamic=False,
                                    invert=self.binarize_invert)
        x_mb = x_mb.to(self.device).float().contiguous()
        bce, _ = ops.recon_loss(self.recon_kind, x_mb_.float().contiguous(), x_mb)
        kl = torch.stack([r.kl for r in reparametrized], dim=-1).contiguous()
        vec = ops.elbo_reduce(bce, kl, beta)
        likelihood = self.log_likelihood(x_mb, n=likelihood_n) if likelihood_n else None
        return BatchStats(vec, beta, bce, kl, likelihood)

    # ------------------------------------------------------------------------------------------ IWAE log-likelihood
    iwae_chunk_rows = 1 << 17  # rows (samples x batch) decoded per launch group; bounds the activation workspace

    @torch.no_grad()
    def log_likelihood(self, x: Tensor, n: int = 500, eps: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
        """vae.py:82-123: importance-weighted Monte-Carlo estimate with n samples per row ->
        (log_p_x [B], mi [B], cov_norm scalar).  `eps` [n, B, sum(n_i)] optionally supplies the standard-normal draws
        (component order = column order, as everywhere else).

        The encoder and the heads run ONCE (the reference does too, :93); per chunk of samples:
        mvae_iwae_latent (z and sum_c log q - log p for every component, from the head pre-activations) ->
        fc_d0 + relu -> logits GEMM with the reconstruction row sums in its epilogue (targets indexed modulo B: no
        x.repeat((n, 1, 1)), no [n, B, D] logits in HBM) ; then ONE streaming logsumexp over the sample axis for both
        estimates, and cov_norm from sum_s z (the sample mean commutes with the bilinear form of :119-121)."""
        B, D, H, P, Sd, Sn = x.shape[0], self.in_dim, self.h_dim, self.desc.ld_ml, self.desc.ld_z, self.desc.ld_eps
        self._sync_radii()
        ws = self._workspace(B)
        self._stage_x(ws, ws.slot, x)
        if self._planes_stale:
            self.refresh_weight_planes()
        self._input_planes(ws, train=False)
        self._gemm("e0_fwd", ws.xp, self.We0p, B, H, D, epilogue=L.EPI_BIAS_RELU, bias=self.fc_e0.bias.data,
                   out_planes=ws.hp)
        if self.latent_gemm:
            ws.ml.zero_()
            self._heads_gemm(ws)
        else:
            ops.skinny_rowdot((ws.hp, 3), self.Wh, H, 1, K=H, N=P, bias=self.bh, out=ws.ml)
        nc = max(1, min(n, self.iwae_chunk_rows // max(B, 1)))
        lw = self._iwae_workspace(B, nc, n)
        lw["recon"].zero_()
        lw["zsum"].zero_()
        epi = L.EPI_BCE_ROWSUM if self.recon_kind == "bce" else L.EPI_NLL_ROWSUM
        for s0 in range(0, n, nc):
            k = min(nc, n - s0)
            e = lw["eps"][:k]
            if eps is None:
                e.normal_()
            else:
                e.copy_(eps[s0:s0 + k], non_blocking=True)
            z = lw["z"][:k]
            ops.iwae_latent(self.desc, ws.ml, e, self._rflat, z, lw["diff"][s0:s0 + k], lw["zsum"])
            if self.latent_gemm:
                ops.split_planes(z.view(k * B, Sd), lw["zp"])
                self._gemm("iwae_d0", lw["zp"], self.Wd0p, k * B, H, Sd, epilogue=L.EPI_BIAS_RELU,
                           bias=self.fc_d0.bias.data, out_planes=lw["ddp"])
            else:
                ops.skinny_expand(z.view(k * B, Sd), self.fc_d0.weight.data, Sd, 1, K=Sd, N=H,
                                  bias=self.fc_d0.bias.data, act=ops.ACT_RELU, out_planes=lw["ddp"])
            self._gemm("iwae_logits", lw["ddp"], self.Wlp, k * B, D, H, epilogue=epi, bias=self.fc_logits.bias.data,
                       aux=ws.x, aux_rows=B, rowsum=lw["recon"][s0:s0 + k])
        log_p_x, mi = ops.iwae_reduce(lw["recon"], lw["diff"])
        cov_norm = ops.iwae_cov_norm(ws.x, lw["zsum"], n)
        return log_p_x, mi, cov_norm.reshape(())

    def _iwae_workspace(self, B: int, nc: int, n: int) -> dict:
        key = (B, nc, n)
        lw = self._iwae_ws.get(key) if hasattr(self, "_iwae_ws") else None
        if lw is None:
            f = dict(device=self.device, dtype=torch.float32)
            lw = {"eps": torch.empty(nc, B, self.desc.ld_eps, **f), "z": torch.empty(nc, B, self.desc.ld_z, **f),
                  "ddp": ops.PlaneBuf(nc * B, self.h_dim, 2, self.device), "recon": torch.zeros(n, B, **f),
                  "diff": torch.empty(n, B, **f), "zsum": torch.zeros(B, self.desc.ld_z, **f)}
            if self.latent_gemm:
                lw["zp"] = ops.PlaneBuf(nc * B, self.desc.ld_z, 3, self.device)
            self._iwae_ws = {key: lw}  # one likelihood workspace at a time (it can be hundreds of MB)
        return lw

    @torch.no_grad()
    def train_step(self, optimizer, x_mb: Tensor, beta: float, eps: Optional[Tensor] = None,
                   sync_stats: bool = True):
        """vae.py:149-166: zero_grad, forward, ELBO, backward, (clip: no 'curvature'-named params here), optimizer
        step, stats to floats.  Returns (BatchStatsFloat | BatchStats, (reparametrized, concat_z, x_mb_)); x_mb_
        (the logits) is not materialised in training — call forward() when it is needed."""
        ws = self._workspace(x_mb.shape[0])
        draw = self._stage(ws, x_mb, eps)
        self._step_kernels(optimizer, ws, beta, draw)
        self._last_ws = ws
        # Outputs of vae.py:166 = (reparametrized, concat_z, x_mb_); the first and the last are built on first use
        # from this step's buffers (valid until the next step on the same batch size)
        out = (LazyReparametrized(lambda: self._reparametrized_of(ws)), ws.z, LazyTensor(lambda: self.decode(ws.z)))
        if sync_stats:
            h = self._stats_wire.cpu().tolist()  # ONE device->host copy (the reference: 3 + C .item() syncs)
            if self.check_finite and int(ws.flag.item()) != 0:
                raise FloatingPointError("non-finite latent sample or KL term (device flag set by mvae_pm_forward)")
            return self._floats_of(h, beta), out
        # not synchronised: the device vector itself (valid until the next step overwrites it; clone() to keep it)
        return BatchStats(self._stats_report, beta), out

    def _floats_of(self, h: List[float], beta: float) -> BatchStatsFloat:
        """Host copy of the statistics wire -> BatchStatsFloat; raises if the data-parallel step reported a peer
        that never arrived (the parameters are frozen from that step on: mvae_dp_step)."""
        n = 3 + self.desc.C
        if len(h) > n and h[n] != 0.0:
            raise RuntimeError(f"data-parallel step failed: a peer did not arrive within the time limit (phase "
                               f"{int(h[n])}); parameters have not been updated since — all ranks must run the same "
                               "number of steps (MVAE_DP_TIMEOUT_S sets the limit)")
        return BatchStatsFloat(h[0], h[1], h[2], h[3:n], beta)

    def _step_kernels(self, optimizer, ws: _Workspace, beta: float, draw_eps: bool = False) -> None:
        self._sync_radii()
        fused = isinstance(optimizer, FusedCurvatureOptimizer)
        # the early launch is THIS optimizer's; held only while its step is being enqueued (a model attribute that kept
        # the bound method would close a model <-> optimizer reference cycle, and cyclic garbage holding CUDA graphs is
        # collected at arbitrary points — e.g. in the middle of a later graph capture)
        self._early_step = optimizer.step_early if fused else None
        try:
            self._run_step(optimizer, ws, beta, draw_eps, fused)
        finally:
            self._early_step = None

    def _run_step(self, optimizer, ws: _Workspace, beta: float, draw_eps: bool, fused: bool) -> None:
        if fused and self.use_cuda_graph:
            self._graphed_step(optimizer, ws, beta, draw_eps)
        else:
            if not fused:
                optimizer.zero_grad()
            self._forward_kernels(ws, beta, train=True, want_mu_sigma=self.train_statistics, logits=None,
                                     draw_eps=draw_eps)
            # (the early optimizer launch belongs to the fused optimizer; with an all-reduce hook the gradients are not
            # final before the hook has run)
            self._backward_kernels(ws, beta, early=fused and self._grad_hook is None, advance=True)
            if self._grad_hook is not None:
                self._grad_hook(self._bucket)  # data-parallel: one SUM all-reduce over [grads | radius grads | stats]
            if not fused:
                self._attach_grads()
                if self._clip_mask is not None:
                    ops.clip_grad_norm(self._gradius, self._clip_mask, 1.0)
            optimizer.step()
            self._join_pending()
            self._planes_stale = not getattr(optimizer, "planes_fresh", False)
            if self._push_stats:
                self._push_stats_kernel()

    @torch.no_grad()
    def train_epoch(self, optimizer, batches, beta: float, eps_batches=None) -> List[BatchStatsFloat]:
        """The batch loop of Trainer._train_epoch (train.py:198-199, 209-210): `for x_mb, y_mb in train_data:
        stats, _ = model.train_step(optimizer, x_mb, beta)` -> list of per-batch BatchStatsFloat, with the host<->device
        traffic of that loop pipelined instead of serialised with the kernels:

          * batch i+1 is copied host->device on a copy stream into the idle input slot while step i runs,
          * the ELBO statistics of step i leave through an asynchronous device->host copy into a pinned ring and are
            converted to floats when the ring wraps / at the end (the reference blocks on 3+C .item() calls per step).

        `batches` yields x_mb or (x_mb, y_mb) like the reference's DataLoader; host tensors should be pinned for the
        copies to overlap.  `eps_batches` optionally supplies the noise (tests)."""
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._slot_ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._slot_free = [torch.cuda.Event(), torch.cuda.Event()]
            for ev in self._slot_ready + self._slot_free:
                ev.record(main)   # torch creates the cudaEvent lazily: the raw handles are used below
        copy = self._copy_stream
        results: List[BatchStatsFloat] = []
        # The stream / event / copy calls of the loop go straight to the CUDA runtime (mvae_rt_*): through torch each
        # of them costs 10-20 us of host time (stream context switches), ~0.2 ms per step in total — more than the
        # step's kernels take, i.e. the end-to-end loop was host-bound.
        rt = L.lib()
        vp = ctypes.c_void_p
        main_h, copy_h = vp(main.cuda_stream), vp(copy.cuda_stream)
        ready_h = [vp(e.cuda_event) for e in self._slot_ready]
        free_h = [vp(e.cuda_event) for e in self._slot_free]
        cap = self.stats_ring_capacity

        def take(item):
            x = item[0] if isinstance(item, (tuple, list)) else item
            return x

        def prefetch(ws, slot, x, first):
            if not first:
                rt.mvae_rt_stream_wait_event(copy_h, free_h[slot])  # the step that last read this slot has finished
            else:
                copy.wait_stream(main)
            n = ws.B * self.in_dim
            if (not x.is_cuda) and x.is_contiguous() and x.numel() == n and x.dtype in (torch.uint8, torch.float32):
                if x.dtype == torch.uint8:
                    ws.x8  # allocate on first use
                    dst, nbytes, ws._u8[slot] = ws.x8buf[slot], n, True
                else:
                    dst, nbytes, ws._u8[slot] = ws.xbuf[slot], 4 * n, False
                ws.adopted = None
                L.check(rt.mvae_rt_memcpy_async(vp(dst.data_ptr()), vp(x.data_ptr()), nbytes, copy_h), "H2D copy")
            else:   # device tensors, other dtypes, views: the general path
                with torch.cuda.stream(copy):
                    self._stage_x(ws, slot, x)
            rt.mvae_rt_event_record(ready_h[slot], copy_h)

        self._ensure_stats_ring()
        state = {"i": 0}

        def fetch():
            """Statistics of the steps of this epoch not fetched yet: one device->host copy of the ring."""
            i = state["i"]
            if i <= len(results):
                return
            rows = self._stats_ring_dev.cpu().tolist()   # synchronises: every enqueued step has finished
            base = self._ring_count - i                  # value of the device counter when this epoch started
            for j in range(len(results), i):
                results.append(self._floats_of(rows[(base + j) % cap], beta))

        it = iter(batches)
        eps_it = iter(eps_batches) if eps_batches is not None else None
        nxt = next(it, None)
        if nxt is None:
            return results
        x = take(nxt)
        ws = self._workspace(x.shape[0])
        slot = ws.slot
        prefetch(ws, slot, x, first=True)
        self._push_stats = True   # the step's graph ends with mvae_ring_push
        try:
            while x is not None:
                i = state["i"]
                B = x.shape[0]
                nxt = next(it, None)
                x_next = take(nxt) if nxt is not None else None
                rt.mvae_rt_stream_wait_event(main_h, ready_h[slot])
                ws.slot = slot
                if x_next is not None and x_next.shape[0] == B:
                    prefetch(ws, 1 - slot, x_next, first=(i == 0))
                eps = next(eps_it) if eps_it is not None else self._eps_override
                if eps is not None:
                    ws.eps.copy_(eps, non_blocking=True)
                self._step_kernels(optimizer, ws, beta, eps is None)
                rt.mvae_rt_event_record(free_h[slot], main_h)
                state["i"] = i + 1
                self._ring_count += 1
                if state["i"] - len(results) >= cap:   # the ring is about to wrap: fetch it (a bubble once per `cap` steps)
                    fetch()
                if x_next is not None and x_next.shape[0] != B:  # ragged last batch: its own workspace, no overlap
                    ws = self._workspace(x_next.shape[0])
                    slot = ws.slot
                    prefetch(ws, slot, x_next, first=True)
                else:
                    slot = 1 - slot
                x = x_next
            fetch()
        finally:
            self._push_stats = False
        self._last_ws = ws
        if self.check_finite and int(ws.flag.item()) != 0:
            raise FloatingPointError("non-finite latent sample or KL term (device flag set by mvae_pm_forward)")
        return results

    # train_epoch: the step parks its statistics in a device-side ring (mvae_ring_push, the last node of its graph); the
    # loop fetches the ring with one device->host copy per `stats_ring_capacity` steps instead of one per step
    _push_stats = False
    stats_ring_capacity = 256

    def _ensure_stats_ring(self) -> None:
        n = self._stats_wire.numel()
        ring = getattr(self, "_stats_ring_dev", None)
        if ring is None or tuple(ring.shape) != (self.stats_ring_capacity, n):
            self._stats_ring_dev = torch.zeros(self.stats_ring_capacity, n, device=self.device)
            self._ring_ctr = torch.zeros(1, device=self.device, dtype=torch.int64)
            self._ring_count = 0   # host mirror of the device counter

    def _push_stats_kernel(self) -> None:
        ops.ring_push(self._stats_wire, self._stats_ring_dev, self._ring_ctr)

    _grad_hook = None
    _early_step = None  # FusedCurvatureOptimizer.step_early of the optimizer driving the step being enqueued
    use_cuda_graph = False
    fold_prologue = os.environ.get("MVAE_FOLD_PROLOGUE", "1") != "0"
    latent_gemm = False
    train_statistics = False  # True: train_step keeps q_z's loc / scale (Trainer --train_statistics, train.py:200-206)

    def _graphed_step(self, optimizer: "FusedCurvatureOptimizer", ws: _Workspace, beta: float,
                      draw_eps: bool = False) -> None:
        """Replay the whole step from CUDA graphs (launch-bound otherwise: ~30 kernels of a few microseconds).
        One graph holds forward + backward + optimizer step (+ refresh of the weight planes); with a gradient hook
        (the NCCL data-parallel path) it is split in two and the all-reduce runs between them, eagerly.  Keyed by everything baked into launch
        parameters: batch size, beta, and whether the curvature optimizers step."""
        # ... and every other value a launch bakes into its parameters (a changed learning rate must not replay the old one)
        adopted = ws.adopted
        key = (ws.B, ws.slot, beta, optimizer.curvature_step_enabled(), id(optimizer), draw_eps, ws.u8,
               self._grad_hook is None, optimizer.lr, optimizer.betas, optimizer.eps, optimizer.curvature_lr,
               self.binarize_seed, self.noise_seed, self.binarize_invert, self.check_finite, self.train_statistics,
               self.fused_latent, self.latent_gemm, 0 if adopted is None else adopted.data_ptr(), self._push_stats)
        entry = self._graphs.get(key)
        # Parameters changed outside the fused optimizer (load_state_dict, broadcast_parameters, an interleaved torch
        # optimizer): the graph's GEMMs read the weight PLANES, so they are rebuilt eagerly before any replay.
        if self._planes_stale:
            self.refresh_weight_planes()
        if entry is None:
            # warm-up on a side stream (lazy func attributes / module loading must not happen during capture)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._forward_kernels(ws, beta, train=True, want_mu_sigma=self.train_statistics, logits=None,
                                     draw_eps=draw_eps)
                self._backward_kernels(ws, beta, early=False)  # no optimizer step follows: nothing is exchanged
            torch.cuda.current_stream().wait_stream(side)
            n0 = ops.launch_count()
            one_graph = self._grad_hook is None  # nothing eager between backward and optimizer: ONE graph, one replay
            saved = optimizer.step_count

            def capture_opt():
                optimizer.step()
                self._join_pending()
                if not getattr(optimizer, "planes_fresh", False):
                    self.refresh_weight_planes()
                if self._push_stats:
                    self._push_stats_kernel()

            def capture_step():
                self._forward_kernels(ws, beta, train=True, want_mu_sigma=self.train_statistics, logits=None,
                                     draw_eps=draw_eps)
                self._backward_kernels(ws, beta, early=one_graph, advance=True)
                if one_graph:
                    capture_opt()

            ga = self._capture(capture_step)
            n1 = ops.launch_count()
            gb = None
            if not one_graph:
                gb = self._capture(capture_opt)
            optimizer.step_count = saved  # capture does not execute
            n2 = ops.launch_count()
            if len(self._graphs) >= 32:   # bound the captured graphs (each pins its adopted input tensor)
                self._graphs.pop(next(iter(self._graphs)))
            entry = self._graphs[key] = (ga, gb, n1 - n0, n2 - n1, adopted)
        ga, gb, la, lb, _ = entry
        ga.replay()
        if gb is not None:
            if self._grad_hook is not None:
                self._grad_hook(self._bucket)
            gb.replay()
        optimizer.step_count += 1
        ops.add_launches(la + lb)
        self._planes_stale = False

    @staticmethod
    def _capture(fn) -> "torch.cuda.CUDAGraph":
        """Capture fn() into a CUDA graph.  `thread_local` error mode: CUDA calls of OTHER threads (NCCL's watchdog, a
        monitoring thread) are none of the capture's business.  A capture that still fails — e.g. because the garbage
        collector released CUDA memory of an old model in the middle of it — is retried once after a full collection;
        nothing has executed at that point (capture only records)."""
        import gc
        for attempt in (0, 1):
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    fn()
                return g
            except RuntimeError:
                if attempt:
                    raise
                del g
                gc.collect()
                torch.cuda.synchronize()
        raise AssertionError("unreachable")

    def _attach_grads(self) -> None:
        """torch optimizers' zero_grad(set_to_none=True) drops .grad; re-attach the bucket views."""
        for name, p in [t[:2] for t in self._net_params()]:
            p.grad = self._grad_views[name]   # (a permuted view for tensors the flat buffer stores in another order)
        for i, rp in enumerate(self._radius_params):
            if rp is not None and rp.requires_grad:
                rp.grad = self._bucket[self._n_net + i]

    def _reparametrized_of(self, ws: _Workspace) -> List[Reparametrized]:
        """q_z / p_z / z / kl of the step that last ran on `ws`.  With train_statistics the step wrote loc and scale
        itself; otherwise they are recomputed here from the head pre-activations and the noise the step kept (with the
        radii as they are NOW: after a curvature step they differ from the step's own by lr * gradient)."""
        if not getattr(ws, "has_mu_sigma", False):
            ops.pm_forward(self.desc, ws.ml, ws.eps, self._rflat, want_mu_sigma=True,
                           out={"z": ws.z, "kl": ws.kl, "mu": ws.mu, "sigma": ws.sigma})
            ws.has_mu_sigma = True
        return self._reparametrized(ws)

    def reparametrized_of_last_step(self) -> List[Reparametrized]:
        return self._reparametrized_of(self._last_ws)


class FusedCurvatureOptimizer:
    """Kernel counterpart of Trainer.build_optimizer (train.py:327-360) + CurvatureOptimizer (utils.py:148-180):
    Adam(lr) over every network parameter (one launch over the flat bucket) and SGD(lr=1e-4) over the radii,
    stepped only when `should_do_curvature_step()` (reference: not fixed_curvature and epoch >= 10)."""

    def __init__(self, model: FusedFeedForwardVAE, learning_rate: float = 1e-3, fixed_curvature: bool = True,
                 should_do_curvature_step=lambda: False, curvature_lr: float = 1e-4, betas=(0.9, 0.999),
                 eps: float = 1e-8) -> None:
        self.model = model
        self.lr, self.betas, self.eps = learning_rate, betas, eps
        self.fixed_curvature = fixed_curvature
        self.curv_condition = should_do_curvature_step
        self.curvature_lr = curvature_lr
        self.exp_avg = torch.zeros_like(model._flat)
        self.exp_avg_sq = torch.zeros_like(model._flat)
        self.step_count = 0
        self.step_dev = torch.zeros(1, device=model._flat.device, dtype=torch.int32)  # 1-based after the first step
        self.param_groups = [{"params": [p for _, p in model._net_params()], "lr": learning_rate}]
        self._dp = self._dp_tail = self._dp_sync = None  # set by parallel.attach_p2p
        self._early_done = False
        self.dp_overlap = os.environ.get("MVAE_DP_OVERLAP", "1") != "0"
        self.dp_early_ctas = int(os.environ.get("MVAE_DP_EARLY_CTAS", "24"))
        self._done = torch.zeros(1, device=model._flat.device, dtype=torch.int32)
        # The early launch over fc_logits keeps ITS OWN step counter (advanced by its own last CTA): the launch that ends
        # the step advances step_dev when it finishes, and the two launches are not ordered against each other (the
        # model joins the early branch AFTER the final launch: _join_pending) — the early one must not read a counter
        # the final one may already have bumped.
        self.step_dev_early = torch.zeros(1, device=model._flat.device, dtype=torch.int32)
        self._done_early = torch.zeros(1, device=model._flat.device, dtype=torch.int32)
        self.planes_fresh = False  # True after a step that also refreshed the model's weight planes

    def zero_grad(self) -> None:
        pass  # the backward kernels overwrite / zero the bucket themselves

    def curvature_step_enabled(self) -> bool:
        return (not self.fixed_curvature) and bool(self.curv_condition())

    def hyper_parameters(self) -> tuple:
        """Everything step() bakes into launch parameters (part of the key of the model's CUDA graphs)."""
        return (float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.curvature_lr))

    def _plane_targets(self):
        """(targets the fused kernels refresh in place, [(weight, planes)] that need their own plane-split launch).
        The fused kernels refresh a weight's planes with 128-bit accesses: rows must be a multiple of 4 floats wide
        (in_dim = 50 of the BDP data is not); such a matrix gets its own launch after the update."""
        targets = self.model.plane_targets()
        late = [(t[3], t[2]) for t in targets if t[2].cols % 4]
        return [t[:3] for t in targets if t[2].cols % 4 == 0], late

    def _dp_launch(self, begin: int, end: int, channel: int, do_tail: bool, max_ctas: int = 0) -> None:
        m = self.model
        targets, late = self._plane_targets()
        inside = lambda off: begin <= off < end  # noqa: E731
        ops.dp_step(self._dp, m._n_net, begin, end, channel, do_tail, 2 * m.desc.C + 3, m.desc.C, self.exp_avg,
                    self.exp_avg_sq, self.lr, self.betas[0], self.betas[1], self.eps, self.step_dev, m._rflat,
                    self.curvature_lr if self.curvature_step_enabled() else 0.0, m._radius_mask, m._clip_mask, 1.0,
                    self._dp_tail, self._dp_sync, [t for t in targets if inside(t[0])], max_ctas=max_ctas)
        for w, buf in late:
            if inside(w.data_ptr() - m._flat.data_ptr() >> 2):
                ops.split_planes(w, buf)

    def _dp_ranges(self):
        """Float ranges of the parameter buffer exchanged by the launches of one data-parallel step."""
        m = self.model
        o = m.dp_early_begin()
        return [(o, m._n_net), (0, o)] if self.dp_overlap else [(0, m._n_net)]

    def step_early(self) -> None:
        """Update (under data parallelism: exchange + update) of fc_logits — half of the parameters — whose gradient is
        complete as soon as the logits weight-gradient and input-gradient GEMMs are: the model calls this on a side
        stream so that it runs UNDER the latent backward pass and the fc_e0 weight gradient (mvae_dp_step, channel 0,
        a few CTAs; on one GPU mvae_opt_step_fused over that range, without advancing the step counter)."""
        if not self.dp_overlap:
            return
        begin, end = self._dp_ranges()[0]
        if self._dp is not None:
            self._dp_launch(begin, end, 0, False, max_ctas=self.dp_early_ctas)
        else:
            self._local_launch(begin, end, last=False)
        self._early_done = True

    @property
    def early_is_independent(self) -> bool:
        """True when nothing orders the early launch against the one that ends the step (one GPU: disjoint parameter
        ranges, separate step counters) — the model may then join the early branch after the final launch."""
        return self._dp is None

    def _local_launch(self, begin: int, end: int, last: bool) -> None:
        """Adam over [begin, end) of the flat buffer + plane refresh of the GEMM weights inside; the launch that ends
        the step (`last`) also steps the radii and advances the device step counter (the early one: its own)."""
        m = self.model
        targets, late = self._plane_targets()
        inside = lambda off: begin <= off < end  # noqa: E731
        ops.opt_step_fused(m._flat[begin:end], m._gnet[begin:end], self.exp_avg[begin:end], self.exp_avg_sq[begin:end],
                           self.lr, self.betas[0], self.betas[1], self.eps, self.step_dev if last else self.step_dev_early,
                           self._done if last else self._done_early,
                           m._rflat if last else None, m._gradius, m._radius_mask,
                           self.curvature_lr if (last and self.curvature_step_enabled()) else 0.0,
                           [(t[0] - begin, t[1], t[2]) for t in targets if inside(t[0])])
        for w, buf in late:
            if inside(w.data_ptr() - m._flat.data_ptr() >> 2):
                ops.split_planes(w, buf)

    def step(self, closure=None) -> None:
        """The Adam step counter lives on the device (mvae_adam_step_dev), so this call can be captured in a CUDA
        graph and replayed."""
        m = self.model
        self.step_count += 1
        if self._dp is not None:
            # data parallel over NVLink peer memory: gradient reduce-scatter + Adam on this rank's slice + parameter
            # all-gather + the clip of the curvature gradients + the radii's SGD step, one kernel (mvae_dp_step) over
            # whatever step_early() has not taken already
            end = self._dp_ranges()[-1][1] if self._early_done else m._n_net
            self._early_done = False
            self._dp_launch(0, end, 1, True)
            m._planes_stale = False
            self.planes_fresh = True
            return
        if m._clip_mask is not None:
            ops.clip_grad_norm(m._gradius, m._clip_mask, 1.0)  # vae.py:161-163
        # Adam + the radii's SGD step + the refresh of the GEMM weight planes + the step counter: one launch over
        # whatever step_early() has not taken already (fixed radii receive no gradient: radius_mask)
        end = self._dp_ranges()[-1][1] if self._early_done else m._n_net
        if not self._early_done:
            # no early launch in this step: its counter follows along (an empty launch: the counter bump only)
            z = m._flat[0:0]
            ops.opt_step_fused(z, z, z, z, self.lr, self.betas[0], self.betas[1], self.eps, self.step_dev_early,
                               self._done_early, None, m._gradius, m._radius_mask, 0.0, [])
        self._early_done = False
        self._local_launch(0, end, last=True)
        m._planes_stale = False
        self.planes_fresh = True

    def state_dict(self):
        """Under the peer-memory data-parallel step every rank keeps only ITS 1/N slice of the moments current
        (mvae_dp_step): the slices are gathered here, so the dict is complete on every rank (collective call)."""
        m, v = self.exp_avg, self.exp_avg_sq
        if self._dp is not None:
            import torch.distributed as dist
            world, rank = int(self._dp.world), int(self._dp.rank)
            m, v = torch.zeros_like(m), torch.zeros_like(v)
            for begin, end in self._dp_ranges():
                n4 = (end - begin) // 4
                per = (n4 + world - 1) // world
                lo, hi = begin + 4 * min(n4, rank * per), begin + 4 * min(n4, (rank + 1) * per)
                m[lo:hi], v[lo:hi] = self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi]
            dist.all_reduce(m)
            dist.all_reduce(v)
        return {"exp_avg": m, "exp_avg_sq": v, "step": int(self.step_dev.item())}

    def load_state_dict(self, sd) -> None:
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_count = int(sd["step"])
        self.step_dev.fill_(self.step_count)
        self.step_dev_early.fill_(self.step_count)
