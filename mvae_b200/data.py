"""Host-side mirror of the reference's `VaeDataset` contract (mt/data/vae_dataset.py:22-44) for the hot path: only
`in_dim`, `img_dims`, `batch_size` and the reconstruction-loss KIND matter to the fused model
(BCE-with-logits: mt/data/image_reconstruction.py:81-82,142-143; unit-variance Gaussian NLL: mt/data/synthetic.py:161-162).
Real data loading stays with the reference's loaders; the classes here generate the synthetic, shape-faithful
batches SURVEY.md §8(d) prescribes for benchmarking."""
from typing import Optional, Tuple

import torch

from . import ops


class VaeDataset:

    recon_kind = "bce"
    binary_inputs = False  # True: every input value is 0 or 1 (exactly representable in one bf16 plane)

    def __init__(self, batch_size: int, in_dim: int, img_dims: Optional[Tuple[int, ...]]) -> None:
        self.batch_size = batch_size
        self._in_dim = in_dim
        self._img_dims = img_dims

    @property
    def img_dims(self):
        return self._img_dims

    @property
    def in_dim(self) -> int:
        return self._in_dim

    def reconstruction_loss(self, x_mb_: torch.Tensor, x_mb: torch.Tensor) -> torch.Tensor:
        """Row sums are what the model consumes (vae.py:131); returned as [B, 1] so that `.sum(dim=-1)` of the
        reference's call site gives the same [B] tensor."""
        rs, _ = ops.recon_loss(self.recon_kind, x_mb_.float().contiguous(), x_mb.float().contiguous())
        return rs.unsqueeze(-1)

    def synthetic_batch(self, batch_size: Optional[int] = None, seed: int = 0, device="cpu") -> torch.Tensor:
        raise NotImplementedError


class SyntheticMnistDataset(VaeDataset):
    """MNIST-shaped: binarised pixels with MNIST's mean density (the reference binarises with x > U(0,1),
    image_reconstruction.py:37-53)."""
    recon_kind = "bce"
    binary_inputs = True

    def __init__(self, batch_size: int) -> None:
        super().__init__(batch_size, in_dim=784, img_dims=(-1, 1, 28, 28))

    def synthetic_batch(self, batch_size=None, seed=0, device="cpu"):
        g = torch.Generator().manual_seed(seed)
        x = (torch.rand(batch_size or self.batch_size, self.in_dim, generator=g) < 0.1307).float()
        return x.to(device)


class SyntheticBdpDataset(VaeDataset):
    """BDP-shaped (in_dim 50, standardised values, Gaussian NLL; synthetic.py:138-162)."""
    recon_kind = "nll"

    def __init__(self, batch_size: int) -> None:
        super().__init__(batch_size, in_dim=50, img_dims=None)

    def synthetic_batch(self, batch_size=None, seed=0, device="cpu"):
        g = torch.Generator().manual_seed(seed)
        return torch.randn(batch_size or self.batch_size, self.in_dim, generator=g).to(device)


class SyntheticCifarDataset(VaeDataset):
    """CIFAR-shaped (in_dim 3072, values in [0,1], BCE with logits; image_reconstruction.py:116-143)."""
    recon_kind = "bce"

    def __init__(self, batch_size: int) -> None:
        super().__init__(batch_size, in_dim=3072, img_dims=(-1, 3, 32, 32))

    def synthetic_batch(self, batch_size=None, seed=0, device="cpu"):
        g = torch.Generator().manual_seed(seed)
        return torch.rand(batch_size or self.batch_size, self.in_dim, generator=g).to(device)


class GenericDataset(VaeDataset):

    def __init__(self, batch_size: int, in_dim: int, recon_kind: str = "bce", binary_inputs: bool = False) -> None:
        super().__init__(batch_size, in_dim, None)
        self.recon_kind = recon_kind
        self.binary_inputs = binary_inputs
