"""mvae_b200 — B200-native (sm_100a) hot path of the mixed-curvature VAE (oskopek/mvae).

Layout:  csrc/ (CUDA kernels + the C ABI of include/mvae_b200.h), _lib.py (ctypes binding), ops.py (tensor-level
wrappers), and the host-side mirror of the reference's operator API (manifolds.py, distributions.py, components.py,
vae.py) that drops in under mt.mvae.  There is no CPU fallback: the CUDA library is the product.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
