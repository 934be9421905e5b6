"""CPU tests: the C-ABI shared library loads without a GPU and exports every symbol include/mvae_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mvae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvae_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    from mvae_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built_lib):
    names = _declared_symbols()
    assert "mvae_pm_forward" in names and "mvae_gemm" in names and len(names) >= 17
    L = ctypes.CDLL(built_lib)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mvae_b200.h but not exported"


def test_binding_matches_header(built_lib):
    from mvae_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _declared_symbols()
    L = _lib.lib()
    assert L.mvae_abi_version() == _lib.ABI_VERSION
    assert L.mvae_strerror(0) == b"ok"
    assert b"invalid" in L.mvae_strerror(-1)
    assert ctypes.sizeof(_lib.PmDesc) == 16 + 32 * _lib.MAX_COMPONENTS


def test_integration_notes_name_every_entry_point():
    """INTEGRATION.md §2 maps every entry point of the header to the reference interface it replaces."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in _declared_symbols() if f"`{n}`" not in doc]
    assert not missing, missing


def test_desc_init_layout_matches_oracle(built_lib, oracle):
    """Host-only entry point: packed descriptor offsets are the oracle's (concat order of vae.py:78)."""
    from mvae_b200 import ops
    for sig, scalar in (("h2,s2,e2", False), ("h6,h6,s6,s6,e6", False), ("h2,s2,p3,e2", True), ("3e1,2p4", False)):
        a = ops.make_desc(sig, scalar_parametrization=scalar)
        b = oracle.make_desc(sig, scalar_parametrization=scalar)
        assert bytes(a) == bytes(b)
    with pytest.raises(Exception):
        ops.make_desc([9], [2])
    with pytest.raises(Exception):
        ops.make_desc([1], [0])


def test_compute_entry_points_fail_loudly_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mvae_b200 import _lib, ops
    with pytest.raises(_lib.MvaeError):
        ops.pm_forward(ops.make_desc("h2"), torch.zeros(4, 4), torch.zeros(4, 2), torch.ones(1))
