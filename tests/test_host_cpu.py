"""CPU tests of the host-side mirror of the reference's operator API (no kernels are launched): the component grammar
of mt/mvae/utils.py:78-140, the attribute / state_dict surface the reference's Trainer and checkpoints rely on
(checked against the parameter names the reference itself produced: tests/golden/model_*.npz), the descriptor the
kernels receive, and the bench contract of the CPU reference arm."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import load_golden, model_golden_names

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_component_grammar_and_attributes():
    from mvae_b200 import _lib as L
    from mvae_b200 import components
    comps = components.parse_components("2h3, s2,d4,p2,u3,e1", fixed_curvature=False)
    assert [c._shortcut() for c in comps] == ["h3", "h3", "s2", "d4", "p2", "u3", "e1"]
    # ambient dimensions: hyperboloid / sphere carry one more coordinate (component.py:114-167), the rest do not
    assert [c.dim for c in comps] == [4, 4, 3, 4, 2, 3, 1]
    assert [c.true_dim for c in comps] == [3, 3, 2, 4, 2, 3, 1]
    assert [c.kind for c in comps] == [L.HYPERBOLOID, L.HYPERBOLOID, L.SPHERE, L.PROJ_SPHERE, L.POINCARE, L.UNIVERSAL,
                                       L.EUCLIDEAN]
    # curvature parameter names: the reference's optimizer groups parameters by these substrings (train.py:329-333)
    assert [c.radius_parameter()[0] for c in comps] == ["_nradius", "_nradius", "_pradius", "_pradius", "_nradius",
                                                        "_curvature", None]
    assert all(c.radius_parameter()[1].requires_grad for c in comps[:-1])
    fixed = components.parse_components("h2,u2", fixed_curvature=True)
    assert not any(c.radius_parameter()[1].requires_grad for c in fixed)
    assert components.canonical_name(comps) == "d4,e1,2h3,p2,s2,u3"
    for bad in ("x2", "h0", "0h2", "h", ""):
        if bad == "":
            assert components.parse_components(bad, True) == []
            continue
        with pytest.raises((ValueError, NotImplementedError)):
            components.parse_components(bad, True)
    for c in comps:
        c.init_layers(16, scalar_parametrization=False)
        assert c.fc_mean.out_features == c.true_dim and c.fc_logvar.out_features == c.true_dim
        assert c.summary_name(3).startswith("comp_003_")
    # 'u': the manifold follows the sign of the curvature (universal.py:64-74), eps = 1e-6
    u = comps[5]
    for kappa, kind in ((-0.5, L.POINCARE), (0.3, L.PROJ_SPHERE), (0.0, L.EUCLIDEAN), (5e-7, L.EUCLIDEAN)):
        with torch.no_grad():
            u._curvature.fill_(kappa)
        assert u.effective_kind() == kind
        if kind != L.EUCLIDEAN:
            assert abs(float(u.manifold.radius) - abs(kappa) ** -0.5) < 1e-6


@pytest.mark.parametrize("name", model_golden_names())
def test_state_dict_keys_are_the_references(name):
    """A checkpoint of the reference loads: same parameter names and shapes, for every component
    letter incl. 'd' (_pradius) and 'u' (_curvature)."""
    from mvae_b200 import components, data, vae
    g, meta = load_golden(name)
    ref = {k[len("param."):]: v for k, v in g.items() if k.startswith("param.")}
    model = vae.FusedFeedForwardVAE(meta["h_dim"], components.parse_components(meta["sig"], meta["fixed_curvature"]),
                                    data.GenericDataset(1, meta["in_dim"], meta["recon"]),
                                    meta["scalar_parametrization"], device="cpu")
    sd = model.state_dict()
    assert sorted(sd) == sorted(ref)
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(np.asarray(v).shape), k
    res = model.load_state_dict({k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in ref.items()})
    assert not res.missing_keys and not res.unexpected_keys
    # the product-manifold descriptor the kernels get: concat order / offsets of vae.py:78
    off = 0
    for i, c in enumerate(model.components):
        d = model.desc.comp[i]
        assert (d.n, d.d, d.z_off) == (c.true_dim, c.dim, off)
        off += c.dim
    assert model.desc.ld_z == model.total_z_dim == off


def test_no_cpu_fallback_for_the_model():
    """The product path fails loudly without a CUDA device instead of computing on the host."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mvae_b200 import _lib, components, data, vae
    model = vae.FusedFeedForwardVAE(16, components.parse_components("h2,e2", True), data.GenericDataset(4, 8, "bce"),
                                    False, device="cpu")
    with pytest.raises(Exception):
        model.train_step(None, torch.zeros(4, 8), 1.0)
    # nothing under mvae_b200/ may import the oracle (test infrastructure)
    pkg = os.path.join(ROOT, "mvae_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, fn


def test_reference_arm_bench_line():
    """`bench.py --impl reference` (the CPU port on the host cores) prints ONE JSON line with the contract's keys."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--workload", "cfg1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "samples/s"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("MNIST e2")


def test_input_slots_keep_their_own_batch_kind():
    """train_epoch stages batch i+1 (host -> device) while step i is being enqueued: each input slot remembers whether
    it holds uint8 pixels (binarised on the device) or a float batch.  Host-side bookkeeping only (CPU tensors)."""
    from mvae_b200 import components, data, vae
    m = vae.FusedFeedForwardVAE(16, components.parse_components("h2,e2", True),
                                data.GenericDataset(4, 8, "bce", binary_inputs=True), False, device="cpu")
    ws = vae._Workspace(m, 4)
    xf = torch.rand(4, 8)
    x8 = (torch.rand(4, 8) * 255).to(torch.uint8)
    m._stage_x(ws, 0, x8)
    m._stage_x(ws, 1, xf)
    ws.slot = 0
    assert ws.u8 and torch.equal(ws.x8, x8)
    ws.slot = 1
    assert not ws.u8 and torch.equal(ws.x, xf)
    m._stage_x(ws, 1, x8.reshape(4, 2, 4))   # image-shaped batches are flattened like the reference's transform
    assert ws.u8 and torch.equal(ws.x8, x8)
