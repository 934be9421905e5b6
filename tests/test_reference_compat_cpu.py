"""CPU tests of "drops in under mt.mvae"  against the reference's OWN classes, imported unmodified from
/root/reference through oracle/ref_harness.py (skipped where the reference is absent, e.g. on the GPU box):

  * FusedFeedForwardVAE consumes the Component objects that mt.mvae.utils.parse_components builds, so the isinstance
    checks of Trainer._train_epoch's radius warm-up (train.py:189-194) hold;
  * that warm-up REBINDS `_pradius.data` / `_nradius.data`: the rebind reaches the flat radius vector the kernels read;
  * Trainer.build_optimizer (train.py:327-360) finds the same parameter groups by name;
  * state_dict keys / shapes equal those of the reference's FeedForwardVAE built from the same components.
Host-side bookkeeping only: no kernel runs here (the kernels' parity is the GPU tests' job)."""
import os
import sys

import pytest

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as rh  # noqa: E402

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    return rh.load_reference()


def _fused(ref, sig, fixed, in_dim=20, h_dim=16):
    from mt.mvae import utils
    from mvae_b200 import data, vae
    torch.manual_seed(0)
    comps = utils.parse_components(sig, fixed)          # the reference's own Component objects
    return vae.FusedFeedForwardVAE(h_dim, comps, data.GenericDataset(4, in_dim, "bce"), False, device="cpu")


def test_reference_components_are_consumed_as_they_are(ref):
    from mt.mvae import components as rc
    from mvae_b200 import _lib as L
    model = _fused(ref, "h2,s2,d3,p2,u2,e2", fixed=False)
    assert [type(c) for c in model.components] == [rc.HyperbolicComponent, rc.SphericalComponent,
                                                   rc.StereographicallyProjectedSphereComponent, rc.PoincareComponent,
                                                   rc.UniversalComponent, rc.EuclideanComponent]
    assert model._kinds == [L.HYPERBOLOID, L.SPHERE, L.PROJ_SPHERE, L.POINCARE, L.UNIVERSAL, L.EUCLIDEAN]
    assert [(model.desc.comp[i].n, model.desc.comp[i].d) for i in range(6)] == [(2, 3), (2, 3), (3, 3), (2, 2), (2, 2),
                                                                                (2, 2)]
    assert model.total_z_dim == sum(c.dim for c in model.components) == 15
    # parameters of the reference's modules now live in the flat buffer the kernels read
    base = model._flat.data_ptr()
    for c in model.components:
        assert base <= c.fc_mean.weight.data_ptr() < base + 4 * model._flat.numel()
    # a sampling procedure outside the hot path is refused, not silently replaced
    from mt.mvae.sampling import EuclideanConstantProcedure
    from mvae_b200 import data, vae
    odd = rc.SphericalComponent(2, True, sampling_procedure=EuclideanConstantProcedure)
    with pytest.raises(NotImplementedError):
        vae.FusedFeedForwardVAE(16, [odd], data.GenericDataset(4, 20, "bce"), False, device="cpu")


def test_state_dict_equals_the_reference_models(ref):
    sig = "h2,s2,d3,p2,u2,e2"
    fused = _fused(ref, sig, fixed=False)
    ref_model = rh.build_model(sig, 20, 16, False, False, "bce", 0, torch.float32)
    want = ref_model.state_dict()
    got = fused.state_dict()
    assert list(got) == list(want)
    for k in want:
        assert tuple(got[k].shape) == tuple(want[k].shape), k
        # same construction order under the same seed => same default initialisation (vae.py:55-57, ffnn_vae.py:35-40)
        assert torch.allclose(got[k].float(), want[k].float()), k
    fused.load_state_dict({k: v.float() * 0.5 for k, v in want.items()})
    assert torch.allclose(fused.fc_e0.weight, want["fc_e0.weight"].float() * 0.5)


def test_trainer_radius_warmup_and_optimizer_groups(ref):
    """Run the reference's Trainer code paths that touch the model object — the warm-up loop of _train_epoch
    (train.py:189-194, copied here statement for statement because the method also needs a DataLoader) and
    build_optimizer (train.py:327-360, called as is) — against the fused model."""
    from mt.mvae.models import train as rt
    model = _fused(ref, "h2,s2,d3,p2,e2", fixed=False)
    assert torch.equal(model._rflat, torch.ones(5))
    for epoch in (0, 4, 9):
        for c in model.components:                                   # train.py:190-194
            if isinstance(c, rt.StereographicallyProjectedSphereComponent) or isinstance(c, rt.SphericalComponent):
                c._pradius.data = torch.ones_like(c._pradius.data) * (11 - epoch)
            elif isinstance(c, rt.PoincareComponent) or isinstance(c, rt.HyperbolicComponent):
                c._nradius.data = torch.ones_like(c._nradius.data) * (11 - epoch)
        assert any(rp is not None and rp.data_ptr() != ptr for rp, ptr in zip(model._radius_params, model._radius_ptrs))
        model._sync_radii()                                          # what every kernel sequence starts with
        want = float(11 - epoch)
        assert model._rflat.tolist() == [want, want, want, want, 1.0]
        for rp, ptr in zip(model._radius_params, model._radius_ptrs):
            assert rp is None or rp.data_ptr() == ptr                # re-homed: optimizers step the kernels' storage
    # in-place writes land in the flat vector directly
    with torch.no_grad():
        model.components[0]._nradius.fill_(2.5)
    assert float(model._rflat[0]) == 2.5

    class _Stats:
        epoch, global_step = 12, 0

    class _T:
        pass

    trainer = _T()
    trainer.model, trainer.stats = model, _Stats()
    type(trainer).epoch = property(lambda self: self.stats.epoch)
    opt = rt.Trainer.build_optimizer(trainer, learning_rate=1e-3, fixed_curvature=False)
    groups = opt.param_groups
    net, neg, pos = groups[0]["params"], groups[1]["params"], groups[2]["params"]
    assert len(neg) == 2 and len(pos) == 2                            # h, p | s, d
    assert all(p.numel() == 1 for p in neg + pos)
    assert len(net) == len(list(model.parameters())) - 4
    assert opt.curv_condition() is True


def test_train_step_output_types(ref):
    """LazyReparametrized / LazyTensor behave like the list / tensor Trainer._train_epoch unpacks (train.py:198)."""
    from mvae_b200 import vae
    calls = []
    lt = vae.LazyTensor(lambda: calls.append(1) or torch.arange(6.0).reshape(2, 3))
    assert not calls
    assert tuple(lt.shape) == (2, 3) and len(calls) == 1
    assert torch.equal(torch.sigmoid(lt), torch.sigmoid(torch.arange(6.0).reshape(2, 3))) and len(calls) == 1
    assert torch.equal(lt[1], torch.tensor([3.0, 4.0, 5.0])) and len(lt) == 2
    lr = vae.LazyReparametrized(lambda: calls.append(2) or ["a", "b"])
    assert list(zip(["x", "y"], lr)) == [("x", "a"), ("y", "b")] and lr[1] == "b" and len(lr) == 2
    assert calls.count(2) == 1


def test_conv_model_state_dict_equals_the_reference_models(ref):
    """FusedConvolutionalVAE stores conv filters in GEMM order ([Co, ky, kx, Ci]) behind parameters that keep the
    reference's names, SHAPES and values (permuted views): same state_dict as ConvolutionalVAE (conv_vae.py:47-55)
    under the same seed, checkpoints load both ways, gradients are exposed in the reference's layout."""
    from mt.mvae import utils
    from mvae_b200 import conv_vae, data
    sig = "h2,s2,e2"
    ref_model = rh.build_model(sig, 3072, 8192, False, False, "bce", 0, torch.float32, architecture="conv")
    want = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
    torch.manual_seed(0)
    fused = conv_vae.FusedConvolutionalVAE(8192, utils.parse_components(sig, False), data.GenericDataset(4, 3072, "bce"),
                                           False, device="cpu")
    got = fused.state_dict()
    assert list(got) == list(want)
    for k in want:
        assert tuple(got[k].shape) == tuple(want[k].shape), k
        assert torch.equal(got[k].float(), want[k].float()), k
    # the flat buffer holds the filters channel-innermost: rows of the GEMM operand
    o, n = fused._slices["e1.weight"]
    master = fused._flat[o:o + n].view(128, 4, 4, 64)
    assert torch.equal(master, want["e1.weight"].permute(0, 2, 3, 1))
    assert torch.equal(fused._W["e1"], want["e1.weight"].permute(0, 2, 3, 1).reshape(128, 1024))
    o, n = fused._slices["d1.weight"]
    assert torch.equal(fused._flat[o:o + n].view(128, 4, 4, 256), want["d1.weight"].permute(0, 2, 3, 1))
    # checkpoints: reference -> fused (in place into the permuted views) and back
    fused.load_state_dict({k: v * 2 for k, v in want.items()})
    assert torch.equal(fused.e2.weight, want["e2.weight"] * 2) and fused._planes_stale
    ref_model.load_state_dict({k: v.clone() for k, v in fused.state_dict().items()})
    assert torch.equal(ref_model.d2.weight, want["d2.weight"] * 2)
    # gradients live in the bucket in master order and are seen through the same permutation
    fused._bucket.copy_(torch.arange(fused._bucket.numel(), dtype=torch.float32))
    o, n = fused._slices["e0.weight"]
    assert torch.equal(fused.e0.weight.grad, fused._bucket[o:o + n].view(64, 4, 4, 3).permute(0, 3, 1, 2))
    fused._attach_grads()
    assert fused.e0.weight.grad.shape == (64, 3, 4, 4) and fused.e0.weight.grad.data_ptr() == fused._bucket[o:].data_ptr()
    # Trainer.build_optimizer sees reference-shaped parameters
    assert {tuple(p.shape) for n_, p in fused.named_parameters() if n_.endswith("d3.weight")} == {(64, 3, 4, 4)}
