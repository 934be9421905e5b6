"""GPU parity tests of the whole hot path (FusedFeedForwardVAE.forward / train_step) against the golden vectors the
reference itself produced (tests/golden/model_*.npz: ModelVAE.forward + compute_batch_stats + backward in float64,
vae.py:69-80,125-160) and against the CPU oracle at the BASELINE config shapes."""
import numpy as np
import pytest

from helpers import load_golden, model_golden_names, normwise

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-4       # BASELINE.json north_star: 1e-4 relative (normwise per tensor)
TOL_SUM = 1e-5   # ELBO / bce / kl sums


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mvae_b200 import _lib
    _lib.lib()
    return torch.device("cuda:0")


def _build(meta, params, dev):
    from mvae_b200 import components, data, vae
    comps = components.parse_components(meta["sig"], meta["fixed_curvature"])
    ds = data.GenericDataset(1, meta["in_dim"], meta["recon"])
    model = vae.FusedFeedForwardVAE(meta["h_dim"], comps, ds, meta["scalar_parametrization"], device=dev)
    sd = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in params.items()}
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model


@pytest.mark.parametrize("name", model_golden_names())
def test_forward_and_gradients_match_reference(dev, name):
    from mvae_b200 import vae
    g, meta = load_golden(name)
    params = {k[len("param."):]: v for k, v in g.items() if k.startswith("param.")}
    model = _build(meta, params, dev)
    # state_dict keys and shapes are the reference's (checkpoint compatibility)
    sd = model.state_dict()
    assert sorted(sd) == sorted(params)
    for k in params:
        assert tuple(sd[k].shape) == tuple(np.asarray(params[k]).shape), k
    x = torch.from_numpy(g["x"].astype(np.float32))
    eps = torch.from_numpy(g["eps"].astype(np.float32)).to(dev)
    beta = meta["beta"]
    rep, concat_z, logits = model.forward(x, eps=eps, beta=beta)
    assert normwise(concat_z.cpu().numpy(), g["z"]) < TOL
    assert normwise(logits.cpu().numpy(), g["logits"]) < TOL
    assert normwise(torch.cat([r.q_z.loc for r in rep], -1).cpu().numpy(), g["mu"]) < TOL
    assert normwise(torch.cat([r.q_z.scale for r in rep], -1).cpu().numpy(), g["sigma"]) < TOL
    assert normwise(torch.stack([r.kl for r in rep], -1).cpu().numpy(), g["kl"]) < TOL
    # bit-exact index work: concat order / offsets of vae.py:78
    off = 0
    for r in rep:
        w = r.z.shape[-1]
        assert torch.equal(r.z, concat_z[:, off:off + w])
        off += w
    stats = model.compute_batch_stats(x, logits, rep, beta=beta)
    assert normwise(stats._bce.cpu().numpy(), g["bce"]) < TOL
    sf = stats.convert_to_float()
    assert abs(sf.elbo - g["elbo"]) < TOL_SUM * abs(g["elbo"])
    assert abs(sf.bce - g["bce_sum"]) < TOL_SUM * abs(g["bce_sum"])
    assert abs(sf.kl - g["kl_sum"]) < TOL_SUM * abs(g["kl_sum"]) + 1e-5
    # one train step with a no-op optimizer: gradients of -ELBO for every parameter

    class NoOpt:
        def zero_grad(self):
            pass

        def step(self):
            pass

    bs, _ = model.train_step(NoOpt(), x, beta, eps=eps)
    assert abs(bs.elbo - g["elbo"]) < TOL_SUM * abs(g["elbo"])
    # 'u' components: train_step clips the 2-norm of the "curvature"-named gradients to 1 (vae.py:161-163); the golden
    # gradients come from a plain backward(), so the same clip is applied to them here
    curv = [k for k, _ in model.named_parameters() if "curvature" in k]
    clip = min(1.0, 1.0 / (float(np.sqrt(sum(float(g["grad." + k]) ** 2 for k in curv))) + 1e-6)) if curv else 1.0
    for k, p in model.named_parameters():
        ref = g["grad." + k] * (clip if k in curv else 1.0)
        if "radius" in k:
            if meta["fixed_curvature"]:
                continue
            got = p.grad.detach().cpu().numpy()
            assert abs(got - ref) < 2e-4 * max(1.0, abs(ref)), (k, got, ref)
            continue
        got = p.grad.detach().cpu().numpy()
        assert normwise(got, ref) < 2e-4, (k, normwise(got, ref))


@pytest.mark.parametrize("sig,B,D,H,recon,fixed", [
    ("h2,s2,e2", 4096, 784, 400, "bce", False),          # BASELINE cfg2
    ("e2", 128, 784, 400, "bce", True),                  # cfg1
    ("h2", 2048, 50, 400, "nll", False),                 # cfg4a (reduced batch for the CPU oracle)
    ("p2", 2048, 50, 400, "nll", False),                 # cfg4b
    ("h6,h6,s6,s6,e6", 1000, 784, 400, "bce", False),    # cfg3 model, ragged batch
])
@pytest.mark.parametrize("mode", ["fused", "skinny", "gemm"])
def test_train_step_vs_oracle(dev, oracle, sig, B, D, H, recon, fixed, mode):
    """Full step at the BASELINE shapes: loss, statistics and every gradient against the float64 oracle, with the
    latent block as one fused kernel per direction (the training path of narrow products), as separate CUDA-core
    kernels, and with its dense layers on the tensor cores (the training path of wide products such as cfg3)."""
    from mvae_b200 import components, data, vae
    torch.manual_seed(0)
    comps = components.parse_components(sig, fixed)
    model = vae.FusedFeedForwardVAE(H, comps, data.GenericDataset(B, D, recon), False, device=dev)
    if mode == "gemm":
        if not model.latent_gemm:
            pytest.skip("narrow product: the latent block stays fused")
        assert not model.fused_latent
    else:
        model.latent_gemm = False
        model.fused_latent = mode == "fused"
    fused_latent = mode == "fused"
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(B, D, generator=g) < 0.1307).float() if recon == "bce" else torch.randn(B, D, generator=g)
    eps = torch.randn(B, model.desc.ld_eps, generator=g)
    params = {k: v.detach().cpu().double().numpy() for k, v in model.state_dict().items()}
    ovae = oracle.OracleVAE(sig, D, H, recon, False)

    class NoOpt:
        def zero_grad(self):
            pass

        def step(self):
            pass

    bs, _ = model.train_step(NoOpt(), x, 0.8, eps=eps.to(dev))
    ws = model._last_ws
    fwd = ovae.step(params, x.double().numpy(), eps.double().numpy(), beta=0.8, backward=False)
    # The two relus make the map parameters -> gradients discontinuous: a hidden unit whose pre-activation is within
    # float32 rounding of zero may fall on the other side than in the float64 oracle (expected ~1 of the 2*B*H units
    # per step at these sizes for ANY float32 implementation; relu'(0) is a convention).  The device's decisions are
    # therefore checked on their own — every unit that differs from the oracle must sit at the kink, and there may
    # only be a handful — and the oracle's backward pass then takes the same decisions, so that every gradient is
    # compared at the tight bar.
    decisions = {}
    h_dev = ws.h32 if fused_latent else ws.hp.to_float()  # the fused latent block reads h as fp32, not as planes
    for name, act, pre in (("h", h_dev, fwd["h_pre"]), ("dd", ws.ddp.to_float(), fwd["dd_pre"])):
        on = act.cpu().numpy() > 0
        diff = on != (pre > 0)
        assert diff.sum() <= 4, (name, int(diff.sum()))
        assert np.all(np.abs(pre[diff]) <= 2e-6 * np.abs(pre).max()), (name, np.abs(pre[diff]).max())
        decisions[name] = on
    ref = ovae.step(params, x.double().numpy(), eps.double().numpy(), beta=0.8, relu_decisions=decisions)
    assert abs(bs.elbo - ref["elbo"]) < TOL_SUM * abs(ref["elbo"])
    assert abs(bs.bce - ref["bce_sum"]) < TOL_SUM * abs(ref["bce_sum"])
    assert abs(bs.kl - ref["kl_sum"]) < TOL_SUM * abs(ref["kl_sum"]) + 1e-3
    np.testing.assert_allclose(bs.component_kl, ref["kl_comp"], rtol=2e-5, atol=1e-2)
    if not fused_latent:  # the fused latent kernels keep gz / gml in shared memory
        assert normwise(ws.gz.cpu().numpy(), ref["gz"]) < TOL
        assert normwise(ws.gml.cpu().numpy(), ref["gml"]) < TOL
    for k, p in model.named_parameters():
        if k not in ref["grads"]:
            continue
        got = p.grad.detach().cpu().numpy()
        r = np.asarray(ref["grads"][k], dtype=np.float64)
        if got.ndim:
            err_max = normwise(got, r)
            err_fro = float(np.linalg.norm(got - r) / max(np.linalg.norm(r), 1e-30))
        else:
            err_max = err_fro = abs(got - r) / max(1.0, abs(r))
        assert err_fro < TOL and err_max < TOL, (k, err_fro, err_max)


def test_optimizer_step_matches_torch(dev):
    """train_step with the fused optimizer == train_step with torch.optim.Adam/SGD on the same gradients
    (Trainer.build_optimizer semantics, train.py:327-360)."""
    from mvae_b200 import components, data, vae
    B, D, H = 256, 64, 32
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(B, D, generator=g) < 0.3).float()
    eps_list = [torch.randn(B, 6, generator=g).to(dev) for _ in range(3)]
    models = []
    for _ in range(2):
        torch.manual_seed(5)
        m = vae.FusedFeedForwardVAE(H, components.parse_components("h2,s2,e2", False), data.GenericDataset(B, D, "bce"),
                                    False, device=dev)
        models.append(m)
    fused = vae.FusedCurvatureOptimizer(models[0], 1e-3, fixed_curvature=False, should_do_curvature_step=lambda: True)
    net = [p for k, p in models[1].named_parameters() if "radius" not in k]
    curv = [p for k, p in models[1].named_parameters() if "radius" in k]
    adam, sgd = torch.optim.Adam(net, lr=1e-3), torch.optim.SGD(curv, lr=1e-4)

    class Both:
        def zero_grad(self):
            adam.zero_grad()
            sgd.zero_grad()

        def step(self):
            adam.step()
            sgd.step()

    for e in eps_list:
        s0, _ = models[0].train_step(fused, x, 1.0, eps=e)
        s1, _ = models[1].train_step(Both(), x, 1.0, eps=e)
        assert abs(s0.elbo - s1.elbo) < 1e-5 * abs(s1.elbo)
    for (k0, p0), (k1, p1) in zip(models[0].named_parameters(), models[1].named_parameters()):
        assert k0 == k1
        assert torch.allclose(p0, p1, rtol=1e-4, atol=1e-6), k0
    # parameters actually moved, and the radii too
    assert abs(float(models[0].components[0]._nradius.detach()) - 1.0) > 0


def test_training_reduces_loss(dev):
    from mvae_b200 import components, data, vae
    torch.manual_seed(0)
    ds = data.SyntheticMnistDataset(512)
    model = vae.FusedFeedForwardVAE(400, components.parse_components("h2,s2,e2", True), ds, False, device=dev)
    opt = vae.FusedCurvatureOptimizer(model, 1e-3, fixed_curvature=True)
    x = ds.synthetic_batch(seed=0)
    first = last = None
    for i in range(60):
        bs, _ = model.train_step(opt, x, 1.0)
        assert np.isfinite(bs.elbo)
        first = bs.elbo if first is None else first
        last = bs.elbo
    assert last > first + 0.2 * abs(first)  # ELBO is a sum over the batch; it must rise markedly when overfitting one batch


def test_uint8_batches_binarised_on_device(dev, oracle):
    """mvae_binarize (ImageDynamicBinarization, image_reconstruction.py:37-53, moved to the device): bit-exact against
    the reference's comparison `ToTensor(x) > U` for supplied draws and against `> 0.5` in evaluation mode; the Philox
    path is deterministic per (seed, step), fresh per step, and unbiased; and a train step fed uint8 pixels equals
    the step fed the float batch the kernel produced (up to the order of the atomic reductions)."""
    from mvae_b200 import components, data, ops, vae
    g = torch.Generator().manual_seed(0)
    for B, D in ((64, 784), (37, 19)):   # vectorised and general kernel
        px = torch.randint(0, 256, (B, D), generator=g, dtype=torch.int32).to(torch.uint8)
        u = torch.rand(B, D, generator=g)
        v = px.to(torch.float32).div(255)                     # torchvision ToTensor
        ref_dyn = (v.double() > u.double()).float()           # float64 default dtype of the reference (run.py:77)
        ref_fix = (v > 0.5).float()
        ref_inv = ((1 - v).double() > u.double()).float()
        xp = ops.PlaneBuf(B, D, 1, dev, ones_col=True)
        x = ops.binarize(px.to(dev), planes=xp, x=torch.empty(B, D, device=dev), u=u.to(dev))
        assert torch.equal(x.cpu(), ref_dyn)
        assert torch.equal(xp.t[0, :, :D].float().cpu(), ref_dyn) and bool((xp.t[0, :, D] == 1).all())
        assert torch.equal(ops.binarize(px.to(dev), dynamic=False).cpu(), ref_fix)
        assert torch.equal(ops.binarize(px.to(dev), u=u.to(dev), invert=True).cpu(), ref_inv)
        ctr = torch.zeros(1, dtype=torch.int64, device=dev)
        a = ops.binarize(px.to(dev), seed=7, offset_dev=ctr)
        b = ops.binarize(px.to(dev), seed=7, offset_dev=ctr)
        assert torch.equal(a, b)
        ctr += 1
        c = ops.binarize(px.to(dev), seed=7, offset_dev=ctr)
        assert not torch.equal(a, c)
        assert not torch.equal(a, ops.binarize(px.to(dev), seed=8, offset_dev=ctr - 1))
    # unbiased: P(x = 1) = v for every pixel value (large sample, 5 sigma)
    Bn = 1 << 16
    px = torch.full((Bn, 8), 0, dtype=torch.uint8)
    levels = torch.tensor([0, 1, 64, 100, 128, 200, 254, 255], dtype=torch.uint8)
    px[:] = levels
    ctr = torch.zeros(1, dtype=torch.int64, device=dev)
    freq = ops.binarize(px.to(dev), seed=3, offset_dev=ctr).mean(0).cpu().double()
    p = levels.double() / 255
    assert bool(((freq - p).abs() <= 5 * (p * (1 - p) / Bn).sqrt() + 1e-12).all()), (freq, p)
    # model level
    torch.manual_seed(0)
    B, D, H = 256, 784, 64
    mk = lambda: vae.FusedFeedForwardVAE(H, components.parse_components("h2,s2,e2", False),
                                         data.GenericDataset(B, D, "bce", binary_inputs=True), False, device=dev)
    m1 = mk()
    m2 = mk()
    m2.load_state_dict(m1.state_dict())
    px = torch.randint(0, 256, (B, D), generator=g, dtype=torch.int32).to(torch.uint8)
    eps = torch.randn(B, m1.desc.ld_eps, generator=g).to(dev)
    o1 = vae.FusedCurvatureOptimizer(m1, 1e-3, fixed_curvature=False, should_do_curvature_step=lambda: True)
    o2 = vae.FusedCurvatureOptimizer(m2, 1e-3, fixed_curvature=False, should_do_curvature_step=lambda: True)
    m1.binarize_seed = 11
    s1, _ = m1.train_step(o1, px, 1.0, eps=eps)
    x_float = m1._workspace(B).x.clone()   # what the kernel produced from the pixels
    assert set(x_float.unique().tolist()) <= {0.0, 1.0} and 0.3 < float(x_float.mean()) < 0.7
    s2, _ = m2.train_step(o2, x_float, 1.0, eps=eps)
    assert abs(s1.elbo - s2.elbo) <= 1e-6 * abs(s2.elbo) and abs(s1.bce - s2.bce) <= 1e-6 * abs(s2.bce)
    for (k, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
        assert normwise(a.cpu().numpy(), b.cpu().numpy()) < 1e-5, k
    # the next step draws different uniforms (device counter), evaluation uses the fixed threshold
    m1.train_step(o1, px, 1.0, eps=eps)
    assert not torch.equal(m1._workspace(B).x, x_float)
    m1.forward(px, eps=eps)
    assert torch.equal(m1._workspace(B).x.cpu(), (px.float().div(255) > 0.5).float())
    # pipelined epoch with uint8 batches and CUDA graphs: runs, finite, statistics per batch
    m1.use_cuda_graph = True
    out = m1.train_epoch(o1, [px.pin_memory()] * 5, 1.0)
    assert len(out) == 5 and all(np.isfinite(s.elbo) for s in out)
    assert int(m1._workspace(B).bin_ctr.item()) >= 6


def test_universal_components_train(dev, oracle):
    """'u' (universal.py): the kernels branch on sign(kappa) per launch; one step's loss and every gradient (incl.
    d/dkappa) against the float64 oracle; the fused optimizer (gradient clip of the curvature parameters, vae.py:161-163,
    then SGD) moves kappa exactly like torch's clip_grad_norm_ + SGD; a curvature that crosses the +-eps band switches
    manifolds without any re-setup."""
    from mvae_b200 import components, data, vae
    sig, B, D, H = "u2,u3,u2,e2", 512, 64, 32
    kappas = [-0.7, 0.9, 5e-7]
    torch.manual_seed(2)
    model = vae.FusedFeedForwardVAE(H, components.parse_components(sig, False), data.GenericDataset(B, D, "bce"), False,
                                    device=dev)
    with torch.no_grad():
        for c, k in zip(model.components, kappas):
            c._curvature.fill_(k)
    assert [c.effective_kind() for c in model.components] == [3, 4, 0, 0]
    g = torch.Generator().manual_seed(4)
    x = (torch.rand(B, D, generator=g) < 0.3).float()
    eps = torch.randn(B, model.desc.ld_eps, generator=g)
    params = {k: v.detach().cpu().double().numpy() for k, v in model.state_dict().items()}
    ref = oracle.OracleVAE(sig, D, H, "bce", False).step(params, x.double().numpy(), eps.double().numpy(), beta=1.0)

    class NoOpt:
        def zero_grad(self):
            pass

        def step(self):
            pass

    kap_before = [float(c._curvature.detach()) for c in model.components[:3]]
    bs, _ = model.train_step(NoOpt(), x, 1.0, eps=eps.to(dev))
    assert abs(bs.elbo - ref["elbo"]) < TOL_SUM * abs(ref["elbo"])
    gk = np.array([float(ref["grads"][f"components.{i}._curvature"]) for i in range(3)])
    assert gk[2] == 0.0 and abs(gk[0]) > 0 and abs(gk[1]) > 0   # no gradient reaches kappa on the Euclidean branch
    clip = min(1.0, 1.0 / (np.linalg.norm(gk) + 1e-6))
    for k, p in model.named_parameters():
        r = ref["grads"][k] * (clip if "curvature" in k else 1.0)
        got = p.grad.detach().cpu().numpy()
        assert normwise(got, r) < 3e-4, (k, normwise(got, r))
    # fused optimizer: clip + SGD(1e-4) on kappa
    opt = vae.FusedCurvatureOptimizer(model, 1e-3, fixed_curvature=False, should_do_curvature_step=lambda: True)
    model.train_step(opt, x, 1.0, eps=eps.to(dev))
    for i in range(3):
        want = kap_before[i] - 1e-4 * clip * gk[i]
        assert abs(float(model.components[i]._curvature.detach()) - want) < 2e-7 + 3e-4 * abs(1e-4 * clip * gk[i])
    # crossing the band: the same model object continues on the other manifold
    with torch.no_grad():
        model.components[0]._curvature.fill_(0.4)
        model.components[1]._curvature.fill_(0.0)
    assert [c.effective_kind() for c in model.components] == [4, 0, 0, 0]
    params = {k: v.detach().cpu().double().numpy() for k, v in model.state_dict().items()}
    ref2 = oracle.OracleVAE(sig, D, H, "bce", False).step(params, x.double().numpy(), eps.double().numpy(), beta=1.0,
                                                          backward=False)
    model.use_cuda_graph = True
    for _ in range(2):   # second call replays the CUDA graph captured by the first
        rep, z, logits = model.forward(x, eps=eps.to(dev))
        assert normwise(z.cpu().numpy(), ref2["z"]) < TOL
        assert normwise(torch.stack([r.kl for r in rep], -1).cpu().numpy(), ref2["kl"]) < TOL
    ll, mi, cov = model.log_likelihood(x.to(dev), n=4)
    assert torch.isfinite(ll).all() and torch.isfinite(mi).all()


@pytest.mark.parametrize("graph", [False, True])
def test_fused_optimizer_with_unaligned_input_width(dev, graph):
    """BDP-shaped model (in_dim = 50: rows of fc_e0.weight are not a multiple of 4 floats): the fused optimizer refreshes
    that matrix's operand planes with a separate launch; after every step the planes equal the updated weights."""
    from mvae_b200 import components, data, vae
    torch.manual_seed(1)
    B, D, H = 300, 50, 64
    model = vae.FusedFeedForwardVAE(H, components.parse_components("h2,p2", False), data.GenericDataset(B, D, "nll"),
                                    False, device=dev)
    model.use_cuda_graph = graph
    opt = vae.FusedCurvatureOptimizer(model, 1e-3, fixed_curvature=False, should_do_curvature_step=lambda: True)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, D, generator=g)
    w0 = model.fc_e0.weight.detach().clone()
    elbos = []
    for _ in range(4):
        bs, _ = model.train_step(opt, x, 1.0)
        elbos.append(bs.elbo)
        assert np.isfinite(bs.elbo)
        assert normwise(model.We0p.to_float().cpu().numpy(), model.fc_e0.weight.detach().cpu().numpy()) < 1e-6
        assert normwise(model.Wlp.to_float().cpu().numpy(), model.fc_logits.weight.detach().cpu().numpy()) < 1e-4
    assert not torch.equal(w0, model.fc_e0.weight.detach())
    assert elbos[-1] > elbos[0]
