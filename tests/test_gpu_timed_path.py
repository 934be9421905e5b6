"""GPU parity tests of THE PATH bench.py TIMES: FusedCurvatureOptimizer + the whole step replayed from ONE CUDA graph
(side-stream branches, autotuned GEMM tiles) + the pipelined train_epoch (both input slots, float and uint8 batches).

k steps with injected noise are followed by the float64 oracle running the reference's update (oracle.OracleTrainer:
OracleVAE.step + Adam + radii SGD, vae.py:149-166, train.py:327-360): per-step ELBO / BCE / KL and every parameter
after k steps.  A missing graph edge, a stale weight plane or a wrong input slot shows up as an O(1) error here."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import normwise

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

K = 5
BETA = 0.8
# (signature, batch, in_dim, h_dim, initial radius).  Wide spheres start at R = 10 like the reference's own schedule
# (Trainer._train_epoch sets R = 11 - epoch during the first ten epochs, train.py:189-194): at R = 1 a few of 8192
# samples of an s6 component land next to the log-det singularity |v| = pi R (spherical.py:58-67), where the radius
# gradient is the sum of a few huge terms of either sign — ill conditioned in ANY float32 arithmetic, and Adam / SGD
# then carry the difference into every later step.  scripts/diag_kstep.py shows the effect.
CONFIGS = [("h2,s2,e2", 4096, 784, 400, 1.0),           # BASELINE cfg2
           ("h6,h6,s6,s6,e6", 8192, 784, 400, 10.0)]    # BASELINE cfg3 (per-GPU batch)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mvae_b200 import _lib
    _lib.lib()
    return torch.device("cuda:0")


def _build(sig, B, D, H, dev, graph=True, seed=0, radius=1.0):
    from mvae_b200 import components, data, vae
    torch.manual_seed(seed)
    model = vae.FusedFeedForwardVAE(H, components.parse_components(sig, False),
                                    data.GenericDataset(B, D, "bce", binary_inputs=True), False, device=dev)
    with torch.no_grad():
        for rp in model._radius_params:
            if rp is not None:
                rp.fill_(radius)
    model.use_cuda_graph = graph
    opt = vae.FusedCurvatureOptimizer(model, 1e-3, fixed_curvature=False, should_do_curvature_step=lambda: True)
    return model, opt


def _params64(model):
    return {k: v.detach().cpu().double().numpy() for k, v in model.state_dict().items()}


def _movement_errors(model, want, start):
    """Per tensor: ||got - want||_F / ||want - start||_F (error relative to how far the oracle moved the tensor)."""
    errs = {}
    for k, v in model.state_dict().items():
        den = float(np.linalg.norm(want[k] - start[k]))
        if den > 0.0:
            errs[k] = float(np.linalg.norm(v.detach().cpu().double().numpy() - want[k])) / den
    return errs


def _batches(B, D, n_eps, seed):
    g = torch.Generator().manual_seed(seed)
    xs = [(torch.rand(B, D, generator=g) < 0.1307).float() for _ in range(K)]
    eps = [torch.randn(B, n_eps, generator=g) for _ in range(K)]
    return xs, eps


def _relu_decisions(model, ovae, params, x, eps):
    """The device's relu decisions of the step that just ran, after checking that every unit that differs from the
    float64 oracle sits at the kink (tests/test_gpu_model.py::test_train_step_vs_oracle explains why)."""
    ws = model._last_ws
    fwd = ovae.step(params, x, eps, beta=BETA, backward=False)
    dec = {}
    h_dev = ws.h32 if ws.fused else ws.hp.to_float()
    for name, act, pre in (("h", h_dev, fwd["h_pre"]), ("dd", ws.ddp.to_float(), fwd["dd_pre"])):
        on = act.cpu().numpy() > 0
        diff = on != (pre > 0)
        # how many units sit within float32 rounding of the kink depends on the batch (dense inputs: more); what
        # matters is that EVERY differing unit does
        assert diff.sum() <= max(8, 1e-4 * diff.size), (name, int(diff.sum()))
        assert np.all(np.abs(pre[diff]) <= 1e-5 * np.abs(pre).max()), (name, np.abs(pre[diff]).max())
        dec[name] = on
    return dec


@pytest.mark.parametrize("sig,B,D,H,R0", CONFIGS)
def test_graphed_fused_steps_and_pipelined_epoch_vs_oracle(dev, oracle, sig, B, D, H, R0):
    from mvae_b200 import ops
    ovae = oracle.OracleVAE(sig, D, H, "bce", False)
    # ---- (1) train_step: fused optimizer + ONE CUDA graph per step, autotuned tiles ----
    model, opt = _build(sig, B, D, H, dev, radius=R0)
    assert model.autotune_gemm
    p0 = _params64(model)
    trainer = oracle.OracleTrainer(ovae, p0)
    xs, eps = _batches(B, D, model.desc.ld_eps, seed=1)
    stats_a = []
    for i in range(K):
        x64, e64 = xs[i].double().numpy(), eps[i].double().numpy()
        bs, _ = model.train_step(opt, xs[i].to(dev), BETA, eps=eps[i].to(dev))
        dec = _relu_decisions(model, ovae, trainer.params, x64, e64)
        ref = trainer.step(x64, e64, BETA, relu_decisions=dec)
        assert abs(bs.elbo - ref["elbo"]) < 2e-5 * abs(ref["elbo"]), (i, bs.elbo, ref["elbo"])
        assert abs(bs.bce - ref["bce_sum"]) < 2e-5 * abs(ref["bce_sum"]), (i, bs.bce, ref["bce_sum"])
        assert abs(bs.kl - ref["kl_sum"]) < 2e-5 * abs(ref["kl_sum"]) + 1e-3, (i, bs.kl, ref["kl_sum"])
        np.testing.assert_allclose(bs.component_kl, ref["kl_comp"], rtol=5e-5, atol=1e-2)
        stats_a.append(bs)
    assert len(model._graphs) == 1 and opt.step_count == K and int(opt.step_dev.item()) == K
    errs = _movement_errors(model, trainer.params, p0)
    worst = max(errs, key=errs.get)
    print(f"\n[{sig}] graphed train_step vs oracle after {K} steps: worst movement error {errs[worst]:.2e} ({worst})")
    # Adam divides by sqrt(v): the update of an entry whose gradient is tiny is a ratio of two tiny numbers, so the
    # bar is on the Frobenius norm of each tensor's MOVEMENT (a race or a stale operand gives O(1); measured: ~5e-6)
    assert errs[worst] < 1e-3, (worst, errs[worst])
    for k, v in model.state_dict().items():   # and on the parameters themselves
        assert normwise(v.detach().cpu().numpy(), trainer.params[k]) < 1e-4, k
    # the operand planes the next step's GEMMs will read are those of the updated weights
    assert normwise(model.We0p.to_float().cpu().numpy(), model.fc_e0.weight.detach().cpu().numpy()) < 1e-6
    assert normwise(model.Wlp.to_float().cpu().numpy(), model.fc_logits.weight.detach().cpu().numpy()) < 1e-4
    ref_params = {k: v.detach().clone() for k, v in model.state_dict().items()}

    # ---- (2) the same batches through train_epoch (H2D of batch i+1 under step i, both input slots) ----
    model_b, opt_b = _build(sig, B, D, H, dev, radius=R0)
    out = model_b.train_epoch(opt_b, [x.pin_memory() for x in xs], BETA, eps_batches=[e.to(dev) for e in eps])
    assert len(out) == K
    assert len(model_b._graphs) == 2   # one graph per input slot
    for i in range(K):   # same kernels; the autotuner may pick other tiles (summation order) and atomics reorder
        assert abs(out[i].elbo - stats_a[i].elbo) < 1e-5 * abs(stats_a[i].elbo), (i, out[i].elbo, stats_a[i].elbo)
        assert abs(out[i].kl - stats_a[i].kl) < 1e-5 * abs(stats_a[i].kl) + 1e-3
    errs_b = _movement_errors(model_b, {k: v.cpu().double().numpy() for k, v in ref_params.items()}, p0)
    worst = max(errs_b, key=errs_b.get)
    print(f"[{sig}] train_epoch vs train_step after {K} steps: worst movement difference {errs_b[worst]:.2e} ({worst})")
    assert errs_b[worst] < 1e-3, (worst, errs_b[worst])
    for k, v in model_b.state_dict().items():
        assert normwise(v.detach().cpu().numpy(), ref_params[k].cpu().numpy()) < 1e-4, k

    # ---- (3) uint8 batches through train_epoch: binarised on the device inside the graph (Philox, device counter) ----
    g = torch.Generator().manual_seed(7)
    pxs = []
    for _ in range(K):   # MNIST-like pixels: ~26 % inked with a uniform intensity -> binarised density ~0.13
        ink = torch.rand(B, D, generator=g) < 0.26
        pxs.append((torch.randint(1, 256, (B, D), generator=g, dtype=torch.int32) * ink).to(torch.uint8))
    model_c, opt_c = _build(sig, B, D, H, dev, radius=R0)
    model_c.binarize_seed = 12345
    out_c = model_c.train_epoch(opt_c, [p.pin_memory() for p in pxs], BETA, eps_batches=[e.to(dev) for e in eps])
    # the batches the kernels saw: same Philox key and step counter, drawn standalone
    ctr = torch.zeros(1, dtype=torch.int64, device=dev)
    x_bin = []
    for i in range(K):
        ctr.fill_(i)
        x_bin.append(ops.binarize(pxs[i].to(dev), seed=12345, offset_dev=ctr).cpu())
    assert int(model_c._bin_ctr.item()) == K
    trainer_c = oracle.OracleTrainer(ovae, p0)
    model_d, opt_d = _build(sig, B, D, H, dev, radius=R0)     # twin fed the float batches one blocking step at a time
    for i in range(K):
        x64, e64 = x_bin[i].double().numpy(), eps[i].double().numpy()
        bs, _ = model_d.train_step(opt_d, x_bin[i].to(dev), BETA, eps=eps[i].to(dev))
        dec = _relu_decisions(model_d, ovae, trainer_c.params, x64, e64)
        ref = trainer_c.step(x64, e64, BETA, relu_decisions=dec)
        assert abs(bs.elbo - ref["elbo"]) < 2e-5 * abs(ref["elbo"])
        assert abs(out_c[i].elbo - ref["elbo"]) < 2e-5 * abs(ref["elbo"]), (i, out_c[i].elbo, ref["elbo"])
        assert abs(out_c[i].bce - ref["bce_sum"]) < 2e-5 * abs(ref["bce_sum"])
    errs = _movement_errors(model_c, trainer_c.params, p0)
    worst = max(errs, key=errs.get)
    print(f"[{sig}] uint8 train_epoch vs oracle after {K} steps: worst movement error {errs[worst]:.2e} ({worst})")
    assert errs[worst] < 1e-3, (worst, errs[worst])


def test_graph_replay_after_parameters_changed_outside(dev):
    """load_state_dict / a changed learning rate between two replays: the graph's GEMMs read weight PLANES, which must
    be rebuilt before the next replay, and values baked into launch parameters must not be replayed stale.  The eager
    path (same kernels, launched one by one) is the reference here."""
    sig, B, D, H = "h2,s2,e2", 512, 784, 400
    xs, eps = _batches(B, D, 6, seed=3)
    models = []
    for graph in (True, False):
        model, opt = _build(sig, B, D, H, dev, graph=graph)
        model.autotune_gemm = False   # identical tiles in both modes
        sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        out = []
        for i in range(2):
            out.append(model.train_step(opt, xs[i].to(dev), BETA, eps=eps[i].to(dev))[0].elbo)
        model.load_state_dict(sd0)                      # back to the start, optimizer moments kept
        out.append(model.train_step(opt, xs[2].to(dev), BETA, eps=eps[2].to(dev))[0].elbo)
        opt.lr = 5e-3                                   # baked into the optimizer launch
        out.append(model.train_step(opt, xs[3].to(dev), BETA, eps=eps[3].to(dev))[0].elbo)
        with torch.no_grad():
            model.fc_logits.weight.mul_(0.5)
        model.mark_parameters_changed()
        out.append(model.train_step(opt, xs[4].to(dev), BETA, eps=eps[4].to(dev))[0].elbo)
        models.append((model, out))
    (mg, og), (me, oe) = models
    for a, b in zip(og, oe):
        assert abs(a - b) < 2e-6 * abs(b), (og, oe)
    for (k, a), (_, b) in zip(mg.state_dict().items(), me.state_dict().items()):
        assert normwise(a.cpu().numpy(), b.cpu().numpy()) < 2e-5, k


def test_device_resident_batches_are_read_in_place(dev):
    """adopt_device_inputs (what bench.py's `value` uses): a float32 batch that lives on the device is read in place —
    same results as the copying path, one graph per batch tensor, and the kernels really see the tensor's CURRENT
    contents."""
    sig, B, D, H = "h2,s2,e2", 1024, 784, 400
    xs, eps = _batches(B, D, 6, seed=13)
    xs_dev = [x.to(dev) for x in xs[:3]]
    outs = []
    for adopt in (False, True):
        model, opt = _build(sig, B, D, H, dev)
        model.autotune_gemm = False
        model.adopt_device_inputs = adopt
        o = [model.train_step(opt, xs_dev[i % 3], BETA, eps=eps[i].to(dev))[0].elbo for i in range(5)]
        outs.append((model, o))
    (m0, o0), (m1, o1) = outs
    assert len(m0._graphs) == 1 and len(m1._graphs) == 3
    for a, b in zip(o0, o1):
        assert abs(a - b) < 2e-6 * abs(a), (o0, o1)
    for (k, a), (_, b) in zip(m0.state_dict().items(), m1.state_dict().items()):
        assert normwise(b.cpu().numpy(), a.cpu().numpy()) < 2e-5, k
    before = m1.train_step(opt, xs_dev[0], BETA, eps=eps[0].to(dev))[0].bce
    xs_dev[0].copy_(1.0 - xs_dev[0])   # same tensor, new contents: the replayed graph must read them
    after = m1.train_step(opt, xs_dev[0], BETA, eps=eps[0].to(dev))[0].bce
    assert abs(after - before) > 0.5 * abs(before)


def test_radius_rebind_reaches_the_kernels(dev, oracle):
    """Trainer._train_epoch REBINDS `c._pradius.data = ones_like(...) * (11 - epoch)` in its first ten epochs
    (train.py:189-194).  The kernels read the flat radius vector: the rebind must land there (and the parameter be
    re-homed) before the next step, inside a CUDA-graph replay too."""
    sig, B, D, H = "h2,s2,p2,e2", 512, 784, 64
    model, opt = _build(sig, B, D, H, dev)
    ovae = oracle.OracleVAE(sig, D, H, "bce", False)
    xs, eps = _batches(B, D, model.desc.ld_eps, seed=5)
    model.train_step(opt, xs[0].to(dev), BETA, eps=eps[0].to(dev))
    for epoch in (0, 3):
        for c in model.components:   # the loop of train.py:189-194, by attribute instead of isinstance
            if hasattr(c, "_pradius"):
                c._pradius.data = torch.ones_like(c._pradius.data) * (11 - epoch)
            elif hasattr(c, "_nradius"):
                c._nradius.data = torch.ones_like(c._nradius.data) * (11 - epoch)
        params = _params64(model)
        for k in params:
            if "radius" in k:
                assert float(params[k]) == 11 - epoch
        bs, _ = model.train_step(opt, xs[1].to(dev), BETA, eps=eps[1].to(dev))
        ref = ovae.step(params, xs[1].double().numpy(), eps[1].double().numpy(), beta=BETA, backward=False)
        assert abs(bs.kl - ref["kl_sum"]) < 2e-5 * abs(ref["kl_sum"]) + 1e-3, (epoch, bs.kl, ref["kl_sum"])
        np.testing.assert_allclose(bs.component_kl, ref["kl_comp"], rtol=5e-5, atol=1e-2)
        for i, c in enumerate(model.components):   # re-homed: the optimizer's SGD step moved the SAME storage
            for nm in ("_pradius", "_nradius"):
                if hasattr(c, nm):
                    assert getattr(c, nm).data_ptr() == model._rflat[i].data_ptr()
                    assert abs(float(getattr(c, nm).detach()) - (11 - epoch)) < 1.0
                    assert float(getattr(c, nm).detach()) != 11 - epoch   # learnable curvature: SGD moved it


def test_train_step_outputs_are_the_references(dev):
    """vae.py:166 returns (reparametrized, concat_z, x_mb_); Trainer._train_epoch iterates `reparametrized` under
    --train_statistics (train.py:198-206).  The fused step builds them on first use."""
    sig, B, D, H = "h2,s2,e2", 256, 784, 64
    model, opt = _build(sig, B, D, H, dev, graph=False)
    opt.curv_condition = lambda: False   # radii fixed: the lazy recomputation sees the step's own radii
    xs, eps = _batches(B, D, model.desc.ld_eps, seed=9)
    x, e = xs[0].to(dev), eps[0].to(dev)
    rep_f, z_f, logits_f = model.forward(x, eps=e)
    mu_f = torch.cat([r.q_z.loc for r in rep_f], -1).clone()
    sg_f = torch.cat([r.q_z.scale for r in rep_f], -1).clone()
    z_f = z_f.clone()
    for stats_mode in (False, True):
        m2, o2 = _build(sig, B, D, H, dev, graph=stats_mode)
        m2.load_state_dict(model.state_dict())
        m2.train_statistics = stats_mode
        o2.lr = 0.0
        o2.curv_condition = lambda: False
        stats, (rep, concat_z, x_mb_) = m2.train_step(o2, x, 1.0, eps=e)
        assert len(rep) == len(m2.components)
        for comp, r in zip(m2.components, rep):          # train.py:201-206
            s = comp.summaries(0, r.q_z, prefix="train/batch")
            assert all(torch.isfinite(v).all() for v in s.values())
        assert normwise(torch.cat([r.q_z.loc for r in rep], -1).cpu().numpy(), mu_f.cpu().numpy()) < 1e-5
        assert normwise(torch.cat([r.q_z.scale for r in rep], -1).cpu().numpy(), sg_f.cpu().numpy()) < 1e-5
        assert normwise(concat_z.cpu().numpy(), z_f.cpu().numpy()) < 1e-5
        assert tuple(x_mb_.shape) == (B, D)
        assert normwise(torch.sigmoid(x_mb_).cpu().numpy(), torch.sigmoid(logits_f).cpu().numpy()) < 1e-4
        assert rep[0].data is not None and rep[-1].data is None   # (u, v) parts of the wrapped normal; none for 'e'


def test_uint8_evaluation_targets(dev):
    """forward(x_u8) + compute_batch_stats(x_u8, ...) — the reference's evaluation call pattern (train.py:235-239,
    eval.py:93-97) with raw pixels: the targets of the reconstruction loss are the BINARISED batch, not 0..255."""
    sig, B, D, H = "h2,s2,e2", 128, 784, 64
    model, _ = _build(sig, B, D, H, dev, graph=False)
    g = torch.Generator().manual_seed(2)
    px = torch.randint(0, 256, (B, D), generator=g, dtype=torch.int32).to(torch.uint8)
    e = torch.randn(B, model.desc.ld_eps, generator=g).to(dev)
    rep, z, logits = model.forward(px, eps=e)
    st8 = model.compute_batch_stats(px, logits, rep, beta=1.0).convert_to_float()
    xf = (px.float().div(255) > 0.5).float()
    rep2, z2, logits2 = model.forward(xf, eps=e)
    stf = model.compute_batch_stats(xf, logits2, rep2, beta=1.0).convert_to_float()
    assert st8.bce >= 0 and np.isfinite(st8.elbo)
    assert abs(st8.bce - stf.bce) < 1e-6 * abs(stf.bce) and abs(st8.elbo - stf.elbo) < 1e-6 * abs(stf.elbo)
    h8, hf = model.encode(px.to(dev)).clone(), model.encode(xf.to(dev)).clone()
    assert torch.equal(h8, hf)


def test_data_parallel_steps_vs_oracle_on_the_global_batch():
    """2 GPUs: NCCL all-reduce, the peer-memory kernel, its overlapped form and the latter inside the CUDA graph, each
    against the float64 oracle on the GLOBAL batch (scripts/dp_check.py).  Skipped on a single-GPU box; the log of a
    2-GPU run is kept under profiles/."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "scripts", "dp_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "dp_check ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_step_prologue_noise_and_zero_fill(dev):
    """mvae_step_prologue: the N(0, I) draw behind Normal.rsample (wrapped_normal.py:72) as Philox4x32 + Box-Muller —
    deterministic per (seed, step counter), fresh per step, standard normal to 5 sigma on 4 M draws — and the zero
    fill that stands in for optimizer.zero_grad() (vae.py:151), for aligned and unaligned spans."""
    from mvae_b200 import ops
    n = 1 << 22
    ctr = torch.zeros(1, dtype=torch.int64, device=dev)
    a, b, c = (torch.empty(n, device=dev) for _ in range(3))
    ops.step_prologue(a, 7, ctr)
    ops.step_prologue(b, 7, ctr)
    assert torch.equal(a, b)
    ops.counter_add(ctr)
    assert int(ctr.item()) == 1
    ops.step_prologue(c, 7, ctr)
    assert not torch.equal(a, c)
    ops.step_prologue(b, 8, ctr)
    assert not torch.equal(b, c)
    x = a.double()
    se = 1.0 / np.sqrt(n)
    assert abs(float(x.mean())) < 5 * se
    assert abs(float(x.var()) - 1.0) < 5 * np.sqrt(2.0) * se
    assert abs(float((x ** 4).mean()) - 3.0) < 5 * np.sqrt(96.0) * se
    assert abs(float((a * c).double().mean())) < 5 * se            # consecutive steps are uncorrelated
    assert abs(float((a[:-1] * a[1:]).double().mean())) < 5 * se   # neighbours within a step too
    assert float(a.abs().max()) < 7.0 and torch.isfinite(a).all()
    # ragged length + unaligned zero spans
    e = torch.full((1003,), 9.0, device=dev)
    buf = torch.full((4099,), 5.0, device=dev)
    z1, z2 = buf[1:1030], buf[2048:4099]
    ops.step_prologue(e[:1001], 3, ctr, [z1, z2])
    assert float(e[1001]) == 9.0 and float(e[1002]) == 9.0 and float(e[:1001].abs().max()) < 6.0
    assert float(buf[0]) == 5.0 and float(buf[1030]) == 5.0 and float(buf[2047]) == 5.0
    assert float(z1.abs().sum()) == 0.0 and float(z2.abs().sum()) == 0.0
    ops.step_prologue(None, 0, None, [buf])
    assert float(buf.abs().sum()) == 0.0


def test_train_step_draws_its_own_noise(dev):
    """Without supplied eps the step draws N(0, I) itself (inside the graph), fresh every step, also through
    train_epoch and forward()."""
    model, opt = _build("h2,s2,e2", 2048, 784, 400, dev)
    xs, _ = _batches(2048, 784, 6, seed=11)
    seen = []
    for i in range(3):
        model.train_step(opt, xs[i].to(dev), 1.0)
        seen.append(model._last_ws.eps.clone())
    model.train_epoch(opt, [x.pin_memory() for x in xs[:2]], 1.0)
    seen.append(model._last_ws.eps.clone())
    model.forward(xs[0].to(dev))
    seen.append(model._last_ws.eps.clone())
    model.forward(xs[0].to(dev))
    seen.append(model._last_ws.eps.clone())
    for i in range(len(seen)):
        e = seen[i].double()
        assert abs(float(e.mean())) < 0.05 and abs(float(e.var()) - 1.0) < 0.05
        for j in range(i):
            assert not torch.equal(seen[i], seen[j]), (i, j)
