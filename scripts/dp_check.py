"""torchrun --nproc-per-node N scripts/dp_check.py — the data-parallel step against the float64 ORACLE ON THE GLOBAL
BATCH (the reference is single-device: N ranks with B/N rows each must reproduce its step at batch B).

For every exchange mode — NCCL all-reduce, the peer-memory kernel (mvae_dp_step), the peer-memory kernel with the
fc_logits range exchanged early on a side branch, and the latter inside the step's CUDA graph (what bench.py times) —
k steps with injected noise are compared step by step (ELBO, BCE, KL of the global batch) and at the end (every
parameter, radii) with oracle.OracleTrainer = OracleVAE.step + Adam + radii SGD fed the concatenated shards.  Also
checked: replicas are bit-identical, the sticky error word of mvae_dp_step is clear, the gathered optimizer state is
complete, and a model with 'u' components (gradient clip between the reduction and the update) takes the peer path.

Prints 'dp_check ok' on rank 0 (tests/test_gpu_timed_path.py runs this under torchrun when the box has >= 2 GPUs)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from mvae_b200 import components, data, parallel, vae  # noqa: E402

rank, world, local = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
K = 4
BETA = 0.9


def gather_rows(t):
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t.contiguous())
    return torch.cat(out, 0)


def run_case(sig, B, D, H, mode, radius=1.0):
    os.environ["MVAE_DP_OVERLAP"] = "1" if mode in ("p2p_overlap", "p2p_overlap_graph") else "0"
    torch.manual_seed(0)
    model = vae.FusedFeedForwardVAE(H, components.parse_components(sig, False),
                                    data.GenericDataset(B, D, "bce", binary_inputs=True), False, device=dev)
    model.use_cuda_graph = mode.endswith("graph")
    opt = vae.FusedCurvatureOptimizer(model, 1e-3, fixed_curvature=False, should_do_curvature_step=lambda: True)
    if mode == "nccl":
        parallel.attach(model)
    else:
        assert parallel.attach_p2p(model, opt), "peer mapping failed"
    with torch.no_grad():
        for c in model.components:   # wide spheres start at R = 10 like the reference's schedule (train.py:189-194):
            for nm in ("_nradius", "_pradius"):   # see tests/test_gpu_timed_path.py on conditioning at R = 1
                if hasattr(c, nm):
                    getattr(c, nm).fill_(radius)
    parallel.broadcast_parameters(model)
    if any(hasattr(c, "_curvature") for c in model.components):  # 'u': curvatures on both sides of zero
        with torch.no_grad():
            for c, kappa in zip([c for c in model.components if hasattr(c, "_curvature")], (-0.6, 0.8, 0.3)):
                c._curvature.fill_(kappa)
    p0 = {k: v.detach().cpu().double().numpy() for k, v in model.state_dict().items()}
    trainer = None
    if rank == 0:
        import oracle as orc
        trainer = orc.OracleTrainer(orc.OracleVAE(sig, D, H, "bce", False), p0)
    g = torch.Generator().manual_seed(100 + rank)
    worst = 0.0
    for i in range(K):
        x = (torch.rand(B, D, generator=g) < 0.13).float().to(dev)
        eps = torch.randn(B, model.desc.ld_eps, generator=g).to(dev)
        stats, _ = model.train_step(opt, x, BETA, eps=eps)
        ws = model._last_ws
        # relu decisions the devices took (relu'(0) is a convention; see tests/test_gpu_model.py), global batch order
        h_on = gather_rows(((ws.h32 if ws.fused else ws.hp.to_float()) > 0).to(torch.uint8))
        dd_on = gather_rows((ws.ddp.to_float() > 0).to(torch.uint8))
        gx, ge = gather_rows(x), gather_rows(eps)
        if rank == 0:
            dec = {"h": h_on.cpu().numpy().astype(bool), "dd": dd_on.cpu().numpy().astype(bool)}
            ref = trainer.step(gx.cpu().double().numpy(), ge.cpu().double().numpy(), BETA, relu_decisions=dec)
            for name, got, want in (("elbo", stats.elbo, ref["elbo"]), ("bce", stats.bce, ref["bce_sum"]),
                                    ("kl", stats.kl, ref["kl_sum"])):
                rel = abs(got - want) / max(abs(want), 1.0)
                worst = max(worst, rel)
                assert rel < 3e-5, (mode, sig, i, name, got, want)
    torch.cuda.synchronize()
    if mode != "nccl":
        assert parallel.dp_error_word(opt) == 0, "mvae_dp_step reported a time-out"
    assert parallel.replicas_identical(model), (mode, sig, "replicas differ")
    sd_opt = opt.state_dict()  # collective under the peer path: gathers the moment slices
    assert sd_opt["step"] == K
    assert float(sd_opt["exp_avg_sq"].abs().sum()) > 0
    msg = ""
    if rank == 0:
        errs = {}
        for k, v in model.state_dict().items():
            got, want, start = v.detach().cpu().double().numpy(), trainer.params[k], p0[k]
            den = float(np.linalg.norm(want - start))
            if den == 0.0:
                continue
            errs[k] = float(np.linalg.norm(got - want)) / den
        kmax = max(errs, key=errs.get)
        # Adam divides by sqrt(v): the update of an entry whose gradient is tiny is the ratio of two tiny numbers, so
        # the bar is on the Frobenius norm of each tensor's MOVEMENT (a missing edge / a wrong slice gives O(1))
        assert errs[kmax] < 5e-3, (mode, sig, kmax, errs[kmax])
        msg = f"{mode:18s} {sig:14s} stats rel {worst:.1e}  worst movement error {errs[kmax]:.1e} ({kmax})"
        print(msg, flush=True)
    dist.barrier()
    return model, opt


for mode in ("nccl", "p2p", "p2p_overlap", "p2p_overlap_graph"):
    run_case("h2,s2,e2", 2048, 784, 400, mode)
run_case("h6,h6,s6,s6,e6", 2048, 784, 400, "p2p_overlap_graph", radius=10.0)   # wide product: latent dense layers on the tensor cores
run_case("u2,u2,u2,e2", 2048, 784, 64, "p2p_overlap_graph")      # clip of the curvature gradients inside the kernel
if rank == 0:
    print("dp_check ok", flush=True)
dist.destroy_process_group()
