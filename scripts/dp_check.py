"""torchrun --nproc-per-node N scripts/dp_check.py — the peer-memory data-parallel step (mvae_dp_adam_step) against the
NCCL all-reduce path on the same shards: parameters, radii and statistics after a few steps must agree."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvae_b200 import components, data, parallel, vae  # noqa: E402

rank, world, local = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
sig, B, D, H = "h2,s2,e2", 1024, 784, 400


def build(use_graph):
    torch.manual_seed(0)
    m = vae.FusedFeedForwardVAE(H, components.parse_components(sig, False),
                                data.GenericDataset(B, D, "bce", binary_inputs=True), False, device=dev)
    m.use_cuda_graph = use_graph
    m.autotune_gemm = False  # identical tiles in every mode: local gradients are then bit-identical across modes
    o = vae.FusedCurvatureOptimizer(m, 1e-3, fixed_curvature=False, should_do_curvature_step=lambda: True)
    return m, o


g = torch.Generator().manual_seed(100 + rank)
xs = [(torch.rand(B, D, generator=g) < 0.13).float().to(dev) for _ in range(4)]
eps = [torch.randn(B, 6, generator=g).to(dev) for _ in range(4)]

results = {}
for mode in ("nccl", "p2p", "p2p_graph"):
    model, opt = build(mode == "p2p_graph")
    if mode == "nccl":
        parallel.attach(model)
    else:
        assert parallel.attach_p2p(model, opt), "peer mapping failed"
    parallel.broadcast_parameters(model)
    stats = None
    for i in range(4):
        stats, _ = model.train_step(opt, xs[i], 1.0, eps=eps[i])
    torch.cuda.synchronize()
    assert parallel.dp_error_word(opt) == 0 if mode != "nccl" else True
    results[mode] = (model._flat.clone(), model._rflat.clone(), stats)
    dist.barrier()

ref = results["nccl"]
for mode in ("p2p", "p2p_graph"):
    got = results[mode]
    dp = (got[0] - ref[0]).abs().max().item() / ref[0].abs().max().item()
    dr = (got[1] - ref[1]).abs().max().item()
    de = abs(got[2].elbo - ref[2].elbo) / abs(ref[2].elbo)
    # replicas must be bit-identical across ranks
    gathered = [torch.empty_like(got[0]) for _ in range(world)]
    dist.all_gather(gathered, got[0])
    same = all(torch.equal(gathered[0], t) for t in gathered)
    if rank == 0:
        print(f"{mode}: params rel diff vs nccl {dp:.2e}, radii diff {dr:.2e}, elbo rel diff {de:.2e}, "
              f"replicas identical: {same}, elbo {got[2].elbo:.3f}", flush=True)
    assert dp < 1e-5 and dr < 1e-6 and de < 1e-6 and same, mode
if rank == 0:
    print("dp_check ok", flush=True)
dist.destroy_process_group()
