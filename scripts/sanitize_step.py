"""Target of `compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_step.py [dp]`: a few train
steps of the path bench.py times (fused optimizer, the step replayed from ONE CUDA graph with its side branches) at
the BASELINE cfg2 and cfg3 model shapes with a reduced batch (the sanitizer slows kernels ~100x), plus — under
torchrun with `dp` — the peer-memory data-parallel step (mvae_dp_step, early fc_logits exchange on the third stream).
Logs are kept under profiles/ (race detection)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvae_b200 import components, data, parallel, vae  # noqa: E402

dp = len(sys.argv) > 1 and sys.argv[1] == "dp"
rank, world, local = parallel.init_from_env("nccl") if dp else (0, 1, 0)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
for sig, B in (("h2,s2,e2", 384), ("h6,h6,s6,s6,e6", 256)):
    torch.manual_seed(0)
    model = vae.FusedFeedForwardVAE(400, components.parse_components(sig, False),
                                    data.GenericDataset(B, 784, "bce", binary_inputs=True), False, device=dev)
    model.autotune_gemm = False
    model.use_cuda_graph = True
    opt = vae.FusedCurvatureOptimizer(model, 1e-3, fixed_curvature=False, should_do_curvature_step=lambda: True)
    if dp:
        assert parallel.attach_p2p(model, opt)
        parallel.broadcast_parameters(model)
    g = torch.Generator().manual_seed(rank)
    x = (torch.rand(B, 784, generator=g) < 0.13).float().to(dev)
    px = torch.randint(0, 256, (B, 784), generator=g, dtype=torch.int32).to(torch.uint8)
    for i in range(3):
        s, _ = model.train_step(opt, x, 1.0)
    out = model.train_epoch(opt, [px.pin_memory()] * 3, 1.0)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"{sig}: elbo {s.elbo:.3f} -> {out[-1].elbo:.3f} ({len(model._graphs)} graphs)", flush=True)
if dp:
    assert parallel.dp_error_word(opt) == 0
    torch.distributed.destroy_process_group()
print("sanitize_step done", flush=True)
