"""CPU tests (gloo, world size 2) of the data-parallel host logic in mvae_b200/parallel.py: contiguous batch sharding
and the single SUM all-reduce over [grads | stats].  The per-shard arithmetic is done by the CPU oracle here (no GPU):
the property under test is that SUM over shards of (gradients, ELBO statistics) equals the full-batch step — the
reason the collective is a SUM and not DDP's mean (mt/mvae/stats.py:200-202 sums over the batch)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import oracle as orc
    from mvae_b200 import parallel
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    sig, D, H = "h2,s2,e2", 20, 16
    rng = np.random.default_rng(0)  # identical on every rank: replicated parameters, one global batch
    ov = orc.OracleVAE(sig, D, H, "bce", False)
    params = {}
    for i, (n, l_n) in enumerate([(2, 2)] * 3):
        params[f"components.{i}.fc_mean.weight"] = rng.standard_normal((n, H)) * 0.3
        params[f"components.{i}.fc_mean.bias"] = rng.standard_normal(n) * 0.1
        params[f"components.{i}.fc_logvar.weight"] = rng.standard_normal((l_n, H)) * 0.3
        params[f"components.{i}.fc_logvar.bias"] = rng.standard_normal(l_n) * 0.1
    params["components.0._nradius"] = np.asarray(1.0)
    params["components.1._pradius"] = np.asarray(1.0)
    for nm, shp in (("fc_e0", (H, D)), ("fc_d0", (H, 8)), ("fc_logits", (D, H))):
        params[nm + ".weight"] = rng.standard_normal(shp) * 0.3
        params[nm + ".bias"] = rng.standard_normal(shp[0]) * 0.1
    x = (rng.random((B, D)) < 0.3).astype(np.float64)
    eps = rng.standard_normal((B, 6))
    lo, hi = parallel.shard_bounds(B, rank, world)
    out = ov.step(params, x[lo:hi], eps[lo:hi], beta=0.7)
    keys = sorted(out["grads"])
    flat = np.concatenate([np.asarray(out["grads"][k], dtype=np.float64).reshape(-1) for k in keys] +
                          [np.asarray([out["bce_sum"], out["kl_sum"], out["elbo"]]), out["kl_comp"]])
    bucket = torch.from_numpy(flat.copy())
    parallel.allreduce_sum_(bucket)
    if rank == 0:
        full = ov.step(params, x, eps, beta=0.7)
        ref = np.concatenate([np.asarray(full["grads"][k], dtype=np.float64).reshape(-1) for k in keys] +
                             [np.asarray([full["bce_sum"], full["kl_sum"], full["elbo"]]), full["kl_comp"]])
        np.save(out_path, np.stack([bucket.numpy(), ref]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [12, 7])
def test_sum_allreduce_over_shards_equals_full_batch(tmp_path, B):
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, _free_port(), B, out), nprocs=2, join=True)
    got, ref = np.load(out)
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-10)


def test_shard_bounds_cover_the_batch_exactly():
    sys.path.insert(0, ROOT)
    from mvae_b200 import parallel
    for B in (1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
