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


def _worker_replicas(rank, world, port, out_path):
    """Host logic of the data-parallel self-checks on gloo: replicas_identical (bit-wise MIN/MAX all-reduce of the
    parameter words), broadcast_parameters, rank-offset noise / binarisation seeds, and the gather of the moment slices
    the optimizer keeps per rank (FusedCurvatureOptimizer.state_dict under the peer-memory path)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from mvae_b200 import components, data, parallel, vae
    parallel.init_from_env("gloo")
    torch.manual_seed(rank)   # different initialisation per rank on purpose
    model = vae.FusedFeedForwardVAE(16, components.parse_components("h2,s2,e2", False), data.GenericDataset(4, 20, "bce"),
                                    False, device="cpu")
    res = {"differ_before": parallel.replicas_identical(model)}
    seeds0 = (model.noise_seed, model.binarize_seed)
    parallel.attach(model)
    parallel.broadcast_parameters(model)
    res["identical_after_broadcast"] = parallel.replicas_identical(model)
    res["planes_stale"] = model._planes_stale
    res["seed_moved"] = (model.noise_seed != seeds0[0], model.binarize_seed != seeds0[1])
    seeds = [None] * world
    dist.all_gather_object(seeds, (model.noise_seed, model.binarize_seed))
    res["seeds_distinct"] = len(set(seeds)) == world
    with torch.no_grad():
        if rank == 1:
            model.fc_d0.bias[3] += 1e-7   # one parameter word on one rank
    res["identical_after_poke"] = parallel.replicas_identical(model)
    # sharded moments: every rank holds only its slice of each exchanged range; state_dict() gathers them
    opt = vae.FusedCurvatureOptimizer(model, 1e-3)

    class _Comm:
        pass

    opt._dp = _Comm()
    opt._dp.world, opt._dp.rank = world, rank
    full = torch.arange(model._n_net, dtype=torch.float32) + 1.0
    opt.exp_avg.zero_()
    opt.exp_avg_sq.zero_()
    for begin, end in opt._dp_ranges():
        n4 = (end - begin) // 4
        per = (n4 + world - 1) // world
        lo, hi = begin + 4 * min(n4, rank * per), begin + 4 * min(n4, (rank + 1) * per)
        opt.exp_avg[lo:hi] = full[lo:hi]
        opt.exp_avg_sq[lo:hi] = 2 * full[lo:hi]
    sd = opt.state_dict()
    res["moments_complete"] = bool(torch.equal(sd["exp_avg"], full) and torch.equal(sd["exp_avg_sq"], 2 * full))
    res["ranges"] = opt._dp_ranges()
    if rank == 0:
        import json
        with open(out_path, "w") as fh:
            json.dump(res, fh)
    dist.barrier()
    dist.destroy_process_group()


def test_replica_checks_and_sharded_optimizer_state_world_size_2(tmp_path):
    import json
    out = str(tmp_path / "res.json")
    mp.spawn(_worker_replicas, args=(2, _free_port(), out), nprocs=2, join=True)
    res = json.load(open(out))
    assert res["differ_before"] is False and res["identical_after_broadcast"] is True
    assert res["identical_after_poke"] is False
    assert res["planes_stale"] is True and res["seed_moved"] == [True, True] and res["seeds_distinct"] is True
    assert res["moments_complete"] is True
    (b0, e0), (b1, e1) = res["ranges"]   # fc_logits first (exchanged early), then the rest: a partition of the buffer
    assert b1 == 0 and e1 == b0 and e0 > b0
