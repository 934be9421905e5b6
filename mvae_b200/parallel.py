"""Data-parallel training of the fused VAE: one process per GPU, batch sharded by rank, ONE SUM all-reduce per step
over the flat bucket [network grads | radius grads | ELBO statistics] (NCCL over NVLink / NVSwitch).

The reference has no distributed mode; its ELBO is a SUM over the batch (mt/mvae/stats.py:200-202),
so the cross-GPU reduction is a SUM, not DDP's mean: N ranks with B/N samples each reproduce the single-GPU step at
batch B up to summation order.  Parameters and optimizer state are replicated; every rank applies the same update.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> tuple:
    """Initialise torch.distributed from torchrun's environment.  Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def bind_to_gpu_numa_node(local_rank: int) -> bool:
    """Pin this process to the CPUs NVML reports as local to its GPU, so that the pinned host batches it allocates
    afterwards (first touch) sit on the socket the GPU's PCIe root hangs off: with one process per GPU and eight
    12.8 MB host->device copies per 0.25 ms step, cross-socket traffic is what saturates first."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False


def shard_bounds(batch: int, rank: int, world: int) -> tuple:
    """Contiguous batch split: rank r owns rows [r*B/N, (r+1)*B/N) — remainder to the first ranks."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(bucket: torch.Tensor, group=None) -> torch.Tensor:
    """The single collective of the data path."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    return bucket


def attach(model, group=None) -> None:
    """Make `model.train_step` all-reduce its gradient/statistics bucket before the optimizer step."""
    model._grad_hook = lambda bucket: allreduce_sum_(bucket, group)
    if dist.is_initialized():
        model.binarize_seed = int(model.binarize_seed) + 0x9E3779B1 * (dist.get_rank(group) + 1)
        model.noise_seed = (int(model.noise_seed) + 0x85EBCA6B * (dist.get_rank(group) + 1)) & (2**62 - 1)


def broadcast_parameters(model, src: int = 0, group=None) -> None:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(model._flat, src=src, group=group)
        dist.broadcast(model._rflat, src=src, group=group)
        model.mark_parameters_changed()


# ---------------------------------------------------------------------------------------------- peer-memory path
class _DeviceRegion:
    """A cudaMalloc'ed region exposed to torch through the CUDA array interface (zero-copy views)."""

    def __init__(self, ptr: int, nbytes: int, owner=None):
        self.ptr, self.nbytes, self.owner = ptr, nbytes, owner

    def view(self, offset_bytes: int, count: int, dtype, device):
        itemsize = torch.empty((), dtype=dtype).element_size()
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]

        class _Holder:
            pass

        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (self.ptr + offset_bytes, False),
                                      "version": 3}
        h._region = self  # keeps the allocation alive as long as the tensor lives
        assert offset_bytes + count * itemsize <= self.nbytes
        return torch.as_tensor(h, device=device)


def attach_p2p(model, optimizer, group=None) -> bool:
    """Data-parallel step over NVLink peer memory (mvae_dp_step): the model's parameter buffer and gradient
    bucket are re-homed into a cudaMalloc'ed region that every peer maps through CUDA IPC; the optimizer's step then
    does gradient reduce-scatter + Adam + parameter all-gather in ONE kernel instead of all-reduce -> Adam.
    Single node, world size <= 8.  Returns False (and leaves the model untouched) if the peers cannot be mapped."""
    import ctypes

    from . import _lib as L
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return False
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world > L.DP_MAX_RANKS:
        return False
    lib = L.lib()
    dev = model.device
    n_flat, n_bucket = model.flat_sizes()
    al = lambda x: (x + 255) // 256 * 256  # noqa: E731
    off_bucket = 0
    off_flat = off_bucket + al(4 * n_bucket)
    off_flags = off_flat + al(4 * n_flat)
    nbytes = off_flags + L.DP_FLAG_BYTES
    ptr = ctypes.c_void_p()
    ok = True
    try:
        L.check(lib.mvae_dp_alloc(nbytes, ctypes.byref(ptr)), "mvae_dp_alloc")
        handle = ctypes.create_string_buffer(L.DP_HANDLE_BYTES)
        L.check(lib.mvae_dp_ipc_export(ptr, handle), "mvae_dp_ipc_export")
        mine = bytes(handle.raw)
    except L.MvaeError:
        ok, mine = False, b""
    handles = [None] * world
    dist.all_gather_object(handles, mine if ok else None, group=group)
    if any(h is None for h in handles):
        return False
    bases = []
    for r in range(world):
        if r == rank:
            bases.append(ptr.value)
            continue
        pp = ctypes.c_void_p()
        rc = lib.mvae_dp_ipc_open(handles[r], ctypes.byref(pp))
        bases.append(pp.value if rc == 0 else None)
    all_ok = [None] * world
    dist.all_gather_object(all_ok, all(b is not None for b in bases), group=group)
    if not all(all_ok):
        return False
    region = _DeviceRegion(ptr.value, nbytes)
    flat = region.view(off_flat, n_flat, torch.float32, dev)
    bucket = region.view(off_bucket, n_bucket, torch.float32, dev)
    model._flatten(storage=(flat, bucket))
    comm = L.DpComm()
    comm.rank, comm.world = rank, world
    for r in range(world):
        comm.bucket[r] = bases[r] + off_bucket
        comm.flat[r] = bases[r] + off_flat
        comm.flags[r] = bases[r] + off_flags
    C = model.desc.C
    optimizer._dp = comm
    optimizer._dp_tail = torch.zeros(2 * C + 3 + 1, device=dev, dtype=torch.float32)  # + the error word (as float)
    optimizer._dp_sync = torch.zeros(L.DP_SYNC_WORDS, device=dev, dtype=torch.int32)
    # moments follow the re-homed parameters (fresh optimizer state: moments AND step counter), statistics are read
    # from the summed tail
    optimizer.exp_avg = torch.zeros_like(model._flat)
    optimizer.exp_avg_sq = torch.zeros_like(model._flat)
    optimizer.step_count = 0
    optimizer.step_dev.zero_()
    optimizer.step_dev_early.zero_()
    model._stats_report = optimizer._dp_tail[C:2 * C + 3]
    model._stats_wire = optimizer._dp_tail[C:]
    model._grad_hook = None
    model._dp_region = region
    # dynamic binarisation (uint8 batches): every rank draws its own uniforms
    model.binarize_seed = int(model.binarize_seed) + 0x9E3779B1 * (rank + 1)
    model.noise_seed = (int(model.noise_seed) + 0x85EBCA6B * (rank + 1)) & (2**62 - 1)  # rank-offset seeds for eps
    torch.cuda.synchronize(dev)
    dist.barrier(group=group)  # every peer has mapped every region before the first step
    return True


def rendezvous(optimizer) -> None:
    """Enqueue a kernel that returns once every rank has enqueued it (mvae_dp_rendezvous): re-aligns the ranks on the
    device without a host barrier.  No-op without the peer-memory path."""
    if getattr(optimizer, "_dp", None) is None:
        return
    import ctypes

    from . import _lib as L
    rc = L.lib().mvae_dp_rendezvous(ctypes.byref(optimizer._dp), optimizer._dp_sync.data_ptr(),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    L.check(rc, "mvae_dp_rendezvous")


def dp_error_word(optimizer) -> int:
    """0 unless a peer failed to arrive at a flag wait of mvae_dp_step within its time limit (sticky: every later step
    is a no-op).  train_step / train_epoch raise on it by themselves — it travels with the step's statistics."""
    return 0 if optimizer._dp_sync is None else int(optimizer._dp_sync[8].item())


def dp_phase_times(optimizer) -> dict:
    """Durations (microseconds) of the phases of the LAST mvae_dp_step launches on this rank, from the %globaltimer
    stamps the kernel leaves in its sync words: the launch at the end of the step (wait for the peers' gradients,
    reduce-scatter + Adam + all-gather, wait for the peers' slices, weight-plane refresh) and the early launch."""
    if optimizer._dp_sync is None:
        return {}
    w = [int(v) & 0xFFFFFFFF for v in optimizer._dp_sync.tolist()]
    d = lambda a, b: ((w[b] - w[a]) & 0xFFFFFFFF) / 1e3  # noqa: E731
    # per-CTA stamps: when the slowest CTA passed phase A / finished its slice, relative to CTA 0 passing phase A
    ctas = [(w[16 + 2 * b], w[17 + 2 * b]) for b in range(256) if w[17 + 2 * b] != 0]
    rel = lambda x: ((x - w[10]) & 0xFFFFFFFF) / 1e3 if ((x - w[10]) & 0xFFFFFFFF) < (1 << 31) else -(((w[10] - x) & 0xFFFFFFFF) / 1e3)  # noqa: E731
    extra = {}
    if ctas:
        extra = {"late_ctas": len(ctas), "late_last_cta_passes_A": max(rel(a) for a, _ in ctas),
                 "late_last_cta_done": max(rel(b) for _, b in ctas),
                 "late_median_cta_done": sorted(rel(b) for _, b in ctas)[len(ctas) // 2]}
    return {**extra, "late_wait_grads": d(9, 10), "late_reduce_adam_gather": d(10, 11), "late_cta0_loads_adam_stores": d(10, 14),
            "late_cta0_fence_sys": d(14, 15), "late_wait_slices": d(11, 12),
            "late_plane_refresh": d(12, 13), "late_total": d(9, 13), "early_total": d(2, 3),
            "early_start_to_late_start": d(2, 9)}


def replicas_identical(model, group=None) -> bool:
    """True when every rank holds bit-identical parameters and radii (collective call)."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return True
    mine = torch.cat([model._flat.detach().reshape(-1), model._rflat.detach().reshape(-1)]).view(torch.int32)
    lo, hi = mine.clone(), mine.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return bool(torch.equal(lo, hi))
