#!/usr/bin/env python
"""bench.py — ELBO training throughput of the mvae hot path on B200 (see BASELINE.json / DESIGN.md §6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4a|cfg4b|cfg1]

A step = one full ModelVAE.train_step (vae.py:149-166) on one synthetic MNIST-shaped batch: forward, ELBO, backward,
optimizer update (+ the single gradient all-reduce when N > 1).  N=1 workload = BASELINE.json configs[1]
(MNIST h2,s2,e2, learnable curvature, batch 4096); N>1 keeps 4096 samples per GPU (weak scaling).

  value   whole-job samples/s (steps/s x global batch), inputs resident in HBM, CUDA-event timed, L2 flushed between steps
  e2e     the same through the public API model.train_step(optimizer, x_host, beta): pinned-host -> device copy of the
          batch and device -> host read of the ELBO statistics inside the timed region
  roofline  the fused product-manifold kernel (the kernel BASELINE names): algorithmic bytes / CUDA-event time vs
            the measured HBM copy bandwidth; `roofline_gemm` does the same for the tcgen05 GEMM against measured bf16
  cpu_baseline / --impl reference   the oracle port of the reference's algorithm (oracle/, numpy BLAS + C/OpenMP) on the
          host cores; the reference itself is pure Python and is not present on the GPU box
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (signature, per-GPU batch, in_dim, h_dim, recon, fixed_curvature, description)
    "cfg1": ("e2", 128, 784, 400, "bce", True, "MNIST e2 fixed curvature batch 128"),
    "cfg2": ("h2,s2,e2", 4096, 784, 400, "bce", False, "MNIST h2,s2,e2 learnable curvature batch 4096"),
    "cfg3": ("h6,h6,s6,s6,e6", 8192, 784, 400, "bce", False, "MNIST h6,h6,s6,s6,e6 batch 8192 per GPU"),
    "cfg4a": ("h2", 16384, 50, 400, "nll", False, "BDP-shaped h2 (hyperboloid) batch 16384"),
    "cfg4b": ("p2", 16384, 50, 400, "nll", False, "BDP-shaped p2 (Poincare ball) batch 16384"),
    "cfg5": ("h2,s2,e2", 256, 3072, 8192, "bce", False,
             "CIFAR-shaped conv VAE (h_dim 8192) h2,s2,e2, batch 2048 over 8 GPUs = 256 per GPU"),
}
CONV_WORKLOADS = ("cfg5",)   # ConvolutionalVAE (conv_vae.py:28-79); the others are FeedForwardVAE
METRIC = "ELBO train throughput (steps/sec x global batch)"
UNIT = "samples/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


KERNEL_SOURCES = {  # files whose edits invalidate an ncu capture of the kernel
    "pm_forward_kernel": ["pm_kernels_impl.cuh", "pm_item.cuh", "pm_math.cuh", "pm_params.cuh"],
    "pm_backward_kernel": ["pm_kernels_impl.cuh", "pm_item.cuh", "pm_math.cuh", "pm_params.cuh", "pm_bwd.cu"],
    "gemm_tcgen05_kernel": ["gemm_sm100.cu"],
}


def source_hash(kernel):
    import hashlib
    h = hashlib.sha1()
    for f in KERNEL_SOURCES[kernel]:
        with open(os.path.join(ROOT, "mvae_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:12]


def ncu_traffic(kernel, signature, samples):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of `kernel` from the committed
    `ncu --set full` capture, profiles/ncu_traffic.json (written by scripts/summarize_profiles.py together with a hash
    of the kernel's sources).  None — with the reason — when there is no capture of this shape or the sources have
    changed since: a number from another build is not reported as this build's."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, "no capture committed"
    with open(path) as fh:
        table = json.load(fh)
    for row in table.get(kernel, []):
        if row.get("signature") == signature and int(row.get("samples", -1)) == int(samples):
            if row.get("source_hash") != source_hash(kernel):
                return None, f"capture {row.get('capture')} predates the current kernel sources"
            return float(row["dram_bytes"]), row.get("capture")
    return None, "no capture of this shape"


def synthetic_x(recon, B, D, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    if D == 3072:   # CIFAR-shaped: real-valued pixels in [0, 1] (ToTensor), BCE on real-valued targets
        return torch.rand(B, D, generator=g)
    if recon == "bce":
        return (torch.rand(B, D, generator=g) < 0.1307).float()
    return torch.randn(B, D, generator=g)


def synthetic_pixels(B, D, seed):
    """MNIST-shaped raw grayscale batch (uint8, what the dataset files hold): ~26 % of the pixels are inked with a
    uniform intensity, so that the dynamically binarised batch (x/255 > U(0,1), image_reconstruction.py:37-53) has
    MNIST's mean density 0.13 like synthetic_x."""
    import torch
    g = torch.Generator().manual_seed(seed)
    ink = torch.rand(B, D, generator=g) < 0.26
    val = torch.randint(1, 256, (B, D), generator=g, dtype=torch.int32)
    return (val * ink).to(torch.uint8)


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled DURING the timed region.

    NVML is polled from a thread every ~2 ms (a timed region of a few dozen 0.3 ms steps is far shorter than one
    `nvidia-smi -lms` period); `nvidia-smi` is the fallback when the NVML binding is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread = index, [], None, None
        self.stop_flag = threading.Event()
        self.nvml, self.handle, self.max_mhz = None, None, None

    def _nvml_open(self):
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.nvml, self.handle = pynvml, h
            return True
        except Exception:
            return False

    def _poll(self):
        nv, h = self.nvml, self.handle
        masks = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                 nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
        while not self.stop_flag.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                self.rows.append([mhz, self.max_mhz] + [bool(bits & m) for m in masks])
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self._nvml_open():
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            try:
                self.rows.append([float(c[0]), float(c[1])] + [v.lower().startswith("active") for v in c[3:7]])
            except (ValueError, IndexError):
                pass

    def mark(self):
        """Samples taken before this call (warm-up, idle) are dropped."""
        self.rows = []

    def stop(self):
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            how = "nvml"
        elif self.proc is not None:
            self.proc.terminate()
            how = "nvidia-smi"
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        rows = list(self.rows)
        sm = sorted(r[0] for r in rows)
        reasons = sorted({n for r in rows for n, v in zip(self.NAMES, r[2:6]) if v})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(r[1] for r in rows) if rows else self.max_mhz,
                "reasons": reasons, "samples": len(sm), "source": how}


# ------------------------------------------------------------------------------------------------ CPU arm
def force_host_threads():
    """All host threads for the CPU arm.  torch.distributed.run pre-sets OMP_NUM_THREADS=1 for its workers: override
    it (and the BLAS variables) BEFORE numpy / torch / the OpenMP oracle are loaded."""
    cores = os.cpu_count() or 1
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = str(cores)
    return cores


def cpu_step_fn(workload, B, seed=0):
    """One train step of the CPU arm (forward + ELBO + backward + Adam + radii SGD) in float32 on the host cores:
    oracle/cpu_baseline.py.  Returns (step function, threads in use)."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_baseline
    sig, _, D, H, recon, fixed, _ = WORKLOADS[workload]
    from mvae_b200 import components
    if workload in CONV_WORKLOADS:
        # convolutional model: ATen / oneDNN CPU convolutions under autograd + the C oracle for the latent chain
        from mvae_b200 import conv_vae, data
        torch.manual_seed(seed)
        m = conv_vae.FusedConvolutionalVAE(H, components.parse_components(sig, fixed), data.GenericDataset(B, D, recon),
                                           False, device="cpu")   # host-side construction only: default initialisation
        params = {k: v.detach().numpy().astype(np.float32).copy() for k, v in m.state_dict().items()}
        for k in params:
            if "radius" in k:
                params[k] = np.asarray(10.0, dtype=np.float32)
        cpu = cpu_baseline.CpuConvTrainStep(sig, params)
        x = synthetic_x(recon, B, D, seed)
        g = torch.Generator().manual_seed(seed)

        def conv_step():
            eps = torch.randn(B, cpu.desc.ld_eps, generator=g)
            return cpu.step(x, eps, beta=1.0)["elbo"]

        return conv_step, torch.get_num_threads()
    torch.manual_seed(seed)
    comps = components.parse_components(sig, fixed)
    # reference-shaped parameters with nn.Linear default init (CPU only; no kernels involved)
    params = {}
    for i, c in enumerate(comps):
        c.init_layers(H, False)
        for nm in ("fc_mean", "fc_logvar"):
            params[f"components.{i}.{nm}.weight"] = getattr(c, nm).weight.detach().numpy().copy()
            params[f"components.{i}.{nm}.bias"] = getattr(c, nm).bias.detach().numpy().copy()
        name, rp = c.radius_parameter()
        if rp is not None and not fixed:
            params[f"components.{i}.{name}"] = np.asarray(10.0, dtype=np.float32)  # as the GPU arm (--radius)
    tz = sum(c.dim for c in comps)
    for nm, (o, i_) in (("fc_e0", (H, D)), ("fc_d0", (H, tz)), ("fc_logits", (D, H))):
        lin = torch.nn.Linear(i_, o)
        params[nm + ".weight"] = lin.weight.detach().numpy().copy()
        params[nm + ".bias"] = lin.bias.detach().numpy().copy()
    cpu = cpu_baseline.CpuTrainStep(sig, D, H, recon, params, curvature_lr=1e-4 * min(1.0, 4096.0 / B))  # as the GPU arm
    x = synthetic_x(recon, B, D, seed)
    g = torch.Generator().manual_seed(seed)
    n_eps = cpu.desc.ld_eps

    def step():
        eps = torch.randn(B, n_eps, generator=g)
        return cpu.step(x, eps, beta=1.0)["elbo"]

    return step, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sig, B, D, H, recon, fixed, desc = WORKLOADS[args.workload]
    # bounded sample: a step is a full train step on the first Bs rows of the workload's batch, Bs chosen so that the
    # whole run (warm-up + K steps) stays near two minutes of CPU time
    Bs = B
    step, threads = cpu_step_fn(args.workload, Bs)
    step()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter() - t0
    budget_s = 120.0
    n_total = args.steps + max(args.warmup, 1)
    if t1 * n_total > budget_s:
        Bs = int(B * budget_s / (t1 * n_total)) // 128 * 128
        Bs = min(B, max(128, Bs))
        step, threads = cpu_step_fn(args.workload, Bs)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    value = Bs / (ms / 1e3)
    sample = (f"{args.steps} full train steps on {Bs} of the {B} rows of the batch ({desc}), float32, ATen CPU dense "
              f"layers + C/OpenMP oracle for the latent chain (oracle/cpu_baseline.py), {threads} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "steps_per_sec": 1e3 / ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "signature": sig, "batch_per_gpu": B, "global_batch": B * max(args.gpus, 1),
                       "in_dim": D, "h_dim": H, "parallelism": f"dp{max(args.gpus, 1)}",
                       "note": "bounded sample: the host cores run full train steps on ONE per-GPU share of the global "
                               "batch (the CPU arm is compute bound at these batch sizes: its samples/s changes little "
                               "with the number of shares it is given)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "omp_num_threads": os.environ.get("OMP_NUM_THREADS")},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def time_kernel(fn, iters, flush):
    import torch
    s = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    e = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for _ in range(3):
        fn()
    for i in range(iters):
        flush()
        s[i].record()
        fn()
        e[i].record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in zip(s, e)) / iters  # ms


def step_rooflines(model, opt, x, peaks, flush, C, Sn, Sd, P, H, B):
    """Rooflines of the kernels the TIMED STEP launches, at the config batch: every GEMM / latent / manifold launch of
    one eager step is recorded, then re-issued alone with a cold L2 (CUDA events, 10 runs each).  GEMMs: algorithmic
    fp32 flops 2 M N K against the measured bf16 tensor peak (the kernel issues 3 or 6 bf16 MMAs per product);
    latent / manifold kernels: algorithmic bytes against the measured HBM bandwidth."""
    import torch
    from mvae_b200 import ops
    names = ["gemm", "latent_forward", "latent_backward", "pm_forward", "pm_backward"]
    orig = {n: getattr(ops, n) for n in names}
    calls = []

    def wrap(name, fn):
        def inner(*a, **k):
            calls.append((name, fn, a, k))
            return fn(*a, **k)
        return inner

    graph = model.use_cuda_graph
    model.use_cuda_graph = False
    for n, f in orig.items():
        setattr(ops, n, wrap(n, f))
    try:
        model.train_step(opt, x, 1.0, sync_stats=False)
    finally:
        for n, f in orig.items():
            setattr(ops, n, f)
        model.use_cuda_graph = graph
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    bytes_of = {  # algorithmic bytes per sample (fp32 unless planes): what the kernel must read and write once
        "latent_forward": 4 * H + 4 * Sn + 4 * (P + Sd + C) + 2 * 2 * H,       # h, eps -> ml, z, kl, dd (2 planes)
        "latent_backward": 4 * H + 4 * H + 4 * (P + Sn + Sd) + 2 * 2 * H,      # gdd, h, ml, eps, z -> gh (2 planes)
        "pm_forward": 4 * (3 * Sn + Sd + C), "pm_backward": 4 * (5 * Sn + Sd)}
    rows, gemm_flops, gemm_us, gemm_issued = [], 0.0, 0.0, 0.0
    for name, fn, a, k in calls:
        ms = time_kernel(lambda: fn(*a, **k), 10, flush)
        if name == "gemm":
            M, N, K = a[2], a[3], a[4]
            fl = 2.0 * M * N * K
            gemm_flops += fl
            gemm_us += ms * 1e3
            # bf16 MMAs the kernel issues per algorithmic product: plane pairs (i, j) with i + j < max(planes)
            pa = k.get("a_planes") or a[0].planes
            pb = k.get("b_planes") or a[1].planes
            pa, pb = min(pa, a[0].planes), min(pb, a[1].planes)
            products = sum(1 for i in range(pa) for j in range(pb) if i + j < max(pa, pb))
            gemm_issued += fl * products
            rows.append({"kernel": "gemm_tcgen05_kernel", "shape": [M, N, K], "epilogue": int(k.get("epilogue", 0)),
                         "us": ms * 1e3, "tflops": fl / (ms * 1e-3) / 1e12,
                         "frac": fl / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"], "bf16_mmas_per_product": products,
                         "issued_frac": fl * products / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"]})
        else:
            by = bytes_of[name] * B
            rows.append({"kernel": name + "_kernel", "us": ms * 1e3, "gbs": by / (ms * 1e-3) / 1e9,
                         "frac": by / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "bytes_per_sample": bytes_of[name]})
    ops.add_launches(n0 - ops.launch_count())
    out = {"kernels": rows, "l2": "cold (flushed before every launch)"}
    if gemm_us:
        tf = gemm_flops / (gemm_us * 1e-6) / 1e12
        out["gemm_total"] = {"launches": sum(1 for r in rows if r["kernel"].startswith("gemm")), "us": gemm_us,
                             "tflops": tf, "peak": peaks["bf16_tflops"], "frac": tf / peaks["bf16_tflops"],
                             "issued_frac": gemm_issued / (gemm_us * 1e-6) / 1e12 / peaks["bf16_tflops"],
                             "bound": "tensor", "unit": "TFLOP/s",
                             "note": "frac: algorithmic fp32 flops; issued_frac: the bf16 MMAs actually issued "
                                     "(3 per product with 2 + 2 planes, 6 with 3 + 3)"}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mvae_b200 import components, data, ops, parallel, vae
    rank, world, local = parallel.init_from_env("nccl")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the mvae_b200 path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_bound = world > 1 and os.environ.get("MVAE_NUMA_BIND", "1") != "0" and parallel.bind_to_gpu_numa_node(local)
    sig, B, D, H, recon, fixed, desc = WORKLOADS[args.workload]
    peaks = measured_peaks()
    torch.manual_seed(0)
    comps = components.parse_components(sig, fixed)
    conv = args.workload in CONV_WORKLOADS
    if conv:
        from mvae_b200 import conv_vae
        model = conv_vae.FusedConvolutionalVAE(H, comps, data.GenericDataset(B, D, recon), False, device=dev)
    else:
        model = vae.FusedFeedForwardVAE(H, comps, data.GenericDataset(B, D, recon, binary_inputs=(recon == "bce")),
                                        False, device=dev)  # MNIST-shaped batches are binarised: one exact bf16 plane
    model.use_cuda_graph = not args.no_graph
    model.adopt_device_inputs = True   # `value`: the resident batches are read in place (one graph per batch tensor)
    # Learnable radii start at R = 10, the value the reference's own schedule gives them in its first training epoch
    # (Trainer._train_epoch: R = 11 - epoch for epoch < 10, train.py:189-194).  The ELBO is a SUM over the batch, so at
    # R = 1 the radius gradient of a 4096 x N batch times the reference's SGD step (1e-4) moves R by O(1) per step: with
    # N >= 2 it reaches the clamp at 1e-8 within a few hundred steps and the run trains on NaNs (seen at N = 2).
    with torch.no_grad():
        for rp in model._radius_params:
            if rp is not None and rp.requires_grad:
                rp.fill_(args.radius)
    # The ELBO is a SUM over the batch (stats.py:200-202), so the radius gradient grows with the global batch while the
    # reference's curvature step is plain SGD(1e-4) (train.py:343-355, tuned for batches of 100): at 4 and 8 ranks the
    # radii ran into their clamp within a few hundred steps and the run trained on NaNs.  The curvature step is
    # therefore divided by the number of ranks — the N-rank run then moves the radii like the 1-rank run does.  (Adam's
    # update of the network parameters is invariant to the gradient's scale.)
    # The same holds for the per-GPU batch: BDP-shaped data at B = 16384 (cfg4a) has a radius gradient of 1.2e5 at
    # initialisation — one reference-sized step takes R from 10 to -2.3 and the ELBO to NaN (reproduced with the CPU
    # oracle; rounds 1-2 reported elbo_finite = false there) — so beyond 4096 rows the step shrinks with the batch.
    curvature_lr = 1e-4 / world * min(1.0, 4096.0 / B)
    opt = vae.FusedCurvatureOptimizer(model, 1e-3, fixed_curvature=fixed, should_do_curvature_step=lambda: True,
                                      curvature_lr=curvature_lr)
    collective = "none"
    if world > 1:
        # one exchange per step: fused into the optimizer kernel over NVLink peer memory (default), or NCCL all-reduce
        if os.environ.get("MVAE_DP", "p2p") != "nccl" and parallel.attach_p2p(model, opt):
            collective = "peer-memory kernel: gradient reduce-scatter + Adam + parameter all-gather (mvae_dp_step)"
        else:
            parallel.attach(model)
            collective = "NCCL all-reduce(SUM) of the gradient/statistics bucket, then Adam"
        parallel.broadcast_parameters(model)
    C, Sn, Sd, P = model.desc.C, model.desc.ld_eps, model.desc.ld_z, model.desc.ld_ml
    # distinct batches per rank, rotated so that consecutive steps never see the same input
    n_rot = 4
    xs_host = [synthetic_x(recon, B, D, 1000 * rank + i).pin_memory() for i in range(n_rot)]
    xs_dev = [x.to(dev) for x in xs_host]
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def flush():
        flush_buf.zero_()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: device-resident inputs, no host sync inside the timed region ----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(max(args.warmup, n_rot)):   # every resident batch once at least: its graph is captured here
        model.train_step(opt, xs_dev[i % n_rot], 1.0, sync_stats=False)
    barrier()
    n0 = ops.launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    sampler.mark()
    t_host = time.perf_counter()
    for i in range(args.steps):
        flush()
        # the flush takes a different time on every rank and would leave the ranks misaligned when their steps
        # start — in training they start aligned (the previous step's exchange ends on all ranks together):
        # re-align them on the device, still outside the timed region
        if world > 1 and not args.no_rendezvous:
            parallel.rendezvous(opt)
        starts[i].record()
        model.train_step(opt, xs_dev[i % n_rot], 1.0, sync_stats=False)
        ends[i].record()
    host_enqueue_s = time.perf_counter() - t_host   # host time to enqueue the timed steps (no synchronisation inside)
    barrier()
    launches = ops.launch_count() - n0
    total_ms = sum(a.elapsed_time(b) for a, b in zip(starts, ends))
    own_total_ms = total_ms
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms = total_ms / args.steps
    clocks = sampler.stop() if rank == 0 else None
    # phases of the LAST timed step's exchange kernel on every rank (steady state), slowest rank per phase
    phases, per_rank = {}, {}
    host_ms = host_enqueue_s / args.steps * 1e3
    if world > 1:
        # per rank: device time of its own steps, host time to ENQUEUE a step, SM clock now; and (peer-memory path) how
        # long its last exchange kernel waited for the peers' gradients — the rank that waits least arrives last
        sm_now = 0.0
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(local).uuid)
            sm_now = float(pynvml.nvmlDeviceGetClockInfo(pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode()),
                                                         pynvml.NVML_CLOCK_SM))
        except Exception:
            pass
        ph = parallel.dp_phase_times(opt) if not collective.startswith("NCCL") else {}
        mine = torch.tensor([own_total_ms / args.steps, host_ms, sm_now, ph.get("late_wait_grads", 0.0),
                             ph.get("late_total", 0.0)], device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        cols = list(zip(*[[round(float(v), 4) for v in t.tolist()] for t in allr]))
        per_rank = {"device_ms_per_step": cols[0], "host_enqueue_ms_per_step": cols[1], "sm_mhz_after": cols[2],
                    "late_wait_grads_us": cols[3], "late_total_us": cols[4]}
        if ph:
            keys = sorted(ph)
            t = torch.tensor([ph[k] for k in keys], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            phases = {k: round(float(v), 2) for k, v in zip(keys, t.tolist())}
    stats_vec = model._stats.clone()

    # ---------------- e2e: public API with host buffers (H2D of x, D2H of the statistics, every step) ----------------
    # model.train_epoch(optimizer, batches, beta) is the batch loop of the reference's Trainer._train_epoch
    # (train.py:198-210): every step copies ITS batch from pinned host memory and returns ITS statistics to the host;
    # the copy of batch i+1 overlaps the kernels of step i.  The serial form (one train_step per call, blocking on the
    # statistics) is timed as well and reported as e2e.serial.
    # Image workloads ship the batch as the dataset stores it — uint8 grayscale pixels, 1 byte per pixel — and binarise
    # it on the device (mvae_binarize: the reference's ImageDynamicBinarization, which runs per sample on the CPU in its
    # DataLoader); float32 host batches (the reference's loader output, 4 bytes per pixel) are timed as well.
    u8_inputs = recon == "bce" and not args.float_inputs and not conv
    xs_e2e = [synthetic_pixels(B, D, 1000 * rank + i).pin_memory() for i in range(n_rot)] if u8_inputs else xs_host

    def host_batches(n, src):
        for i in range(n):
            yield src[i % n_rot]

    def timed_epoch(src):
        model.train_epoch(opt, host_batches(max(args.warmup, 3), src), 1.0)
        barrier()
        t0 = time.perf_counter()
        out = model.train_epoch(opt, host_batches(args.steps, src), 1.0)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, out

    e2e_s, stats_list = timed_epoch(xs_e2e)
    bs = stats_list[-1]
    assert len(stats_list) == args.steps
    float_s = timed_epoch(xs_host)[0] if u8_inputs else e2e_s
    for i in range(3):
        model.train_step(opt, xs_e2e[i % n_rot], 1.0)
    barrier()
    n_serial = min(args.steps, 200)
    t0 = time.perf_counter()
    for i in range(n_serial):
        model.train_step(opt, xs_e2e[i % n_rot], 1.0)
    torch.cuda.synchronize()
    serial_s = (time.perf_counter() - t0) / n_serial
    # strict drop-in pattern: the reference's literal call, `stats, _ = model.train_step(optimizer, x_mb, beta)` with the
    # float32 batch its DataLoader yields, blocking on the statistics every step (train.py:197-198)
    for i in range(3):
        model.train_step(opt, xs_host[i % n_rot], 1.0)
    barrier()
    t0 = time.perf_counter()
    for i in range(n_serial):
        model.train_step(opt, xs_host[i % n_rot], 1.0)
    torch.cuda.synchronize()
    strict_s = (time.perf_counter() - t0) / n_serial
    if world > 1:
        t = torch.tensor([e2e_s, serial_s, float_s, strict_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, serial_s, float_s, strict_s = (float(v) for v in t.tolist())

    # ---------------- N > 1: the run proves itself ----------------
    dp_check = None
    if world > 1:
        # (a) replicas bit-identical after everything above; (b) the sticky error word of mvae_dp_step;
        # (c) one more step with supplied noise: the statistics the exchange produced (rank-summed inside the fused
        #     kernel) against an NCCL all-reduce of every rank's own statistics
        identical = parallel.replicas_identical(model)
        g = torch.Generator(device=dev).manual_seed(4242 + rank)
        eps_chk = torch.randn(B, Sn, device=dev, generator=g)
        bs_chk, _ = model.train_step(opt, xs_dev[0], 1.0, eps=eps_chk)
        local = model._stats.clone()          # this rank's [bce, kl, elbo, kl_c] (the tail of its own bucket)
        if collective.startswith("NCCL"):
            local = None                      # the bucket itself was all-reduced in place
        rel = None
        if local is not None:
            dist.all_reduce(local, op=dist.ReduceOp.SUM)
            rel = abs(bs_chk.elbo - float(local[2].item())) / abs(float(local[2].item()))
        dp_check = {"replicas_identical": bool(identical and parallel.replicas_identical(model)),
                    "phases_us_max_over_ranks": phases, "per_rank": per_rank,
                    "elbo_vs_nccl_rel": rel, "dp_error_word": parallel.dp_error_word(opt),
                    "overlap": bool(getattr(opt, "dp_overlap", False) and opt._dp is not None)}
    e2e_ms = e2e_s / args.steps * 1e3
    float_ms = float_s / args.steps * 1e3

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    gb = B * world
    line = {"metric": METRIC, "value": gb / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, n_rot), "ms_per_step": ms, "steps_per_sec": 1e3 / ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (split-bf16 tensor-core GEMMs, fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": desc, "signature": sig, "batch_per_gpu": B, "global_batch": gb, "in_dim": D,
                       "h_dim": H, "parallelism": f"dp{world}", "l2": "flushed between timed steps (256 MiB memset)" +
                       ("; ranks re-aligned after the flush by a peer-memory rendezvous kernel, outside the timed region"
                        if world > 1 and not args.no_rendezvous and not collective.startswith("NCCL") else ""),
                       "cuda_graph": bool(model.use_cuda_graph),
                       "optimizer": f"Adam(1e-3) + SGD({curvature_lr:g} = 1e-4 / ranks x min(1, 4096 / batch per GPU)) on radii",
                       "initial_radius": args.radius,
                       "collective": collective, "numa_bound": bool(numa_bound)},
            "e2e": {"value": gb / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": B * D * (1 if u8_inputs else 4), "d2h_bytes_per_step": (3 + C) * 4,
                    "api": "model.train_epoch(optimizer, pinned host batches, beta): H2D of batch i+1 overlaps step i",
                    "inputs": ("uint8 grayscale batches as the dataset stores them, dynamic binarisation on the device "
                               "(mvae_binarize = ImageDynamicBinarization, image_reconstruction.py:37-53)") if u8_inputs
                    else "float32 batches",
                    "float32_batches": {"value": gb / (float_ms / 1e3), "ms_per_step": float_ms,
                                        "h2d_bytes_per_step": B * D * 4},
                    "serial": {"value": gb / serial_s, "ms_per_step": serial_s * 1e3,
                               "api": "model.train_step(optimizer, x_host, beta), blocking"},
                    "strict": {"value": gb / strict_s, "ms_per_step": strict_s * 1e3, "h2d_bytes_per_step": B * D * 4,
                               "api": "the reference's literal call pattern: model.train_step(optimizer, float32 host "
                                      "batch, beta), blocking on the statistics every step (train.py:197-198)"}},
            "gpu_launches": int(launches), "clocks": clocks, "elbo_per_sample": float(bs.elbo) / gb,
            "host_enqueue_ms_per_step": host_ms,
            "elbo_finite": bool(all(s.elbo == s.elbo and abs(s.elbo) < float("inf") for s in stats_list)),
            "peaks": peaks["source"]}
    if dp_check is not None:
        line["dp_check"] = dp_check

    # ---------------- roofline of the fused product-manifold kernel (HBM bound) ----------------
    if not args.skip_roofline:
        bytes_fwd = 4 * (3 * Sn + Sd + C)   # read ml (2 Sn = m|l), eps (Sn); write z (Sd), kl (C)   [SURVEY §8d]
        bytes_bwd = 4 * (5 * Sn + Sd)       # read ml, eps, gz; write gml
        Bbig = 1 << 22
        g = torch.Generator(device=dev).manual_seed(0)
        ml = torch.randn(Bbig, P, device=dev, generator=g) * 0.5
        eps = torch.randn(Bbig, Sn, device=dev, generator=g)
        z = torch.empty(Bbig, Sd, device=dev)
        kl = torch.empty(Bbig, C, device=dev)
        out = {"z": z, "kl": kl}
        ms_f = time_kernel(lambda: ops.pm_forward(model.desc, ml, eps, model._rflat, out=out), 20, flush)
        gz = torch.randn(Bbig, Sd, device=dev, generator=g)
        gml = torch.empty_like(ml)
        gR = torch.zeros(C, device=dev)
        ms_b = time_kernel(lambda: ops.pm_backward(model.desc, ml, eps, model._rflat, gz, None, 1.0, gml=gml, gradius=gR),
                           20, flush)
        ws = model._workspace(B)
        ms_f_cfg = time_kernel(lambda: ops.pm_forward(model.desc, ws.ml, ws.eps, model._rflat,
                                                      out={"z": ws.z, "kl": ws.kl}), 20, flush)
        ach = Bbig * bytes_fwd / (ms_f * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic("pm_forward_kernel", sig, Bbig)
        line["roofline"] = {"kernel": "pm_forward_kernel", "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"],
                            "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                            "traffic": traffic, "traffic_source": traffic_src,
                            "regime": "asymptotic: 2^22 samples per launch, not a launch of the timed step (the step "
                                      "runs the fused latent block / this kernel at the config batch: roofline_step)",
                            "algorithmic_bytes": Bbig * bytes_fwd,
                            "bytes_per_sample": bytes_fwd, "samples_per_launch": Bbig, "us_per_launch": ms_f * 1e3,
                            "us_at_config_batch": ms_f_cfg * 1e3, "peak_source": peaks["source"]}
        achb = Bbig * bytes_bwd / (ms_b * 1e-3) / 1e9
        traffic_b, traffic_b_src = ncu_traffic("pm_backward_kernel", sig, Bbig)
        line["roofline_backward"] = {"kernel": "pm_backward_kernel", "bound": "hbm", "achieved": achb,
                                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achb / peaks["hbm_gbs"],
                                     "traffic": traffic_b, "traffic_source": traffic_b_src,
                                     "bytes_per_sample": bytes_bwd, "samples_per_launch": Bbig,
                                     "us_per_launch": ms_b * 1e3}
        del ml, eps, z, kl, gz, gml
        # tcgen05 GEMM (tensor bound): encoder layer of the workload, algorithmic flops 2*B*D*H
        if conv:   # the largest forward GEMM of the convolutional stack: e2 as [B*16, 2048] x [512, 2048]^T
            gm, gn, gk, what, products = (B * 16, 512, 2048,
                                          "e2 forward: im2col(a1) . W^T, 3 + 3 planes = 6 bf16 MMAs per product", 6)
            ms_g = time_kernel(lambda: ops.gemm(ws.A[2], model._Wp["e2"], gm, gn, gk, epilogue=1,
                                                bias=model._bias["e2"], out_planes=ws.a[2], tile=(128, 1)), 20, flush)
        else:
            products = 3 if model.input_planes == 1 else 6   # x planes (1 for binarised inputs, else 3) x 3 weight planes
            gm, gn, gk, what = B, H, D, f"fc_e0 forward; {products} bf16 MMAs per product (split planes)"
            ms_g = time_kernel(lambda: ops.gemm(ws.xp, model.We0p, B, H, D, epilogue=1, bias=model.fc_e0.bias.data,
                                                out_planes=ws.hp), 20, flush)
        tf = 2.0 * gm * gn * gk / (ms_g * 1e-3) / 1e12
        line["roofline_gemm"] = {"kernel": "gemm_tcgen05_kernel", "shape": [gm, gn, gk], "bound": "tensor",
                                 "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                 "frac": tf / peaks["bf16_tflops"], "us_per_launch": ms_g * 1e3,
                                 "bf16_mmas_per_product": products,
                                 "issued_frac": tf * products / peaks["bf16_tflops"],
                                 "note": "frac: algorithmic fp32 flops; issued_frac: the bf16 MMAs actually issued; " + what}

        line["roofline_step"] = step_rooflines(model, opt, xs_dev[0], peaks, flush, C, Sn, Sd, P, H, B)

    # ---------------- CPU baseline (oracle/cpu_baseline.py on the host cores), bounded sample ----------------
    if world == 1 and not args.skip_cpu:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        step, threads = cpu_step_fn(args.workload, B)
        step()
        step()
        n_cpu, t0 = 0, time.perf_counter()
        while n_cpu < 200 and time.perf_counter() - t0 < 12.0:   # ~12 s of CPU work
            step()
            n_cpu += 1
        dt = (time.perf_counter() - t0) / n_cpu
        line["cpu_baseline"] = {"value": B / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{n_cpu} full train steps of batch {B}, float32, ATen CPU dense layers + "
                                          "C/OpenMP oracle for the latent chain (oracle/cpu_baseline.py)",
                                "ms_per_step": dt * 1e3, "omp_num_threads": os.environ.get("OMP_NUM_THREADS")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--radius", type=float, default=10.0, help="initial value of the learnable radii")
    ap.add_argument("--no-rendezvous", action="store_true",
                    help="N > 1: do not re-align the ranks after the L2 flush that precedes every timed step")
    ap.add_argument("--skip-roofline", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--float-inputs", action="store_true",
                    help="e2e with float32 host batches (4 bytes per pixel) instead of uint8 pixels binarised on the device")
    args = ap.parse_args()
    if args.impl == "reference" or int(os.environ.get("WORLD_SIZE", "1")) == 1:
        force_host_threads()  # the CPU legs use every host core (torchrun pre-sets OMP_NUM_THREADS=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
