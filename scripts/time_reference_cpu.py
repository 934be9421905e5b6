"""Build-container only (needs /root/reference): times the UNMODIFIED reference's ModelVAE.train_step
(mt/mvae/models/vae.py:149-166, optimizer of Trainer.build_optimizer, train.py:327-360) on the host cores next to the
oracle port that bench.py's CPU arm runs on the GPU box (where the reference does not exist), on the same workload,
dtype and thread count.  Output goes into BASELINE.md §3 so that the port's number has context.

    python scripts/time_reference_cpu.py [workload] [steps]"""
import json
import os
import sys
import time

cores = os.cpu_count() or 1
for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[k] = str(cores)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

import bench  # noqa: E402
import ref_harness as rh  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
sig, B, D, H, recon, fixed, desc = bench.WORKLOADS[wl]
torch.set_num_threads(cores)
rh.load_reference()
from mt.mvae.models import train as rt  # noqa: E402

res = {"workload": desc, "cores": cores, "torch": torch.__version__, "steps": steps}
for dtype, name in ((torch.float32, "f32"), (torch.float64, "f64")):
    model = rh.build_model(sig, D, H, fixed, False, recon, 0, dtype)

    class _Stats:
        epoch, global_step = 12, 0

    class _T:
        epoch = property(lambda self: self.stats.epoch)

    t = _T()
    t.model, t.stats = model, _Stats()
    opt = rt.Trainer.build_optimizer(t, learning_rate=1e-3, fixed_curvature=fixed)
    x = bench.synthetic_x(recon, B, D, 0).to(dtype)
    with rh.default_dtype(dtype):
        for _ in range(2):
            model.train_step(opt, x, beta=1.0)
        t0 = time.perf_counter()
        for _ in range(steps):
            model.train_step(opt, x, beta=1.0)
        ms = (time.perf_counter() - t0) / steps * 1e3
    res[f"reference_{name}_ms_per_step"] = ms
    res[f"reference_{name}_samples_per_s"] = B / ms * 1e3
step, _ = bench.cpu_step_fn(wl, B)
for _ in range(2):
    step()
t0 = time.perf_counter()
for _ in range(steps):
    step()
ms = (time.perf_counter() - t0) / steps * 1e3
res["port_f32_ms_per_step"] = ms
res["port_f32_samples_per_s"] = B / ms * 1e3
res["port_over_reference_f32"] = res["port_f32_ms_per_step"] / res["reference_f32_ms_per_step"]
print(json.dumps(res, indent=1))
