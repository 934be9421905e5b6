#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python scripts/host_profile.py 2>&1 | tail -10
timeout -s KILL 900 python -m pytest tests/test_gpu_timed_path.py tests/test_gpu_model.py -m gpu -q -p no:cacheprovider > gpurun_out/test_timed.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/test_timed.log
ARGS="--steps 500 --warmup 20 --skip-roofline --skip-cpu"
timeout -s KILL 300 python bench.py --gpus 1 $ARGS > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"
timeout -s KILL 300 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 $ARGS > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for f in ("bench_n1", "bench_n2"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        e = d["e2e"]
        print(f, "ms/step", round(d["ms_per_step"], 4), "e2e ms", round(e["ms_per_step"], 4), "float32", round(e["float32_batches"]["ms_per_step"], 4),
              "serial", round(e["serial"]["ms_per_step"], 4), "strict", round(e["strict"]["ms_per_step"], 4), "finite", d.get("elbo_finite"))
    except Exception as ex:
        print(f, "failed", ex)
PY
