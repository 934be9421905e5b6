#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/test_gpu.log | tail -8
ARGS="--steps 500 --warmup 20 --skip-roofline --skip-cpu"
timeout -s KILL 300 python bench.py --gpus 1 $ARGS > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"
MVAE_PDL=1 timeout -s KILL 300 python bench.py --gpus 1 $ARGS > gpurun_out/bench_n1_pdl.json 2> gpurun_out/bench_n1_pdl.err; echo "bench1 pdl rc=$?"
timeout -s KILL 300 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 $ARGS > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
MVAE_PDL=1 timeout -s KILL 300 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 $ARGS > gpurun_out/bench_n2_pdl.json 2> gpurun_out/bench_n2_pdl.err; echo "bench2 pdl rc=$?"
python - <<'PY'
import json
for f in ("bench_n1", "bench_n1_pdl", "bench_n2", "bench_n2_pdl"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "ms/step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 4), "host enqueue ms", round(d.get("host_enqueue_ms_per_step", 0), 4), "finite", d.get("elbo_finite"),
              json.dumps((d.get("dp_check") or {}).get("per_rank")), json.dumps((d.get("dp_check") or {}).get("phases_us_max_over_ranks")))
    except Exception as e:
        print(f, "failed", e)
PY
