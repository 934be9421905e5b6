#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 900 python -m pytest tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider -s > gpurun_out/test_conv.log 2>&1; echo "conv tests rc=$?"
grep -E "^\[|passed|failed|FAILED|Error|^E  " gpurun_out/test_conv.log | tail -30
timeout -s KILL 300 $TR --nproc-per-node 2 --master-port 29511 scripts/dp_check.py > gpurun_out/dp_check.log 2>&1; echo "dp_check rc=$?"
grep -v "^W\|^\*\*\*" gpurun_out/dp_check.log | tail -4
timeout -s KILL 300 python bench.py --gpus 1 --steps 500 --warmup 20 --skip-roofline --skip-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"
timeout -s KILL 300 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 500 --warmup 20 --skip-roofline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
timeout -s KILL 400 python bench.py --workload cfg5 --steps 100 --warmup 10 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "bench cfg5 rc=$?"
tail -3 gpurun_out/bench_cfg5.err
python - <<'PY'
import json
for f in ("bench_n1", "bench_n2", "bench_cfg5"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "ms/step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 4), d.get("dp_check"), d.get("cpu_baseline", {}).get("ms_per_step"))
        if f == "bench_cfg5":
            for r in d.get("roofline_step", {}).get("kernels", []):
                print("   ", r)
    except Exception as e:
        print(f, "failed", e)
PY
