#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 300 $TR --nproc-per-node 2 --master-port 29511 scripts/dp_check.py > gpurun_out/dp_check.log 2>&1; echo "dp_check rc=$?"
grep -v "^W\|^\*\*\*" gpurun_out/dp_check.log | tail -8
timeout -s KILL 300 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 500 --warmup 20 --skip-roofline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for f in ("bench_n2",):
    d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
    print(f, "ms/step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 4), d.get("dp_check"))
PY
