#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
ARGS="--steps 500 --warmup 20 --skip-roofline --skip-cpu"
timeout -s KILL 300 $TR --nproc-per-node 4 --master-port 29511 scripts/dp_check.py > gpurun_out/dp_check.log 2>&1; echo "dp_check rc=$?"
grep -v "^W\|^\*\*\*\|UserWarning\|return func" gpurun_out/dp_check.log | tail -7
timeout -s KILL 300 python bench.py --gpus 1 $ARGS > gpurun_out/p_n1.json 2> gpurun_out/p_n1.err; echo "n1 rc=$?"
for n in 2 4; do
  timeout -s KILL 300 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n $ARGS > gpurun_out/p_n$n.json 2> gpurun_out/p_n$n.err; echo "n$n rc=$?"
done
python - <<'PY'
import json
for f in ("p_n1", "p_n2", "p_n4"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "ms/step", round(d["ms_per_step"], 4), "e2e ms", round(d["e2e"]["ms_per_step"], 4), "finite", d.get("elbo_finite"), json.dumps((d.get("dp_check") or {}).get("phases_us_max_over_ranks")))
    except Exception as e:
        print(f, "failed", e)
PY
