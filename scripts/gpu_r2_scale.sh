#!/bin/bash
# Round 2, scaling run (eight GPUs): cfg2 at N = 1/2/4/8, cfg3 at N = 1/8, cfg5 at N = 8; dp_check at N = 8.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
ARGS="--steps 500 --warmup 20 --skip-roofline --skip-cpu"
timeout -s KILL 300 python bench.py --gpus 1 $ARGS > gpurun_out/scale_cfg2_n1.json 2> gpurun_out/scale_cfg2_n1.err; echo "cfg2 n1 rc=$?"
for n in 2 4 8; do
  timeout -s KILL 300 $TR --nproc-per-node $n --master-port 2951$n bench.py --gpus $n $ARGS > gpurun_out/scale_cfg2_n$n.json 2> gpurun_out/scale_cfg2_n$n.err; echo "cfg2 n$n rc=$?"
done
timeout -s KILL 300 python bench.py --gpus 1 --workload cfg3 $ARGS > gpurun_out/scale_cfg3_n1.json 2> gpurun_out/scale_cfg3_n1.err; echo "cfg3 n1 rc=$?"
for n in 2 4 8; do
  timeout -s KILL 300 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --workload cfg3 $ARGS > gpurun_out/scale_cfg3_n$n.json 2> gpurun_out/scale_cfg3_n$n.err; echo "cfg3 n$n rc=$?"
done
timeout -s KILL 300 $TR --nproc-per-node 8 --master-port 29538 bench.py --gpus 8 --workload cfg5 --steps 200 --warmup 10 --skip-roofline --skip-cpu > gpurun_out/scale_cfg5_n8.json 2> gpurun_out/scale_cfg5_n8.err; echo "cfg5 n8 rc=$?"
timeout -s KILL 300 $TR --nproc-per-node 8 --master-port 29539 scripts/dp_check.py > gpurun_out/dp_check_n8.log 2>&1; echo "dp_check n8 rc=$?"
grep -v "^W\|^\*\*\*\|UserWarning\|return func" gpurun_out/dp_check_n8.log | tail -8
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step", round(d["ms_per_step"], 4), "M/s", round(d["value"] / 1e6, 2), "e2e ms", round(d["e2e"]["ms_per_step"], 4),
              "e2e M/s", round(d["e2e"]["value"] / 1e6, 2), "finite", d.get("elbo_finite"), json.dumps(d.get("dp_check")))
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json", ".err")).read()[-400:])
PY
