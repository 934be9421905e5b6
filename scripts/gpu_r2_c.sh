#!/bin/bash
# Round 2, run C (two GPUs): timed-path tests + data-parallel check, bench N=1/2 with phase timings of mvae_dp_step.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 900 python -m pytest tests/test_gpu_timed_path.py -m gpu -q -p no:cacheprovider -s > gpurun_out/test_timed.log 2>&1; echo "tests rc=$?"
grep -E "^\[|passed|failed|FAILED|Error|p2p|nccl " gpurun_out/test_timed.log | tail -30
timeout -s KILL 300 python bench.py --gpus 1 --steps 500 --warmup 20 --skip-roofline --skip-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"
timeout -s KILL 300 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 500 --warmup 20 --skip-roofline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
MVAE_DP_EARLY_CTAS=8 timeout -s KILL 300 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 500 --warmup 20 --skip-roofline > gpurun_out/bench_n2_e8.json 2> gpurun_out/bench_n2_e8.err; echo "bench2 e8 rc=$?"
MVAE_DP_OVERLAP=0 timeout -s KILL 300 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 500 --warmup 20 --skip-roofline > gpurun_out/bench_n2_nooverlap.json 2> gpurun_out/bench_n2_nooverlap.err; echo "bench2 no-overlap rc=$?"
python - <<'PY'
import json
for f in ("bench_n1", "bench_n2", "bench_n2_e8", "bench_n2_nooverlap"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"] / 1e6, 2), "M  e2e ms", round(d["e2e"]["ms_per_step"], 4),
              "strict", round(d["e2e"]["strict"]["ms_per_step"], 4), "finite", d.get("elbo_finite"), d.get("dp_check"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/bench_n2.err
