#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 500 compute-sanitizer --tool racecheck --target-processes all --print-limit 10 $TR --nproc-per-node 2 --master-port 29515 scripts/sanitize_step.py dp > gpurun_out/sanitizer_racecheck_dp.log 2>&1; echo "racecheck dp rc=$?"
grep -E "RACECHECK SUMMARY|sanitize_step|elbo|Error" gpurun_out/sanitizer_racecheck_dp.log | tail -8
timeout -s KILL 500 compute-sanitizer --tool memcheck --target-processes all --print-limit 10 $TR --nproc-per-node 2 --master-port 29516 scripts/sanitize_step.py dp > gpurun_out/sanitizer_memcheck_dp.log 2>&1; echo "memcheck dp rc=$?"
grep -E "ERROR SUMMARY|sanitize_step" gpurun_out/sanitizer_memcheck_dp.log | tail -6
timeout -s KILL 400 compute-sanitizer --tool memcheck --print-limit 10 python scripts/sanitize_step.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|sanitize_step" gpurun_out/sanitizer_memcheck.log | tail -3
for wl in cfg4a cfg4b; do
  python bench.py --workload $wl --steps 300 --warmup 20 --skip-roofline --skip-cpu | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"
done
