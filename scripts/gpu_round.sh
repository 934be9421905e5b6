#!/bin/bash
# On the GPU box: GPU tests, bench, ncu launch list of one step, full ncu capture of that step's kernels -> gpurun_out/
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/test_gpu.log
timeout -s KILL 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout -s KILL 300 python scripts/step_breakdown.py cfg2 2>&1 | tee gpurun_out/step_breakdown_cfg2.log | tail -16
if [ "$1" == "ncu" ]; then
  timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/prof_step.py cfg2 > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_step python scripts/prof_step.py cfg2 > gpurun_out/ncu_step.log 2>&1
  echo "ncu step rc=$?"
  ls -la gpurun_out | tail -12
fi
