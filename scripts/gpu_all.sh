#!/bin/bash
# On the GPU box: full GPU test suite, smoke, step breakdown, pm microbench, bench.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/test_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2
timeout -s KILL 300 python scripts/step_breakdown.py cfg2 2>&1 | tee gpurun_out/step_breakdown_cfg2.log
MVAE_PM_DEBUG=1 timeout -s KILL 300 python scripts/pm_bench.py "h2,s2,e2" "h6,h6,s6,s6,e6" 2>&1 | grep -v "B=    8192\|B=   16384" | awk '!seen[$0]++' | tee gpurun_out/pm_bench_auto.log
timeout -s KILL 600 python bench.py --steps 300 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
