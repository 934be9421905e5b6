#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/test_gpu.log
timeout -s KILL 300 python scripts/step_breakdown.py cfg2 2>&1 | tee gpurun_out/step_breakdown_cfg2.log
timeout -s KILL 300 python bench.py --steps 500 --warmup 20 --skip-roofline --skip-cpu 2>gpurun_out/bench_quick.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'elbo', d['elbo_per_sample'])" || tail -5 gpurun_out/bench_quick.err
