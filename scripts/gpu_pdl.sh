#!/bin/bash
mkdir -p gpurun_out
echo skip tests

for pdl in 1 0; do
  MVAE_PDL=$pdl timeout -s KILL 300 python bench.py --steps 500 --warmup 20 --skip-roofline --skip-cpu 2>gpurun_out/bench_pdl$pdl.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('PDL=$pdl', 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), 'elbo', d['elbo_per_sample'])" || tail -5 gpurun_out/bench_pdl$pdl.err
done
