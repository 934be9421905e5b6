#!/bin/bash
# On the GPU box: parity tests of the product-manifold kernels, then the microbenchmark at both tile heights.
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "pm or golden or flag or desc" > gpurun_out/test_pm.log 2>&1; echo "pm tests rc=$?"
tail -15 gpurun_out/test_pm.log
timeout -s KILL 300 python scripts/pm_bench.py 2>&1 | tee gpurun_out/pm_bench_auto.log
MVAE_PM_TILE=32 timeout -s KILL 300 python scripts/pm_bench.py "h2,s2,e2" "h6,h6,s6,s6,e6" 2>&1 | tee gpurun_out/pm_bench_32.log
