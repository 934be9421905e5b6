#!/bin/bash
# On the GPU box: parity tests of the product-manifold kernels, then the microbenchmark over tile shapes.
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "pm or golden or flag or desc" > gpurun_out/test_pm.log 2>&1; echo "pm tests rc=$?"
tail -5 gpurun_out/test_pm.log
MVAE_PM_DEBUG=1 timeout -s KILL 300 python scripts/pm_bench.py "h2,s2,e2" "h6,h6,s6,s6,e6" "h2" "e2" 2>&1 | grep -v "B=    8192\|B=   16384" | awk '!seen[$0]++' | tee gpurun_out/pm_bench_auto.log
rm -f gpurun_out/pm_bench_tune.log
for t in "1,4" "2,4" "3,4" "4,4" "6,3" "6,4" "8,3" "8,2"; do
  MVAE_PM_TUNE=$t timeout -s KILL 300 python scripts/pm_bench.py "h2,s2,e2" "h6,h6,s6,s6,e6" 2>&1 | grep "4194304" | sed "s/tile=auto/tune=$t/" | tee -a gpurun_out/pm_bench_tune.log
done
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:pm_ -s 6 -c 2 -f -o gpurun_out/prof_pm python scripts/prof_driver.py pm "h2,s2,e2" > gpurun_out/ncu_pm.log 2>&1
echo "ncu pm rc=$?"
