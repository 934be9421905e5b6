#!/bin/bash
# On the GPU box: ncu launch list of one eager step + full captures of the product-manifold kernels -> gpurun_out/
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-graph --skip-roofline --skip-cpu > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:pm_ -s 6 -c 2 -f -o gpurun_out/prof_pm python scripts/prof_driver.py pm "h2,s2,e2" > gpurun_out/ncu_pm.log 2>&1
echo "ncu pm rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:pm_ -s 6 -c 2 -f -o gpurun_out/prof_pm_cfg3 python scripts/prof_driver.py pm "h6,h6,s6,s6,e6" > gpurun_out/ncu_pm3.log 2>&1
echo "ncu pm cfg3 rc=$?"
ls -la gpurun_out
