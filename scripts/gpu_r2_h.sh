#!/bin/bash
# Round 2, run H (one GPU): ncu launch list + full captures, summarised ON the box (the .ncu-rep files are too large to
# travel back), plus the conv k-step test and the per-workload bench lines.
mkdir -p gpurun_out gpurun_out/profiles_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches.csv
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/test_conv.log 2>&1; echo "gpu tests rc=$?"
grep -E "^\[|passed|failed|FAILED" gpurun_out/test_conv.log | tail -8
timeout -s KILL 500 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench cfg2 rc=$?"
for wl in cfg1 cfg3 cfg4a cfg4b cfg5; do
  timeout -s KILL 400 python bench.py --workload $wl --steps 300 --warmup 20 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"
done
timeout -s KILL 300 python scripts/step_breakdown.py cfg2 > gpurun_out/step_breakdown_cfg2.log 2>&1
timeout -s KILL 300 python scripts/step_breakdown.py cfg3 > gpurun_out/step_breakdown_cfg3.log 2>&1
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/prof_step.py cfg2 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_step python scripts/prof_step.py cfg2 > gpurun_out/ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:pm_ -s 6 -c 2 -f -o gpurun_out/prof_pm python scripts/prof_driver.py pm "h2,s2,e2" > gpurun_out/ncu_pm.log 2>&1; echo "ncu pm rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:pm_ -s 6 -c 2 -f -o gpurun_out/prof_pm_cfg3 python scripts/prof_driver.py pm "h6,h6,s6,s6,e6" > gpurun_out/ncu_pm3.log 2>&1; echo "ncu pm cfg3 rc=$?"
MVAE_PROF_DIR=gpurun_out/profiles_out python scripts/summarize_profiles.py 02 b > gpurun_out/summarize.log 2>&1; echo "summarize rc=$?"
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/profiles_out; du -sh gpurun_out
