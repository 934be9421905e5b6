#!/bin/bash
# Round 2, final run (one GPU, ~2 minutes of budget left): ncu launch list + full capture of one eager cfg2 step with the
# final kernels, then bench lines of the other workloads as far as the time allows.  Summaries are made afterwards from
# the files that come back (scripts/summarize_profiles.py runs without a GPU).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches.csv
timeout -s KILL 60 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/prof_step.py cfg2 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout -s KILL 90 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_step python scripts/prof_step.py cfg2 > gpurun_out/ncu_step.log 2>&1; echo "ncu step rc=$?"
for wl in cfg3 cfg5 cfg1 cfg4a cfg4b; do
  timeout -s KILL 60 python bench.py --workload $wl --skip-roofline --skip-cpu > gpurun_out/final_bench_$wl.json 2> gpurun_out/final_bench_$wl.err; echo "bench $wl rc=$?"
done
ls -la gpurun_out | tail -12
