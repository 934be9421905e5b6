#!/bin/bash
# On the GPU box: smoke, bench (graphs / eager), and optionally ncu captures -> gpurun_out/
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee gpurun_out/bench_summary.txt
timeout -s KILL 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "bench rc=$?" | tee -a gpurun_out/bench_summary.txt
timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --no-graph --skip-roofline --skip-cpu > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; echo "bench eager rc=$?" | tee -a gpurun_out/bench_summary.txt
if [ "$1" == "ncu" ]; then
  timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-graph --skip-roofline --skip-cpu > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?" | tee -a gpurun_out/bench_summary.txt
fi
tail -5 gpurun_out/smoke.log
cat gpurun_out/bench_graph.json; tail -5 gpurun_out/bench_graph.err
cat gpurun_out/bench_eager.json; tail -5 gpurun_out/bench_eager.err
