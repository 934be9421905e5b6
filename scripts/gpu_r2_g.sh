#!/bin/bash
# Round 2, run G (one GPU): the whole GPU suite, bench lines of every BASELINE workload, ncu launch list + full captures.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches.csv
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/test_gpu.log 2>&1; echo "tests rc=$?"
grep -E "^\[|passed|failed|FAILED" gpurun_out/test_gpu.log | tail -20
timeout -s KILL 500 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench cfg2 rc=$?"
for wl in cfg1 cfg3 cfg4a cfg4b cfg5; do
  timeout -s KILL 400 python bench.py --workload $wl --steps 300 --warmup 20 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"
done
python - <<'PY'
import json
for f in ("cfg2", "cfg1", "cfg3", "cfg4a", "cfg4b", "cfg5"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/bench_{f}.json") if l.startswith("{")][-1])
        print(f, "ms/step", round(d["ms_per_step"], 4), "M samples/s", round(d["value"] / 1e6, 2), "e2e ms", round(d["e2e"]["ms_per_step"], 4),
              "strict", round(d["e2e"]["strict"]["ms_per_step"], 4), "cpu ms", round(d.get("cpu_baseline", {}).get("ms_per_step", 0), 2),
              "pm fwd/bwd frac", round(d["roofline"]["frac"], 3), round(d["roofline_backward"]["frac"], 3),
              "gemm total", d.get("roofline_step", {}).get("gemm_total", {}).get("us"))
    except Exception as e:
        print(f, "failed", e)
PY
timeout -s KILL 300 python scripts/step_breakdown.py cfg2 > gpurun_out/step_breakdown_cfg2.log 2>&1; tail -16 gpurun_out/step_breakdown_cfg2.log
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/prof_step.py cfg2 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_step python scripts/prof_step.py cfg2 > gpurun_out/ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:pm_ -s 6 -c 2 -f -o gpurun_out/prof_pm python scripts/prof_driver.py pm "h2,s2,e2" > gpurun_out/ncu_pm.log 2>&1; echo "ncu pm rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:pm_ -s 6 -c 2 -f -o gpurun_out/prof_pm_cfg3 python scripts/prof_driver.py pm "h6,h6,s6,s6,e6" > gpurun_out/ncu_pm3.log 2>&1; echo "ncu pm cfg3 rc=$?"
ls -la gpurun_out | tail -20
