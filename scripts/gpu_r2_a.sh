#!/bin/bash
# Round 2, run A (one GPU): all GPU parity tests, a short bench line, compute-sanitizer over the graphed step.
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/test_gpu.log 2>&1; echo "tests rc=$?"
grep -E "^\[|passed|failed|FAILED|Error" gpurun_out/test_gpu.log | tail -40
timeout -s KILL 500 python bench.py --steps 300 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout -s KILL 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_step.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/sanitizer_memcheck.log
timeout -s KILL 600 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_step.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -5 gpurun_out/sanitizer_racecheck.log
