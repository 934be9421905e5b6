#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/test_gpu.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 7 -f -o gpurun_out/prof_gemm python scripts/prof_driver.py gemm > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
timeout -s KILL 300 python scripts/step_breakdown.py cfg2 2>&1 | tee gpurun_out/step_breakdown_cfg2.log
