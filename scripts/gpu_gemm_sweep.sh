#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/gemm_sweep.log
for t in "0,0,0" "80,1,0" "96,1,0" "112,1,0" "112,2,0" "128,1,0" "160,1,0" "208,1,0" "256,1,0" "64,2,0" "48,2,0" "80,1,3" "80,1,4" "80,2,4" "128,1,4" "128,1,8" "208,1,8" "208,1,10" "256,1,8" "64,1,4" "64,2,6" "64,1,8"; do
  MVAE_GEMM_TUNE=$t timeout -s KILL 120 python scripts/gemm_sweep.py cfg2 2>&1 | grep "us" >> gpurun_out/gemm_sweep.log
done
cat gpurun_out/gemm_sweep.log
