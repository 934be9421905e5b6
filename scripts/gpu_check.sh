#!/bin/bash
# Run on the GPU box through gpurun: parity tests under their own timeouts + diagnostics into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/gpu.txt 2>&1
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider "$@" > gpurun_out/test_gpu.log 2>&1
echo "tests rc=$?" | tee gpurun_out/summary.txt
tail -60 gpurun_out/test_gpu.log
