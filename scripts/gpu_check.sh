#!/bin/bash
# Run on the GPU box through gpurun: kernel parity tests first, then the GEMM probe + GEMM tests under their own
# timeouts so that a hung tcgen05 pipeline cannot eat the whole call.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/gpu.txt 2>&1
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not gemm" -p no:cacheprovider > gpurun_out/test_kernels.log 2>&1
echo "kernels rc=$?" | tee -a gpurun_out/summary.txt
timeout -s KILL 300 python scripts/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1
echo "probe rc=$?" | tee -a gpurun_out/summary.txt
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "gemm" -p no:cacheprovider > gpurun_out/test_gemm.log 2>&1
echo "gemm rc=$?" | tee -a gpurun_out/summary.txt
tail -30 gpurun_out/test_kernels.log
tail -40 gpurun_out/gemm_probe.log
tail -30 gpurun_out/test_gemm.log
