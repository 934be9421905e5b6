#!/bin/bash
# usage: gpu_scale.sh N  (run under gpurun --gpus N): weak-scaling bench at N GPUs, peer-memory and NCCL collectives
N=$1
mkdir -p gpurun_out
for m in p2p nccl; do
  MVAE_DP=$m timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 500 --warmup 20 --skip-roofline 2>gpurun_out/scale_${N}_$m.err | tail -1 > gpurun_out/scale_${N}_$m.json
  python -c "import json,sys; d=json.load(open('gpurun_out/scale_${N}_$m.json')); print('N=$N', '$m', 'ms/step', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'value', round(d['value']))" || tail -5 gpurun_out/scale_${N}_$m.err
done
