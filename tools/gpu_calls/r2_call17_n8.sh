#!/bin/bash
# GPU call 17 (8 GPUs): the scaling run — N = 8 with NCCL and with the peer-memory kernel, N = 1 on the same box
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="--steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29617"
timeout 600 python bench.py $B > gpurun_out/r2c17_n1.json.log 2>&1
MB200_DP_EXCHANGE=nccl timeout 900 $TR bench.py --gpus 8 $B > gpurun_out/r2c17_n8_nccl.json.log 2>&1
MB200_DP_EXCHANGE=peer timeout 900 $TR bench.py --gpus 8 $B > gpurun_out/r2c17_n8_peer.json.log 2>&1
MB200_DP_EXCHANGE=nccl MB200_DP_TRACE=1 timeout 900 $TR bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2c17_n8_trace_nccl.log 2>&1
echo done
