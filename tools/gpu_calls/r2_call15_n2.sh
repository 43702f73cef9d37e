#!/bin/bash
# GPU call 15 (2 GPUs): event timeline of the data-parallel step (peer kernel and NCCL), then bench lines
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="--steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29615"
MB200_DP_TRACE=1 timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c15_trace_peer.log 2>&1
MB200_DP_TRACE=1 MB200_DP_EXCHANGE=nccl timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c15_trace_nccl.log 2>&1
MB200_DP_TRACE=1 MB200_DP_DIAG_NO_EXCHANGE=1 timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c15_trace_noexchange.log 2>&1
B="--steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager"
timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c15_n2_peer.json.log 2>&1
MB200_DP_EXCHANGE=nccl timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c15_n2_nccl.json.log 2>&1
echo done
