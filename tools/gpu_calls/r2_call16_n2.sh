#!/bin/bash
# GPU call 16 (2 GPUs): in-place peer-memory exchange — parity, timeline, bench vs NCCL
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q -p no:cacheprovider -rA -s --timeout 600 > gpurun_out/r2c16_dp_parity.log 2>&1
echo "dp parity exit $?" | tee -a gpurun_out/r2c16_dp_parity.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29616"
MB200_DP_TRACE=1 timeout 900 $TR bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2c16_trace_peer.log 2>&1
B="--steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager"
timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c16_n2_peer.json.log 2>&1
MB200_DP_EXCHANGE=nccl timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c16_n2_nccl.json.log 2>&1
MB200_OPT_BLOCKS_PER_SM=16 timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c16_n2_peer_opt16.json.log 2>&1
MB200_DP_PEER_BLOCKS=64 timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c16_n2_peer_64blk.json.log 2>&1
echo done
