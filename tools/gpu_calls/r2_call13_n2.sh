#!/bin/bash
# GPU call 13 (2 GPUs): peer-memory gradient exchange — parity (peer / nccl / nccl-bf16), N = 2 bench with each
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q -p no:cacheprovider -rA -s --timeout 600 > gpurun_out/r2c13_dp_parity.log 2>&1
echo "dp parity exit $?" | tee -a gpurun_out/r2c13_dp_parity.log
B="--steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613"
timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c13_n2_peer.json.log 2>&1
MB200_DP_EXCHANGE=nccl timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c13_n2_nccl.json.log 2>&1
MB200_DP_PEER_BLOCKS=32 timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c13_n2_peer_32blk.json.log 2>&1
MB200_DP_DIAG_NO_EXCHANGE=1 timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c13_n2_noexchange.json.log 2>&1
timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c13_n2_peer_b.json.log 2>&1
echo done
