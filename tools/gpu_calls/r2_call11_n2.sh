#!/bin/bash
# GPU call 11 (2 GPUs): DP parity rerun on the final tree, N = 1 on each GPU of the box, N = 2 with / without the
# gradient exchange (diagnostic), N = 2 with the optimizer in-stream
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q -p no:cacheprovider -rA -s --timeout 600 > gpurun_out/r2c11_dp_parity.log 2>&1
echo "dp parity exit $?" | tee -a gpurun_out/r2c11_dp_parity.log
B="--steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py $B > gpurun_out/r2c11_n1_gpu0.json.log 2>&1
CUDA_VISIBLE_DEVICES=1 timeout 600 python bench.py $B > gpurun_out/r2c11_n1_gpu1.json.log 2>&1
timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c11_n2.json.log 2>&1
MB200_DP_DIAG_NO_EXCHANGE=1 timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c11_n2_noexchange.json.log 2>&1
MB200_PIPELINE_OPT=0 timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c11_n2_opt_instream.json.log 2>&1
timeout 900 $TR bench.py --gpus 2 $B > gpurun_out/r2c11_n2_b.json.log 2>&1
echo done
