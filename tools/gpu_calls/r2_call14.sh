#!/bin/bash
# GPU call 14: co-resident kernels ask for the GEMM's shared-memory carve-out
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="--steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline"
timeout 900 python bench.py $B > gpurun_out/r2c14_bench_n1.json.log 2>&1
MB200_PIPELINE_OPT=0 timeout 900 python bench.py $B > gpurun_out/r2c14_bench_n1_opt_instream.json.log 2>&1
MB200_OPT_BLOCKS_PER_SM=4 timeout 900 python bench.py $B > gpurun_out/r2c14_bench_n1_4blk.json.log 2>&1
timeout 900 python bench.py $B > gpurun_out/r2c14_bench_n1_b.json.log 2>&1
echo done
