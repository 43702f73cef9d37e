#!/bin/bash
# GPU call 12: optimizer grid capped at 2 blocks / SM beside the GEMMs (pipelined), grid-independent sumsq
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_training_paths_gpu.py -m gpu -x -q -k "sumsq or own_stream or ops_group or checkpoint_resume" > gpurun_out/r2c12_tests.log 2>&1
echo "tests exit $?"
B="--steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline"
timeout 900 python bench.py $B > gpurun_out/r2c12_bench_n1.json.log 2>&1
MB200_PIPELINE_OPT=0 timeout 900 python bench.py $B > gpurun_out/r2c12_bench_n1_opt_instream.json.log 2>&1
MB200_OPT_BLOCKS_PER_SM=1 timeout 900 python bench.py $B > gpurun_out/r2c12_bench_n1_1blk.json.log 2>&1
MB200_OPT_BLOCKS_PER_SM=16 timeout 900 python bench.py $B > gpurun_out/r2c12_bench_n1_16blk.json.log 2>&1
timeout 900 python bench.py $B > gpurun_out/r2c12_bench_n1_b.json.log 2>&1
echo done
