#!/bin/bash
# GPU call 18: decode — frozen-weight tiles requested ahead of the PDL wait (1-CTA GEMM) and the MLP chain of a layer on
# a second stream beside the attention chain; A/B of both switches
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python tools/gemm_check.py > gpurun_out/r2c18_gemm_check.log 2>&1
echo "gemm_check exit $?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c18_gpu_tests.log 2>&1
echo "tests exit $?"
D="--workload decode --steps 3 --warmup 1"
timeout 900 python bench.py $D > gpurun_out/r2c18_bench_decode.json.log 2>&1
MB200_SIDE_STREAM=0 timeout 900 python bench.py $D > gpurun_out/r2c18_bench_decode_onechain.json.log 2>&1
MB200_B_PRELOAD=0 timeout 900 python bench.py $D > gpurun_out/r2c18_bench_decode_nopreload.json.log 2>&1
MB200_SIDE_STREAM=0 MB200_B_PRELOAD=0 timeout 900 python bench.py $D > gpurun_out/r2c18_bench_decode_neither.json.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > gpurun_out/r2c18_bench_n1.json.log 2>&1
echo done
