#!/bin/bash
# GPU call 9: default kernel without the stream-K paths (spill fix), sigmoid-form GELU, optimizer on its own stream
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c9_gpu_tests.log 2>&1
echo "gpu tests exit $?"
timeout 600 python tools/epi_bench.py --only block,adapter,vit > gpurun_out/r2c9_epi_bench.log 2>&1
timeout 300 python tools/gemm_check.py --group streamk > gpurun_out/r2c9_streamk.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > gpurun_out/r2c9_bench_n1.json.log 2>&1
MB200_PIPELINE_OPT=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > gpurun_out/r2c9_bench_n1_opt_instream.json.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > gpurun_out/r2c9_bench_n1_b.json.log 2>&1
echo done
