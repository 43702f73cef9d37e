#!/bin/bash
# GPU call 10: pair kernel launched at 152 regs/thread + setmaxnreg (other kernels can be resident beside it)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python tools/gemm_check.py > gpurun_out/r2c10_gemm_check.log 2>&1
echo "gemm_check exit $?"
timeout 600 python tools/epi_bench.py --only block,adapter,vit > gpurun_out/r2c10_epi_bench.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > gpurun_out/r2c10_bench_n1.json.log 2>&1
MB200_PIPELINE_OPT=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > gpurun_out/r2c10_bench_n1_opt_instream.json.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > gpurun_out/r2c10_bench_n1_b.json.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c10_gpu_tests.log 2>&1
echo "gpu tests exit $?"
echo done
