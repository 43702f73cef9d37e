#!/bin/bash
# GPU call 21: QuickGELU epilogue with the approximate divide (ViT c_fc GEMM)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python tools/epi_bench.py --only vit > gpurun_out/r2c21_epi_bench.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2c21_pytest_gpu.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/r2c21_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c21_bench_n1.json.log 2>&1
echo done
