#!/bin/bash
# GPU call 3: epilogue v3 (TMA store) + weight preload before the PDL wait
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -p no:cacheprovider -rA --timeout 300 -x > gpurun_out/r2c3_gemm_tests.log 2>&1
echo "gemm tests exit $?" | tee -a gpurun_out/r2c3_gemm_tests.log
timeout 600 python tools/epi_bench.py > gpurun_out/r2c3_epi_bench.log 2>&1
MB200_B_PRELOAD=0 timeout 600 python tools/epi_bench.py --only sweep,block > gpurun_out/r2c3_epi_bench_nopreload.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -rA --timeout 600 --deselect tests/test_gemm_gpu.py > gpurun_out/r2c3_gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/r2c3_gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c3_bench_n1.json.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --profile-from-start off --csv --log-file gpurun_out/r2c3_launches_step.csv python tools/profile_step.py \
  > gpurun_out/r2c3_profile_step.log 2>&1
echo done
