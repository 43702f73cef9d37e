#!/bin/bash
# GPU call 7: stream-K of the last wave
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python tools/gemm_check.py --group streamk > gpurun_out/r2c7_streamk.log 2>&1
echo "streamk exit $?" | tee -a gpurun_out/r2c7_streamk.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 -x > gpurun_out/r2c7_gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/r2c7_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2c7_bench_n1.json.log 2>&1
MB200_STREAMK=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2c7_bench_n1_nostreamk.json.log 2>&1
echo done
