#!/bin/bash
# GPU call 5: N-edge fix of the TMA-store epilogue, input prefetch depth 2, graph-replayed decode loop
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/gemm_check.py --group pair > gpurun_out/r2c5_pair.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -rA --timeout 600 > gpurun_out/r2c5_gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/r2c5_gpu_tests.log
timeout 600 python tools/epi_bench.py --only sweep,block,adapter > gpurun_out/r2c5_epi_bench.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2c5_bench_n1.json.log 2>&1
timeout 900 python bench.py --workload decode --steps 3 --warmup 1 > gpurun_out/r2c5_bench_decode.json.log 2>&1
MB200_DECODE_GRAPH=0 timeout 900 python bench.py --workload decode --steps 2 --warmup 1 > gpurun_out/r2c5_bench_decode_hostloop.json.log 2>&1
echo done
