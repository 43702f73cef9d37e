#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2c23_bench_n1.json.log 2>&1
echo done
