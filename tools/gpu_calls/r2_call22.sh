#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python tools/epi_bench.py --only block > gpurun_out/r2c22_epi_bench.log 2>&1
echo done
