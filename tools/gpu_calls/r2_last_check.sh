#!/bin/bash
# last GPU call of round 2: the committed tree once more — tests, smoke, the default bench line
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2z_pytest_gpu.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/r2z_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/r2z_smoke.log
timeout 900 python bench.py > gpurun_out/r2z_bench_default.json.log 2>&1
echo done
