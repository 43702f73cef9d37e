#!/bin/bash
# GPU call 6 (2 GPUs): hardware DP parity test + data-parallel knob sweep
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q -p no:cacheprovider -rA -s --timeout 600 > gpurun_out/r2c6_dp_parity.log 2>&1
echo "dp parity exit $?" | tee -a gpurun_out/r2c6_dp_parity.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n1', round(d['value'],1), round(d['ms_per_step'],2))" > gpurun_out/r2c6_n2_sweep.log 2>&1
bash tools/gpu_calls/n2_sweep.sh >> gpurun_out/r2c6_n2_sweep.log 2>&1
echo done
