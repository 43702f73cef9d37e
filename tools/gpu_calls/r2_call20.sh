#!/bin/bash
# GPU call 20: full-size parity logs (config 2 training, config 5 KV decode) + ncu sections of every non-GEMM kernel of a step
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_model_gpu.py -m gpu -q -s -k "full_size" > gpurun_out/r2c20_full_size.log 2>&1
echo "full-size tests exit $?" | tee -a gpurun_out/r2c20_full_size.log
timeout 900 ncu --profile-from-start off --clock-control none --section SpeedOfLight --section MemoryWorkloadAnalysis \
  --section LaunchStats --section Occupancy \
  -k regex:'attn_|layernorm_|adamw|sumsq|ce_row|colsum|splitk_finalize' -c 340 \
  -o /tmp/r2c20_nongemm python tools/profile_step.py > gpurun_out/r2c20_ncu_nongemm.log 2>&1
ncu -i /tmp/r2c20_nongemm.ncu-rep --page raw --csv > gpurun_out/r2c20_nongemm_raw.csv 2>/dev/null
ls -la /tmp/r2c20_nongemm.ncu-rep gpurun_out/ >> gpurun_out/r2c20_ncu_nongemm.log 2>&1
echo done
