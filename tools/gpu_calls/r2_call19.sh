#!/bin/bash
# GPU call 19: full-size parity (config 2 training, config 5 KV decode) + ncu --set full of the non-GEMM kernels of a step
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_model_gpu.py -m gpu -q -s -k "full_size" > gpurun_out/r2c19_full_size.log 2>&1
echo "full-size tests exit $?" | tee -a gpurun_out/r2c19_full_size.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:'attn_|layernorm_|adamw|sumsq|ce_row|colsum|splitk_finalize' -c 60 \
  -o gpurun_out/r2c19_nongemm_full python tools/profile_step.py > gpurun_out/r2c19_ncu_nongemm.log 2>&1
echo done
