#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget was spent, in one gpurun.
#   gpurun --timeout 1500 -- 'bash tools/gpu_calls/r2_first_gpu_call.sh'
# Writes gpurun_out/r2_*.log. Nothing here changes defaults; the knobs are environment variables.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
# 1. the unverified tests, with real outcomes (xfail marks off) ------------------------------------------------------
timeout 900 python -m pytest tests/test_zz_unverified_gpu.py -m gpu -q -p no:cacheprovider --runxfail -rA --timeout 300 \
  > gpurun_out/r2_unverified_tests.log 2>&1
echo "unverified tests exit $?" | tee -a gpurun_out/r2_unverified_tests.log
# 2. the verified suite (must stay green) -----------------------------------------------------------------------------
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_zz_unverified_gpu.py \
  > gpurun_out/r2_verified_tests.log 2>&1
echo "verified tests exit $?" | tee -a gpurun_out/r2_verified_tests.log
# 3. baseline bench line + launch list + clocks ----------------------------------------------------------------------
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --profile-from-start off --csv --log-file gpurun_out/r2_launches_step.csv python tools/profile_step.py \
  > gpurun_out/r2_profile_step.log 2>&1
# 3b. the same bench line with the LM on the general host-only schedule (candidate replacement of engine.cu's host code)
MB200_FORCE_GENERAL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline \
  > gpurun_out/r2_bench_n1_general_schedule.json.log 2>&1
# 4. trainable-encoder step time (freeze_img_encoder: false) ----------------------------------------------------------
nvidia-smi > gpurun_out/r2_nvsmi.log 2>&1
echo done

# A second call, on 2 GPUs, for the data-parallel knobs (each line: samples/s at N = 2):
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_calls/n2_sweep.sh'
