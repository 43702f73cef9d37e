#!/bin/bash
# GPU call 8: default tree after the stream-K A/B (off by default): adapter scratch A/B, train / conv / decode bench lines,
# launch list of the final step, ncu --set full of the main GEMM shapes and of the attention kernels
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/gemm_check.py --group streamk > gpurun_out/r2c8_streamk.log 2>&1
timeout 600 python tools/epi_bench.py --only adapter,vit > gpurun_out/r2c8_epi_bench.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c8_bench_n1.json.log 2>&1
timeout 900 python bench.py --workload conv --steps 10 --warmup 3 > gpurun_out/r2c8_bench_conv.json.log 2>&1
timeout 900 python bench.py --workload decode --steps 3 --warmup 1 > gpurun_out/r2c8_bench_decode.json.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --profile-from-start off --csv --log-file gpurun_out/r2c8_launches_step.csv python tools/profile_step.py \
  > gpurun_out/r2c8_profile_step.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm2_tcgen05 \
  -o gpurun_out/r2c8_gemm_full python tools/profile_step.py --gemm-only > gpurun_out/r2c8_ncu_gemm.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --profile-from-start off --csv --log-file gpurun_out/r2c8_launches_decode.csv python tools/profile_step.py --decode \
  > gpurun_out/r2c8_profile_decode.log 2>&1
echo done
