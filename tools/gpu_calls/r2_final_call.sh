#!/bin/bash
# Final 1-GPU evidence call of round 2: tests, smoke, the three bench lines, launch lists, ncu --set full of the GEMM core
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2f_pytest_gpu.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/r2f_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/r2f_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_n1.json.log 2>&1
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2f_bench_reference.json.log 2>&1
timeout 900 python bench.py --workload conv --steps 10 --warmup 3 > gpurun_out/r2f_bench_conv.json.log 2>&1
timeout 900 python bench.py --workload decode --steps 3 --warmup 1 > gpurun_out/r2f_bench_decode.json.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --profile-from-start off --csv --log-file gpurun_out/r2f_launches_step.csv python tools/profile_step.py \
  > gpurun_out/r2f_profile_step.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm2_tcgen05 \
  -o gpurun_out/r2f_gemm_full python tools/profile_step.py --gemm-only > gpurun_out/r2f_ncu_gemm.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --profile-from-start off --csv --log-file gpurun_out/r2f_launches_decode.csv python tools/profile_step.py --decode \
  > gpurun_out/r2f_profile_decode.log 2>&1
echo done
