#!/bin/bash
# GPU call 4: short-key attention kernel, pair sentinel report, ncu source-level capture of the epilogue-bound GEMM
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/gemm_check.py --group pair > gpurun_out/r2c4_pair.log 2>&1
timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q -p no:cacheprovider -rA --timeout 300 > gpurun_out/r2c4_attn_tests.log 2>&1
echo "attn tests exit $?" | tee -a gpurun_out/r2c4_attn_tests.log
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider --timeout 600 -k "vit or golden or group" > gpurun_out/r2c4_model_tests.log 2>&1
echo "model tests exit $?" | tee -a gpurun_out/r2c4_model_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2c4_bench_n1.json.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm2_tcgen05 -s 4 -c 2 -o gpurun_out/r2c4_epi \
  python tools/ncu_epi.py > gpurun_out/r2c4_ncu_epi.log 2>&1
ls -la gpurun_out/*.ncu-rep >> gpurun_out/r2c4_ncu_epi.log 2>&1
echo done
