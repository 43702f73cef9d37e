#!/bin/bash
# GPU call 2 of round 2: full GPU suite on the consolidated runtime (+ flash attention, full-size parity), the epilogue /
# fixed-cost microbenchmark, the bench line, decode baseline.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -rA --timeout 600 > gpurun_out/r2c2_gpu_tests.log 2>&1
echo "gpu tests exit $?" | tee -a gpurun_out/r2c2_gpu_tests.log
timeout 600 python tools/epi_bench.py > gpurun_out/r2c2_epi_bench.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_bench_n1.json.log 2>&1
timeout 600 python tools/decode_bench.py > gpurun_out/r2c2_decode_bench.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c2_smoke.log 2>&1
echo done
