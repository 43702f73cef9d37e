#!/bin/bash
# 2-GPU data-parallel sweep (gpurun --gpus 2): the SM carve-out window, bf16 exchange, NCCL CTA cap, bucket count.
# Each line: tag, samples/s (device-resident), ms/step, e2e samples/s.
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d.get('allreduce'))" ; }
run default X=1
run gemm140 MB200_DP_GEMM_SMS=140
run gemm132 MB200_DP_GEMM_SMS=132
run bf16 MB200_DP_BF16=1
run gemm140_bf16 MB200_DP_GEMM_SMS=140 MB200_DP_BF16=1
run gemm140_cta8 MB200_DP_GEMM_SMS=140 NCCL_MAX_CTAS=8
run gemm140_cta8_bf16 MB200_DP_GEMM_SMS=140 NCCL_MAX_CTAS=8 MB200_DP_BF16=1
run gemm132_cta16 MB200_DP_GEMM_SMS=132 NCCL_MAX_CTAS=16
run gemm140_b7 MB200_DP_GEMM_SMS=140 MB200_DP_BUCKETS=7
run gemm140_b2 MB200_DP_GEMM_SMS=140 MB200_DP_BUCKETS=2
