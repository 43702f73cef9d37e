run() { tag=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag', d['value'], d['ms_per_step'], d['e2e']['value'])" ; }
run default X=1
run ch8 NCCL_MAX_NCHANNELS=8
run ch4 NCCL_MAX_NCHANNELS=4
run ch8_b7 NCCL_MAX_NCHANNELS=8 MB200_DP_BUCKETS=7
run b7 MB200_DP_BUCKETS=7
# knobs added after the round-1 GPU budget was spent: keep SMs free for NCCL, bf16 gradient exchange
run gemm140 MB200_DP_GEMM_SMS=140
run gemm132 MB200_DP_GEMM_SMS=132
run bf16 MB200_DP_BF16=1
run gemm140_bf16 MB200_DP_GEMM_SMS=140 MB200_DP_BF16=1
run gemm140_b7 MB200_DP_GEMM_SMS=140 MB200_DP_BUCKETS=7
