run() { tag=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag', d['value'], d['ms_per_step'], d['e2e']['value'])" ; }
run default X=1
run ch8 NCCL_MAX_NCHANNELS=8
run ch4 NCCL_MAX_NCHANNELS=4
run ch8_b7 NCCL_MAX_NCHANNELS=8 MB200_DP_BUCKETS=7
run b7 MB200_DP_BUCKETS=7
