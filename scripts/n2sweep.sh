# quick N=2 sweep of NCCL settings against the bench (tuning aid)
run() { echo "== $1"; env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 10 --warmup 3 --no-cfg3-leg 2>/dev/null | grep '^{' | python -c "import json,sys;j=json.loads(sys.stdin.read());print(round(j['value']),round(j['ms_per_step'],2),round(j['e2e']['ms_per_step'],2))"; }
run "NCCL_MIN_NCHANNELS=16" 29531
run "NCCL_MIN_NCHANNELS=32" 29532
run "NCCL_NVLS_ENABLE=0" 29533
run "NCCL_PROTO=Simple" 29534
