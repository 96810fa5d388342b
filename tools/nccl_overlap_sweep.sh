run() { echo "== $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT tools/ddp_check.py 2>&1 | grep "^3\."; PORT=$((PORT+1)); }
PORT=29600
run A=1
run NCCL_MAX_CTAS=4
run NCCL_MAX_CTAS=8
run NCCL_MAX_CTAS=16
run JPDSE_BUCKET_MB=128
run JPDSE_BUCKET_MB=256 NCCL_MAX_CTAS=8
run JPDSE_BUCKET_MB=16
