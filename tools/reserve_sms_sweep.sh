# 2 GPUs: exposed all-reduce time of the generator backward when the compute grids leave SMs to NCCL
run() { echo "== $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT tools/ddp_check.py 2>&1 | grep "^3\."; PORT=$((PORT+1)); }
PORT=29700
run A=1
run JPDSE_RESERVE_SMS=8 NCCL_MAX_CTAS=8
run JPDSE_RESERVE_SMS=16 NCCL_MAX_CTAS=16
run JPDSE_RESERVE_SMS=16
run JPDSE_RESERVE_SMS=24 NCCL_MAX_CTAS=24
