cd /root/repo; mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/dbg/dp_equiv_nccl.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -32
