set -x
O=gpurun_out/r02; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/nccl_gather_check.py > $O/nccl_gather_${N}gpu.txt 2>&1; tail -8 $O/nccl_gather_${N}gpu.txt
