n=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n$n.log 2> gpurun_out/scale_n$n.err
tail -1 gpurun_out/scale_n$n.log | cut -c1-300
