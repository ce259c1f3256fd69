for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n$n.log 2> gpurun_out/scale_n$n.err
  tail -1 gpurun_out/scale_n$n.log | cut -c1-400
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 --workload c5 > gpurun_out/scale_c5_n8.log 2> gpurun_out/scale_c5_n8.err
tail -1 gpurun_out/scale_c5_n8.log | cut -c1-400
