set -x
O=gpurun_out/r02; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > $O/topo_${N}gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -q -k "two_devices" > $O/pytest_two_devices.txt 2>&1; tail -2 $O/pytest_two_devices.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/nccl_gather_check.py > $O/nccl_gather_${N}gpu.txt 2>&1; tail -5 $O/nccl_gather_${N}gpu.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_c4_${N}gpu.json 2> $O/bench_c4_${N}gpu.err
tail -3 $O/bench_c4_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('$O/bench_c4_${N}gpu.json') if l.startswith('{')][-1])
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'attn', d['attn_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'roof', d['e2e']['frac_of_copy_roof'], 'c5', d.get('scaling_c5'))
PY
