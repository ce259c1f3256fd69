set -x
timeout 900 python -m pytest tests/test_gpu_api.py -m gpu -q -x -k "host_buffer" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e16.json 2> gpurun_out/bench_e2e16.err; tail -3 gpurun_out/bench_e2e16.err
python -c "
import json
d=json.load(open('gpurun_out/bench_e2e16.json'))
print(d['e2e'])
"
