set -x
timeout 300 python tools/quant_ab.py block=quantizedmha_b200/lib/libqmha.so --gran block
timeout 600 python bench.py --workload c4pv8 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e16 --e2e-steps 1 > gpurun_out/bench_c4pv8.json 2> gpurun_out/bench_c4pv8.err; tail -3 gpurun_out/bench_c4pv8.err
python -c "
import json
d=json.load(open('gpurun_out/bench_c4pv8.json'))
print('pv8', 'step', d['ms_per_step'], 'attn', d['attn_ms'], 'prep', d['prep_ms'], d['prep']['frac_algorithmic'])
"
timeout 900 python -m pytest tests/test_gpu_pv8.py -m gpu -q -x 2>&1 | tail -3
