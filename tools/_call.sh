set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -3 gpurun_out/bench_c4.err
python -c "
import json
d=json.load(open('gpurun_out/bench_c4.json'))
print('step', d['ms_per_step'], 'attn', d['attn_ms'], 'prep', d['prep_ms'], d['roofline'], d['e2e'])
"
