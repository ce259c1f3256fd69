timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; python -c "
import json
d=json.load(open('gpurun_out/bench_final.json'))
print('step', d['ms_per_step'], 'attn', d['attn_ms'], 'prep', d['prep_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['clocks'], d['parity'])
"
