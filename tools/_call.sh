set -x
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_extensions.py -m gpu -q -x > gpurun_out/r2_pytest5.log 2>&1
tail -5 gpurun_out/r2_pytest5.log
for mode in stream cluster; do
  if [ $mode = cluster ]; then export QMHA_CLUSTER_QUANT=1; fi
  timeout 600 python bench.py --workload c4 --scales head --steps 10 --warmup 3 --no-cpu-baseline --no-signed --e2e-steps 1 > gpurun_out/bench_head_$mode.json 2> gpurun_out/bench_head_$mode.err; tail -2 gpurun_out/bench_head_$mode.err
  python -c "
import json
d=json.load(open('gpurun_out/bench_head_$mode.json'))
print('$mode', 'prep', d['prep_ms'], d['prep']['frac_algorithmic'], 'attn', d['attn_ms'], d['parity']['ok'])
"
done
unset QMHA_CLUSTER_QUANT
timeout 600 python bench.py --workload c4 --scales tensor --steps 10 --warmup 3 --no-cpu-baseline --no-signed --e2e-steps 1 > gpurun_out/bench_tensor.json 2> gpurun_out/bench_tensor.err; tail -2 gpurun_out/bench_tensor.err
python -c "
import json
d=json.load(open('gpurun_out/bench_tensor.json'))
print('tensor', 'prep', d['prep_ms'], d['prep']['frac_algorithmic'], 'attn', d['attn_ms'], d['parity']['ok'])
"
