set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/final_pytest.log
python bench.py > gpurun_out/final_bench.log 2> gpurun_out/final_bench.err
python bench.py --workload c4f16 --no-cpu-baseline > gpurun_out/final_bench_f16.log 2> gpurun_out/final_bench_f16.err
python tools/prof_one.py f16 8,32,8192,128 2 > gpurun_out/prof_f16_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 1 -f -o gpurun_out/prof_attn_f16 python tools/prof_one.py f16 8,32,8192,128 2 > gpurun_out/prof_f16_ncu.log 2>&1
cat gpurun_out/final_pytest.log; tail -c 600 gpurun_out/final_bench.log; tail -c 300 gpurun_out/final_bench_f16.log
