# End-of-session check on the GPU box: full GPU test suite, default bench line, ncu launch list of the
# same bench command and one full ncu capture of the block quantiser.  Outputs under gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/final_pytest.log
python bench.py > gpurun_out/final_bench.log 2> gpurun_out/final_bench.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_launches.log 2>&1
python tools/prof_one.py int8 8,32,8192,128 1 block > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:block_quantize -c 1 -f -o gpurun_out/prof_blockq_final python tools/prof_one.py int8 8,32,8192,128 1 block > gpurun_out/prof_blockq_ncu.log 2>&1
cat gpurun_out/final_pytest.log; tail -c 400 gpurun_out/final_bench.log
