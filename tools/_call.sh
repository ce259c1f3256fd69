set -x
V=quantizedmha_b200/lib/variants
timeout 600 python tools/ab_libs.py hard=$V/libqmha_hardrel.so soft=quantizedmha_b200/lib/libqmha.so --rounds 3 > gpurun_out/ab_soft.log 2>&1
tail -4 gpurun_out/ab_soft.log
timeout 600 python tools/ab_libs.py hard=$V/libqmha_hardrel.so soft=quantizedmha_b200/lib/libqmha.so --rounds 2 --kernel f16 --gran head > gpurun_out/ab_soft_f16.log 2>&1
tail -3 gpurun_out/ab_soft_f16.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1
tail -5 gpurun_out/r2_pytest2.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_soft.json 2> gpurun_out/bench_soft.err
tail -c 3000 gpurun_out/bench_soft.json; tail -5 gpurun_out/bench_soft.err
