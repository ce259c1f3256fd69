V=quantizedmha_b200/lib/variants
timeout 900 python tools/ab_libs.py base=$V/libqmha_base.so nokv=$V/libqmha_nokv.so --rounds 3 --reps 60 > gpurun_out/ab_nokv.log 2>&1
tail -4 gpurun_out/ab_nokv.log
