V=quantizedmha_b200/lib/variants
timeout 900 python tools/ab_libs.py base=$V/libqmha_base.so lean=$V/libqmha_lean.so --rounds 4 --reps 40 > gpurun_out/ab_lean.log 2>&1
tail -3 gpurun_out/ab_lean.log
