V=quantizedmha_b200/lib/variants
timeout 900 python tools/ab_libs.py base=$V/libqmha_base.so ps1=$V/libqmha_ps1.so ps1a=$V/libqmha_ps1a.so ps0=$V/libqmha_ps0.so --rounds 3 --reps 40 > gpurun_out/ab_ps.log 2>&1
tail -5 gpurun_out/ab_ps.log
