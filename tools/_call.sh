# compute-sanitizer memcheck over the shipped kernels on small shapes (ragged N, every head dim, every kernel)
O=gpurun_out/sanitize.txt; : > $O
for args in "--kernel=fa_tc_int8_b --N=300 --d_model=256 --h=2" "--kernel=fa_tc_int8_b --N=520 --d_model=128 --h=2 --random" "--kernel=fa_tc_int8_b --N=257 --d_model=64 --h=2 --random --rope" "--kernel=fa_tc_v2a --N=300 --d_model=256 --h=2 --random" "--kernel=fa_b200_bf16 --N=130 --d_model=128 --h=2 --random" "--kernel=int8_pv8 --N=300 --d_model=256 --h=2 --random"; do
  echo "== compute-sanitizer --tool memcheck bin/profile_fa_tc_int8_b $args --warmup=0 --runs=1" >> $O
  timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 bin/profile_fa_tc_int8_b $args --warmup=0 --runs=1 2>&1 | grep -v "^=========\s*$" | tail -6 >> $O
  echo "exit ${PIPESTATUS[0]}" >> $O
done
cat $O
