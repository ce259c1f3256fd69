#!/bin/bash
# Per-launch GPU durations of the reference's own kernels (sm_100 build of the unmodified sources,
# baseline/Makefile) on the B200, via the ncu launch list.  One solve() = 32 heads x (3 extract_mat
# + fa_kernel + concat_mat).  Output: gpurun_out/ref_gpu_launches_<kernel>.csv
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/qmha_refgpu && cd /tmp/qmha_refgpu
for k in fa_tc_int8_b fa_tc_v1b; do
  exe=/root/repo/baseline/_ref/profile_$k
  $exe --no-check --warmup=0 --runs=1 > /root/repo/gpurun_out/ref_gpu_plain_$k.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file /root/repo/gpurun_out/ref_gpu_launches_$k.csv $exe --no-check --warmup=0 --runs=1 \
      > /root/repo/gpurun_out/ref_gpu_ncu_$k.log 2>&1
  echo "$k: rc=$?"
done
