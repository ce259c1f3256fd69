# Round-2 evidence run on one B200 (gpurun): GPU test suite, bench lines of every BASELINE configuration, the ncu
# launch list of the default bench command and full ncu captures of the shipped kernels, summarised on the box
# (the reports are too large to pull).  Outputs under gpurun_out/r02/.
set -x
O=gpurun_out/r02; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
python bench.py > $O/bench_c4_1gpu.json 2> $O/bench_c4_1gpu.err; tail -c 300 $O/bench_c4_1gpu.json
for w in c2 c3 c5 c4f16 c4bf16; do
  python bench.py --workload $w --steps 10 --warmup 3 > $O/bench_${w}_1gpu.json 2> $O/bench_${w}_1gpu.err
done
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2>&1
# launch list of the same command (only after it ran clean without ncu)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-signed --e2e-steps 1 > $O/b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_c4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-signed --e2e-steps 1 > $O/ncu_launches.log 2>&1
python tools/launch_shares.py $O/launches_bench_c4.csv "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-signed --e2e-steps 1" > $O/launch_shares_c4.txt; cat $O/launch_shares_c4.txt
# full captures, one kernel each
cap() {  # name, kernel regex, prof_one args
  python tools/prof_one.py $3 $4 1 $5 > $O/prof_plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -f -o /tmp/$1 python tools/prof_one.py $3 $4 1 $5 > $O/ncu_$1.log 2>&1 && \
  python tools/ncu_hot.py /tmp/$1.ncu-rep 25 > $O/ncu_$1.txt 2>&1
  rm -f /tmp/$1.ncu-rep
}
cap attn_fwd_int8_block_c4 attn_fwd int8 8,32,8192,128 block
cap block_quantize_c4 block_quantize int8 8,32,8192,128 block
cap attn_fwd_f16_c4 attn_fwd f16 8,32,8192,128 head
cap attn_fwd_bf16_c4 attn_fwd bf16 8,32,8192,128 head
cap fused_quantize_c4 fused_quantize int8 8,32,8192,128 head
grep -h "gpu__time_duration\|dram__bytes\|pipe_tensor\|pipe_xu" $O/ncu_*.txt
