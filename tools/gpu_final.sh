#!/bin/bash
# Round-end measurement on one B200: tests, smoke, the bench line, ncu launch list of the bench command and --set full
# captures of the dominant kernels.  Everything lands in gpurun_out/ (kept under 64 MiB).
#   FULL_FWD=1  also re-capture the forward kernels (3x3 conv, qkv GEMM, attention, GroupNorm apply) and the per-launch
#               DRAM traffic of one forward;  WITH_REFERENCE_ARM=1 also times `bench.py --impl reference`.
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -5 > $O/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> $O/final_tests.log
cat $O/final_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_final.json 2> $O/bench_final.err; tail -2 $O/bench_final.err
if [ -n "$WITH_REFERENCE_ARM" ]; then
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -2 $O/bench_reference.err
fi
# (1) launch list of the bench command itself (driver contract): per-launch device time, cold-cache + serialised
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4200 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-reps 1 > $O/launches_bench.log 2>&1
gzip -f $O/launches_bench.csv
# (2) --set full captures of the backward kernels (one launch each) from one training step at B = 2
for spec in "conv_wgrad_kernel 6 wgrad" "attn_bwd_kernel 3 attn_bwd" "gn_bwd_apply_kernel 2 gn_bwd_apply"; do
  set -- $spec
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$1 -s $2 -c 1 -o $O/full_$3 -f python tools/gpu_ncu_train.py 2 > $O/ncu_full_$3.log 2>&1
  python tools/ncu_summarize.py $O/full_$3.ncu-rep $O/full_$3_summary.txt > /dev/null 2>&1
  find $O -name "full_$3.ncu-rep" -size +14M -delete
done
if [ -n "$FULL_FWD" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file $O/launches_forward.csv python tools/gpu_ncu_forward.py 4 > $O/launches_forward.log 2>&1
  gzip -f $O/launches_forward.csv
  for spec in "conv_gemm_kernel 3 conv3x3_l0" "conv_gemm_kernel 23 qkv_bn256" "attention64_kernel 0 attn64_self" "gn_apply_kernel 0 gn_apply_l0"; do
    set -- $spec
    timeout 600 ncu --set full --import-source on --clock-control none -k regex:$1 -s $2 -c 1 -o $O/full_$3 -f python tools/gpu_ncu_forward.py 4 > $O/ncu_full_$3.log 2>&1
    python tools/ncu_summarize.py $O/full_$3.ncu-rep $O/full_$3_summary.txt > /dev/null 2>&1
    find $O -name "full_$3.ncu-rep" -size +14M -delete
  done
fi
ls -la $O | grep -E "full_|launches_|bench_final"
du -sh $O
