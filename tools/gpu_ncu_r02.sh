#!/bin/bash
# Round-2 profiling evidence on one B200 (run under gpurun).  Everything lands in gpurun_out/ (kept small):
#   (1) launch list of the bench command (per-launch device time, ncu --metrics gpu__time_duration.sum)
#   (2) ncu --set full captures of the step families VERDICT names, selected through the NVTX range each plan step
#       carries on the eager path (MMD_NO_GRAPH=1): 1x1 GEMMs with the fused GroupNorm apply, attention, GroupNorm
#   (3) per-launch DRAM traffic of one forward (dram__bytes_read/write) for the roofline `traffic` field
# ENV: extra environment for the profiled process, e.g. ENV="MMD_ATTN_PAIR=1"
mkdir -p gpurun_out
O=gpurun_out
EXTRA_ENV=$(echo "$ENV" | tr ',' ' ')
timeout 900 env $EXTRA_ENV ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-extras --profile-reps 1 > $O/launches_bench.log 2>&1
gzip -f $O/launches_bench.csv
for spec in "conv1x1_qkv 4 qkv" "conv1x1_out 0 out_l0" "conv1x1_proj 2 proj" "cross_attention 0 xattn" "self_attention 0 sattn" \
            "group_norm 0 gn" "conv_temporal 0 tconv" "conv3x3_spatial 0 conv3x3"; do
  set -- $spec
  timeout 600 env $EXTRA_ENV ncu --set full --import-source on --clock-control none --nvtx --nvtx-include "$1/" -s $2 -c 1 \
      -o $O/full_$3 -f python tools/gpu_ncu_forward.py 4 > $O/ncu_full_$3.log 2>&1
  python tools/ncu_summarize.py $O/full_$3.ncu-rep $O/full_$3_summary.txt > /dev/null 2>&1
  ncu -i $O/full_$3.ncu-rep --page source --csv 2>/dev/null | gzip > $O/full_$3_src.csv.gz
  rm -f $O/full_$3.ncu-rep   # the merged gpurun_out/ is capped at 64 MiB: keep the summary and the source page only
done
timeout 900 env $EXTRA_ENV ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $O/launches_forward.csv python tools/gpu_ncu_forward.py 4 > $O/launches_forward.log 2>&1
gzip -f $O/launches_forward.csv
ls -la $O | grep -E "full_|launches_"
du -sh $O
