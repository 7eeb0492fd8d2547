#!/bin/bash
# ncu captures of one production forward (B=4, un-graphed).  $1 = kernel regex, $2 = skip, $3 = count, $4 = tag
mkdir -p gpurun_out
K=${1:-conv_gemm}; S=${2:-0}; C=${3:-20}; TAG=${4:-cap}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -o gpurun_out/$TAG -f python tools/gpu_ncu_forward.py 4 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active > gpurun_out/${TAG}_summary.csv 2>/dev/null
wc -l gpurun_out/${TAG}_summary.csv
