#!/bin/bash
# Lean ncu capture of a few launches of one production forward (B=4, un-graphed).
#   $1 kernel regex   $2 launch-skip   $3 launch-count   $4 tag   [$5 = "full" for --set full, default: a 5-section set]
# Keeps gpurun_out small: exports the raw metric page as CSV next to the report and drops reports over 24 MB.
mkdir -p gpurun_out
K=${1:-conv_gemm}; S=${2:-0}; C=${3:-4}; TAG=${4:-cap}; MODE=${5:-lean}
if [ "$MODE" = "full" ]; then SECT="--set full --import-source on"; else
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy"; fi
timeout 900 ncu $SECT --clock-control none -k regex:$K -s $S -c $C -o gpurun_out/$TAG -f python tools/gpu_ncu_forward.py 4 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/$TAG.ncu-rep gpurun_out/${TAG}_raw.csv
find gpurun_out -name "$TAG.ncu-rep" -size +24M -delete
