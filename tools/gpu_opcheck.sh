#!/bin/bash
# Run the operator parity tests on the GPU box, one process per op family (a trapping kernel
# poisons its CUDA context), with hang guards; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
: > gpurun_out/opcheck.log
for k in ${OPS:-conv_pointwise conv_spatial conv_temporal conv_audio conv_heads group_norm resample test_attention temporal_attention}; do
  echo "=== $k" >> gpurun_out/opcheck.log
  timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 120 -k "$k" 2>&1 | tail -${TAILN:-25} >> gpurun_out/opcheck.log
done
grep -E "^===|passed|failed|error|Error|timeout|assert " gpurun_out/opcheck.log | head -150
