#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_determinism.py > gpurun_out/determinism.log 2>&1
cat gpurun_out/determinism.log | tail -20
MMD_NO_GRAPH=1 timeout 600 python tools/gpu_determinism.py 2>&1 | tail -3 > gpurun_out/determinism_nograph.log
cat gpurun_out/determinism_nograph.log
timeout 900 python -m pytest tests/test_forward_gpu.py -m gpu -q -s --timeout 600 > gpurun_out/fwdcheck.log 2>&1
grep -E "rel-L2|passed|failed" gpurun_out/fwdcheck.log
