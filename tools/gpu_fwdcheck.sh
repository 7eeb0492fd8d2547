#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py -m gpu -q -s --timeout 600 -x 2>&1 | tail -40 > gpurun_out/fwdcheck.log
cat gpurun_out/fwdcheck.log
