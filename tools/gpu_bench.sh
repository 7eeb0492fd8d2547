#!/bin/bash
# forward parity + first bench line + clocks; everything lands in gpurun_out/
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py -m gpu -q -s --timeout 600 > gpurun_out/fwdcheck.log 2>&1
grep -E "rel-L2|passed|failed" gpurun_out/fwdcheck.log
timeout 900 python bench.py --steps ${STEPS:-20} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
cat gpurun_out/bench.json
