#!/bin/bash
# Backward-kernel parity on one B200, one process per kernel family (a trapped kernel only takes its own group down).
mkdir -p gpurun_out
O=gpurun_out
run() {  # name, pytest args...
  local name=$1; shift
  timeout ${T:-240} python -m pytest "$@" -m gpu -q -s --timeout 200 -p no:cacheprovider > $O/bwd_$name.log 2>&1
  echo "== $name rc=$?"; grep -E "^\[bwd|passed|failed|error|Error|mmd:|FAILED|timeout" $O/bwd_$name.log | cut -c1-260 | head -${LINES_MAX:-40}
}
run wgrad tests/test_ops_bwd_gpu.py -k "wgrad"
run dgrad tests/test_ops_bwd_gpu.py -k "dgrad"
run gn tests/test_ops_bwd_gpu.py -k "group_norm or resample"
run tattn tests/test_ops_bwd_gpu.py -k "temporal_attention"
run attn tests/test_ops_bwd_gpu.py -k "test_attention_bwd"
run head tests/test_ops_bwd_gpu.py -k "head"
run model tests/test_backward_gpu.py
if [ -n "$WITH_FWD" ]; then
  run fwd tests/test_ops_gpu.py tests/test_forward_gpu.py tests/test_diffusion_gpu.py
fi
