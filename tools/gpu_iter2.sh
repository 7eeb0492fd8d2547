#!/bin/bash
# forward/diffusion parity + bench under both capture modes
mkdir -p gpurun_out
: > gpurun_out/iter.log
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_diffusion_gpu.py -m gpu -q -s --timeout 600 2>&1 | grep -E "rel-L2|passed|failed|Error|error" >> gpurun_out/iter.log
cat gpurun_out/iter.log
for mode in 1 0; do
  MMD_ONE_STREAM=$mode timeout 600 python bench.py --steps ${STEPS:-20} --warmup 3 --batch ${BATCH:-4} --no-cpu-baseline > gpurun_out/bench_iter_$mode.json 2> gpurun_out/bench_iter.err
  tail -3 gpurun_out/bench_iter.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_iter_$mode.json"))
print("ONE_STREAM=$mode ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "fwd_ungraphed", d["forward_ms_ungraphed"], "launches", d["launches_per_step"])
if $mode == 0:
    for k,v in d["families"].items(): print(f"  {k:20s} {v['ms']:8.3f} ms  {v['launches']:4d} launches  {v['tflops']:8.1f} TF/s  {v['gbs']:8.1f} GB/s")
PY
done
