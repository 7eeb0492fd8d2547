#!/bin/bash
# iteration loop on the GPU box: selected op tests, forward parity, one bench line
mkdir -p gpurun_out
: > gpurun_out/iter.log
if [ -n "$OPS" ]; then
  for k in $OPS; do
    timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 120 -k "$k" 2>&1 | tail -15 >> gpurun_out/iter.log
  done
fi
timeout 900 python -m pytest tests/test_forward_gpu.py ${EXTRA_TESTS} -m gpu -q -s --timeout 600 2>&1 | grep -E "rel-L2|passed|failed|Error|error" >> gpurun_out/iter.log
cat gpurun_out/iter.log
timeout 600 python bench.py --steps ${STEPS:-20} --warmup 3 --batch ${BATCH:-4} --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_iter.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "fwd_ungraphed", d["forward_ms_ungraphed"], "launches", d["launches_per_step"])
for k,v in d["families"].items(): print(f"  {k:20s} {v['ms']:8.3f} ms  {v['launches']:4d} launches  {v['tflops']:8.1f} TF/s  {v['gbs']:8.1f} GB/s")
print(d["clocks"])
PY
