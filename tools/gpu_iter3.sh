#!/bin/bash
# full GPU test suite + smoke + one bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/iter.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> gpurun_out/iter.log
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
for ce in $COMPARE_ENV; do
  env $ce timeout 600 python bench.py --steps ${STEPS:-20} --warmup 3 --batch ${BATCH:-4} --no-cpu-baseline > gpurun_out/bench_cmp.json 2> gpurun_out/bench_cmp.err
  CE=$ce python - <<'PY'
import json,os
d=json.load(open("gpurun_out/bench_cmp.json"))
f=d["families"]
print(os.environ.get("CE"), "-> ms/step", d["ms_per_step"], "value", d["value"], "| 3x3", f["conv3x3_spatial"]["ms"], "qkv", f["conv1x1_qkv"]["ms"], "out", f["conv1x1_out"]["ms"], "proj", f["conv1x1_proj"]["ms"], "tconv", f["conv_temporal"]["ms"], "audio", f["conv_audio_k3"]["ms"])
PY
done
