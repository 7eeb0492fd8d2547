#!/bin/bash
mkdir -p gpurun_out
for b in 1 2 8; do
  timeout 600 python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_b$b.json"))
print("B=$b ms/step", d["ms_per_step"], "per-sample ms", d["ms_per_step"]/$b, "value", d["value"])
print({k:(v["ms"],v["launches"]) for k,v in d["families"].items()})
PY
done
