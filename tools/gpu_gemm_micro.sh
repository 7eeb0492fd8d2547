#!/bin/bash
# tools/gpu_gemm_micro.py once per kernel variant (run under gpurun):  VARIANTS="tag:ENV=V,ENV2=V ..." (default: base only)
mkdir -p gpurun_out
for v in ${VARIANTS:-base:}; do
  tag=${v%%:*}; envs=$(echo "${v#*:}" | tr ',' ' ')
  env $envs timeout 300 python tools/gpu_gemm_micro.py ${REPS:-200} > gpurun_out/micro_$tag.jsonl 2> gpurun_out/micro_$tag.err || tail -3 gpurun_out/micro_$tag.err
done
python - <<'PY'
import glob, json, os
runs = {}
for f in sorted(glob.glob("gpurun_out/micro_*.jsonl")):
    tag = os.path.basename(f)[6:-6]
    runs[tag] = {r["shape"]: r for r in map(json.loads, open(f)) }
tags = list(runs)
shapes = list(next(iter(runs.values())).keys()) if runs else []
print("shape".ljust(22) + "".join(t.rjust(9) for t in tags) + "   us@tensor  us@hbm")
for s in shapes:
    r0 = runs[tags[0]][s]
    print(s.ljust(22) + "".join((f"{runs[t][s]['us']:.1f}" if s in runs[t] else "-").rjust(9) for t in tags) + f"   {r0['us_tensor_peak']:8.1f} {r0['us_hbm_peak']:8.1f}")
PY
