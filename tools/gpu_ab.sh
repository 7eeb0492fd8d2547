#!/bin/bash
# A/B of two library builds in one box session (alternating, to cancel clock / thermal drift):
#   A = mm_diffusion_b200/libmmdiff.so, B = $1 (path to another build)
mkdir -p gpurun_out
B=${1:-$PWD/mm_diffusion_b200/libmmdiff_prev.so}
for rep in 1 2 3; do
  for which in A B; do
    if [ $which = A ]; then unset MMD_LIB; else export MMD_LIB=$B; fi
    timeout 600 python bench.py --steps 30 --warmup 5 --batch ${BATCH:-4} --no-cpu-baseline --profile-reps 2 > gpurun_out/ab_$which.json 2> gpurun_out/ab_$which.err
    W=$which python - <<'PY'
import json,os
w=os.environ["W"]; d=json.load(open(f"gpurun_out/ab_{w}.json")); f=d["families"]
print(w, "ms/step", d["ms_per_step"], "| 3x3", f["conv3x3_spatial"]["ms"], "qkv", f["conv1x1_qkv"]["ms"], "out", f["conv1x1_out"]["ms"], "proj", f["conv1x1_proj"]["ms"], "tconv", f["conv_temporal"]["ms"], "audio", f["conv_audio_k3"]["ms"], "gn", f["group_norm"]["ms"])
PY
  done
done
