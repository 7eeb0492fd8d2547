#!/bin/bash
# A/B of two builds of the library on the bench line (same box, back to back): libmmdiff_old.so vs libmmdiff.so
mkdir -p gpurun_out
for tag in old new old new; do
  lib=/root/repo/mm_diffusion_b200/libmmdiff.so
  [ $tag = old ] && lib=/root/repo/mm_diffusion_b200/libmmdiff_old.so
  MMD_LIB=$lib timeout 300 python bench.py --steps ${STEPS:-30} --warmup 3 --no-cpu-baseline --profile-reps 2 > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  TAG=$tag python - <<'PY'
import json,os
t=os.environ["TAG"]
d=json.load(open(f"gpurun_out/ab_{t}.json"))
f=d["families"]
print(t, "ms/step", d["ms_per_step"], "value", d["value"], "| qkv", f["conv1x1_qkv"]["ms"], "out", f["conv1x1_out"]["ms"], "proj", f["conv1x1_proj"]["ms"], "3x3", f["conv3x3_spatial"]["ms"], "tconv", f["conv_temporal"]["ms"], "audio", f["conv_audio_k3"]["ms"], "gn", f["group_norm"]["ms"])
PY
done
