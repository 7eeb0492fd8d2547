#!/bin/bash
# A few ncu --set full captures of one production forward, selected by kernel-name regex (run under gpurun).
#   SPECS="tag|regex|skip|count|ENV1=V,ENV2=V  tag2|..."   (fields separated by '|', specs by spaces; env optional)
# The regex matches the demangled name (template arguments included).  Summaries (tools/ncu_summarize.py) land in
# gpurun_out/full_<tag>_summary.txt, the per-instruction source page in full_<tag>_src.csv.gz; reports over 12 MB are dropped.
mkdir -p gpurun_out
O=gpurun_out
for spec in $SPECS; do
  IFS='|' read -r tag regex skip count envs <<< "$spec"
  envs=$(echo "$envs" | tr ',' ' ')
  timeout 600 env $envs ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$regex" -s ${skip:-0} -c ${count:-1} \
      -o $O/full_$tag -f python tools/gpu_ncu_forward.py 4 > $O/ncu_full_$tag.log 2>&1
  tail -1 $O/ncu_full_$tag.log
  python tools/ncu_summarize.py $O/full_$tag.ncu-rep $O/full_${tag}_summary.txt > /dev/null 2>&1
  ncu -i $O/full_$tag.ncu-rep --page source --csv 2>/dev/null | gzip > $O/full_${tag}_src.csv.gz
  find $O -name "full_$tag.ncu-rep" -size +12M -delete
done
ls -la $O | grep full_
