#!/bin/bash
# One GPU-box session (run under gpurun): any subset of
#   TESTS="<pytest args>"      e.g. TESTS="tests/test_ops_gpu.py -k fused_gn" (default: nothing)
#   SMOKE=1                    __graft_entry__.smoke()
#   VARIANTS="tag:ENV=V,ENV2=V2 tag2: ..."   one sampling bench line per variant (env applied to bench.py), summarised
#   STEPS_DUMP="tag:ENV=V ..." per-launch profile of one forward (tools/gpu_profile_steps.py) per variant
#   EXTRA="<shell command>"    anything else, run last
# Everything lands in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
if [ -n "$TESTS" ]; then
  timeout ${TEST_TIMEOUT:-1200} python -m pytest $TESTS -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -${TEST_TAIL:-15} | tee $O/session_tests.log
fi
if [ -n "$SMOKE" ]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/session_smoke.log
fi
for v in $VARIANTS; do
  tag=${v%%:*}; envs=$(echo "${v#*:}" | tr ',' ' ')
  env $envs timeout 600 python bench.py --steps ${STEPS:-20} --warmup 3 --batch ${BATCH:-4} --no-cpu-baseline --no-gpu-baseline --no-extras --profile-reps 2 \
      > $O/bench_$tag.json 2> $O/bench_$tag.err || tail -3 $O/bench_$tag.err
  python tools/bench_summary.py $O/bench_$tag.json $tag
done
for v in $STEPS_DUMP; do
  tag=${v%%:*}; envs=$(echo "${v#*:}" | tr ',' ' ')
  env $envs timeout 300 python tools/gpu_profile_steps.py ${BATCH:-4} > /dev/null 2> $O/steps_$tag.err && mv $O/steps_b${BATCH:-4}.txt $O/steps_$tag.txt
done
if [ -n "$EXTRA" ]; then bash -c "$EXTRA"; fi
du -sh $O
