#!/bin/bash
# scripts/gpu_profile.sh <workload> <kernel-regex> <skip> <count> [launch-list-skip launch-list-count]  — run under gpurun.
# Plain run first (must exit 0), then (optionally) the ncu launch list of one step and one `ncu --set full`
# capture of the kernels matching the regex (recipe: /opt/skills/guides/B200_PROFILING.md).
# ncu costs ~0.1 s per profiled launch: keep the counts small.  Outputs land in gpurun_out/.
set -u
W=$1; K=$2; S=$3; N=$4; LS=${5:-}; LN=${6:-}
CMD="python bench.py --workload $W --steps 1 --warmup 1 --only-value"
TAG=${K%%|*}
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$W.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$W.log; exit 1; }
cat gpurun_out/plain_$W.log | tail -1
if [ -n "$LN" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -s $LS -c $LN --csv --log-file gpurun_out/launches_$W.csv $CMD > gpurun_out/ncu_launches_$W.log 2>&1
  echo "launch list rc=$?"
fi
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c $N -f -o gpurun_out/prof_${W}_${TAG} $CMD > gpurun_out/ncu_full_$W.log 2>&1
echo "full capture rc=$?"
tail -2 gpurun_out/ncu_full_$W.log
