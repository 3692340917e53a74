#!/bin/bash
# scripts/gpu_profile.sh <workload> <kernel-regex> <skip> [count]  — run under gpurun.
# Plain run first (must exit 0), then the launch list and one `ncu --set full` capture of the kernel
# (recipe: /opt/skills/guides/B200_PROFILING.md).  Outputs land in gpurun_out/.
set -u
W=$1; K=$2; S=$3; N=${4:-2}
CMD="python bench.py --workload $W --steps 1 --warmup 3 --cpu-iters 1"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$W.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$W.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_$W.csv $CMD > gpurun_out/ncu_launches_$W.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $N -f -o gpurun_out/prof_${W}_${K%%|*} $CMD > gpurun_out/ncu_full_$W.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_full_$W.log
