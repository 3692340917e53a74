#!/bin/bash
# scripts/gpu_bench_all.sh — under gpurun: phase timeline of the c2 kernel + one bench line per secondary workload
# (outputs gpurun_out/bench_<workload>.json, picked up by scripts/summarize_profiles.py).
mkdir -p gpurun_out
PPO_B200_PHASE_DEBUG=1 timeout -k 10 120 python scripts/phase_debug.py > gpurun_out/phase_c2.log 2>&1; echo "phase rc=$?"; tail -22 gpurun_out/phase_c2.log
for W in ${@:-c3 c4 c4bf16 c5 adam gather c1}; do
  timeout -k 10 500 python bench.py --workload $W --steps 5 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  echo "$W rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$W.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("$W", "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], r.get("kernel"), r.get("achieved"), r.get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    print(json.dumps(d.get("kernels"))[:1500])
except Exception as e:
    print("$W parse failed", e)
PY
done
