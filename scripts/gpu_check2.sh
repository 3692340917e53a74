#!/bin/bash
# scripts/gpu_check2.sh <tag> — phase timeline + GPU tests + bench (with secondary)
TAG=${1:-x}
mkdir -p gpurun_out
PPO_B200_PHASE_DEBUG=1 timeout -k 10 120 python scripts/phase_debug.py > gpurun_out/${TAG}_phase.log 2>&1; echo "phase rc=$?"; cat gpurun_out/${TAG}_phase.log | tail -8
timeout -k 10 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout -k 10 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
echo "bench rc=$?"; tail -c 400 gpurun_out/${TAG}_bench_c2.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_c2.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], "return", d.get("mean_episode_return"))
print(json.dumps(d["kernels"]))
print(json.dumps(d.get("cpu_baseline")))
print(json.dumps(d.get("secondary"), indent=0)[:3000])
PY
