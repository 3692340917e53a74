#!/bin/bash
# scripts/gpu_check.sh <tag> [pytest -k expr]  — under gpurun: GPU tests, then the default bench line (c2) with its kernels.
TAG=${1:-x}; K=${2:-}
mkdir -p gpurun_out
if [ -n "$K" ]; then timeout -k 10 900 python -m pytest tests -m gpu -q -x -s -k "$K" > gpurun_out/${TAG}_pytest.log 2>&1; else timeout -k 10 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/${TAG}_pytest.log 2>&1; fi
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout -k 10 400 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
echo "bench rc=$?"; tail -c 400 gpurun_out/${TAG}_bench_c2.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_c2.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], "return", d.get("mean_episode_return"))
print(json.dumps(d["kernels"]))
print(json.dumps({k: v for k, v in (d.get("roofline") or {}).items() if k != "per_kernel"}))
PY
