#!/bin/bash
# scripts/gpu_multi.sh <N> [workloads...] — under `gpurun --gpus N`: DP equivalence test on all N GPUs, then one bench line per workload at N GPUs.
N=${1:-2}; shift
mkdir -p gpurun_out
if [ -z "${SKIP_DIST:-}" ]; then timeout -k 10 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x -s > gpurun_out/dist_equivalence_${N}gpu.txt 2>&1; echo "dist rc=$?"; tail -5 gpurun_out/dist_equivalence_${N}gpu.txt; fi
for W in ${@:-c2}; do
  timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $W --steps 5 --warmup 3 --no-secondary > gpurun_out/bench_${W}_g$N.json 2> gpurun_out/bench_${W}_g$N.err
  echo "$W g$N rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${W}_g$N.json").read().strip().splitlines()[-1])
    print("$W g$N", "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "e2e ms %.3f" % d["e2e"].get("ms_per_step", 0))
    print(json.dumps(d.get("kernels"))[:800])
except Exception as e:
    print("$W parse failed", e)
PY
  tail -c 300 gpurun_out/bench_${W}_g$N.err
done
