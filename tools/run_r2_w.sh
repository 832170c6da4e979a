#!/bin/bash
# round 2, call W: the clock sampler started before warm-up -- does the device-timed step still carry stray milliseconds?
TAG=${1:-r2w}
for i in 1 2; do
for a in "" "--kernel linear" "--workload cfg5" "--workload cfg1"; do
  timeout 600 python bench.py --no-cpu-baseline --no-pipeline $a > gpurun_out/${TAG}_b.json 2> gpurun_out/${TAG}_b.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_b.json").read().strip().splitlines()[-1])
    print("[$a]", "ms %.3f e2e_ms %.3f"%(d["ms_per_step"], d["e2e"]["ms_per_step"]), "kernel", round(d["stages"]["fused"]["ms_per_step"]+d["stages"]["tens"]["ms_per_step"],3), "clocks", d["clocks"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/${TAG}_b.err").read()[-1500:])
PY
done
done
