#!/bin/bash
TAG=${1:-r2k}
timeout 120 python tools/debug_tc.py > gpurun_out/${TAG}_tc.log 2>&1; echo "debug_tc rc=$?"; tail -6 gpurun_out/${TAG}_tc.log | cut -c1-250
timeout 600 python -m pytest tests/test_gpu_parallel.py tests/test_gpu_conditioning.py tests/test_gpu_large.py -m gpu -q 2>&1 | tail -5
run() {
  timeout 600 env $3 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline $2 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$1.json").read().strip().splitlines()[-1])
    print("$1", "value %.3e e2e %.3e ms %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), "fused", round(d["stages"]["fused"]["ms_per_step"],2), "tens", round(d["stages"]["tens"]["ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), "parity", {k:v for k,v in d["parity"].items() if k not in ("random_entries",)})
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/${TAG}_$1.err").read()[-1500:])
PY
}
run cfg3_tc "--workload cfg3" ""
run cfg5_tc "--workload cfg5" ""
