#!/bin/bash
TAG=${1:-r2j}
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gputests_full.log 2>&1
grep -E "AssertionError: |Error|passed|failed" gpurun_out/${TAG}_gputests_full.log | sort | uniq -c | sort -rn | head -30
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
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
run cfg3_cc "--workload cfg3" "GPSIG_TENS_TC=0"
run cfg5_tc "--workload cfg5" ""
run cfg5_cc "--workload cfg5" "GPSIG_TENS_TC=0"
run cfg4 "" ""
