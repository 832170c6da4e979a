#!/bin/bash
TAG=${1:-r2g}
timeout 300 python tools/debug_l45.py > gpurun_out/${TAG}_debug.log 2>&1; cat gpurun_out/${TAG}_debug.log | tail -30
run() {
  timeout 600 env $3 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-pipeline $2 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$1.json").read().strip().splitlines()[-1])
    print("$1", "value %.3e e2e %.3e ms %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), "fused", round(d["stages"]["fused"]["ms_per_step"],2), "frac", round(d["roofline"]["frac"],3), "parity", d["parity"].get("random_entries",{}).get("max_abs_err_over_max_abs_ref"))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/${TAG}_$1.err").read()[-1500:])
PY
}
run rbf12 "" ""
run rbf8regs "" "GPSIG_WARPFUSED_WARPS=8"
run lin12 "--kernel linear" ""
