#!/bin/bash
TAG=${1:-r2f}
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gputests_full.log 2>&1
grep -E "AssertionError: |Error|passed|failed" gpurun_out/${TAG}_gputests_full.log | sort | uniq -c | sort -rn | head -40
run() { # name, extra args, env
  timeout 600 env $3 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-pipeline $2 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$1.json").read().strip().splitlines()[-1])
    print("$1", "value %.3e e2e %.3e ms %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), "fused", round(d["stages"]["fused"]["ms_per_step"],2), "tens", round(d["stages"]["tens"]["ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), "parity", d["parity"])
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/${TAG}_$1.err").read()[-1500:])
PY
}
run rbf12 "" ""
run rbf8regs "" "GPSIG_WARPFUSED_WARPS=8"
run lin12 "--kernel linear" ""
run lin8regs "--kernel linear" "GPSIG_WARPFUSED_WARPS=8"
run cfg2 "--workload cfg2" ""
run cfg3 "--workload cfg3" ""
run cfg5 "--workload cfg5" ""
run cfg1 "--workload cfg1" ""
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sigkern_warpfused -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_wf_rbf python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-pipeline > gpurun_out/${TAG}_ncu_wf_rbf.log 2>&1
