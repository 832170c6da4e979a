#!/bin/bash
TAG=${1:-r2p}
timeout 120 python tools/debug_tc.py > gpurun_out/${TAG}_tc.log 2>&1; rc=$?; echo "debug_tc rc=$rc"; tail -5 gpurun_out/${TAG}_tc.log | cut -c1-250; if [ $rc -ne 0 ]; then echo "tcgen05 kernel failed its own check: stopping"; exit 1; fi
timeout 1800 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/${TAG}_gputests_full.log 2>&1
grep -E "AssertionError: |Error|passed|failed" gpurun_out/${TAG}_gputests_full.log | sort | uniq -c | sort -rn | head -20
run() {
  timeout 900 env $3 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline $2 > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$1.json").read().strip().splitlines()[-1])
    print("$1", "value %.3e e2e %.3e ms %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), "fused", round(d["stages"]["fused"]["ms_per_step"],2), "tens", round(d["stages"]["tens"]["ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), "parity", {k:(v if not isinstance(v,dict) else v.get("max_abs_err_over_max_abs_ref")) for k,v in d["parity"].items()})
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/${TAG}_$1.err").read()[-1500:])
PY
}
run cfg3 "--workload cfg3" ""
run cfg5 "--workload cfg5" ""
run cfg4 "" ""
run cfg4_linear "--kernel linear" ""
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tens_seq_tc -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_tc python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_tc.log 2>&1
