#!/bin/bash
# round 2, call R: tcgen05 Kuf with the padding blocks of the last tile skipped -- own check, full GPU suite, cfg3 / cfg5 lines
TAG=${1:-r2r}
timeout 120 python tools/debug_tc.py > gpurun_out/${TAG}_tc.log 2>&1; rc=$?; echo "debug_tc rc=$rc"; tail -5 gpurun_out/${TAG}_tc.log | cut -c1-250; if [ $rc -ne 0 ]; then echo "tcgen05 kernel failed its own check: stopping"; exit 1; fi
timeout 1800 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/${TAG}_gputests_full.log 2>&1
grep -E "AssertionError: |Error|passed|failed" gpurun_out/${TAG}_gputests_full.log | sort | uniq -c | sort -rn | head -12
for wl in cfg3 cfg5; do
  timeout 900 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_$wl.json").read().strip().splitlines()[-1])
    print("$wl", "value %.3e e2e %.3e ms %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), "stages", {k:round(v["ms_per_step"],3) for k,v in d["stages"].items() if v["ms_per_step"]>0}, "roof", d["roofline"]["bound"], round(d["roofline"]["frac"],3), "parity", d["parity"])
except Exception as e:
    print("$wl FAILED", e); print(open("gpurun_out/${TAG}_bench_$wl.err").read()[-1500:])
PY
done
timeout 600 python tools/bench_configs.py --cfg op > gpurun_out/${TAG}_op.jsonl 2> gpurun_out/${TAG}_op.err; cut -c1-400 gpurun_out/${TAG}_op.jsonl; tail -3 gpurun_out/${TAG}_op.err
