#!/bin/bash
TAG=${1:-r2e}
timeout 1800 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/${TAG}_gputests_full.log
grep -E "AssertionError: |Error|passed|failed" gpurun_out/${TAG}_gputests_full.log | sort | uniq -c | sort -rn | head -60
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
timeout 600 python bench.py --kernel linear --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1_linear.json 2>> gpurun_out/${TAG}_bench_n1.err
python - <<PY
import json
for f in ("n1","n1_linear"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.3e e2e %.3e ms %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), "stages", {k:round(v["ms_per_step"],3) for k,v in d["stages"].items()}, "parity", d["parity"], "pipeline ms", (d.get("pipeline") or {}).get("ms_per_step"), "clocks", d.get("clocks"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -5 gpurun_out/${TAG}_bench_n1.err
